import sys, time
sys.path.insert(0,'/root/repo')
import numpy as np, torch
from jubjub_schnorr_b200 import BatchVerifier, workload as wl
bv=BatchVerifier([0])
n=1<<20
pks,off,sig,msg,exp,_=wl.make_aggregate_batch(bv,n,0.05)
hp=[torch.from_numpy(x).pin_memory() for x in (pks,sig,msg)]
for it in range(3):
    t=time.perf_counter(); st=bv.verify_aggregate(hp[0].numpy(),off,hp[1].numpy(),hp[2].numpy()); print('host call',time.perf_counter()-t, (st==exp).all())
t=time.perf_counter(); x=np.ascontiguousarray(off,dtype=np.uint32); print('ascontig',time.perf_counter()-t)
