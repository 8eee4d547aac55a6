"""Randomised size / variant / invalid-fraction stress of the host, device-pointer and bitmap entry points against the
expectation known by construction (sizes around the warp, slice, sub-chunk and scratch-half boundaries).  Run on a GPU box:
    python tools/stress_sizes.py"""
import sys, numpy as np, torch
import os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from jubjub_schnorr_b200 import BatchVerifier
from jubjub_schnorr_b200 import workload as wl
rng = np.random.default_rng(12345)
with BatchVerifier([0]) as bv:
    dev = torch.device("cuda", 0)
    for it in range(24):
        variant = int(rng.integers(0, 3))
        n = int(rng.choice([1, 2, 31, 32, 33, 127, 129, 65535, 65537, 262143, 262144, 262145, 300001, 524287, 524288, 524289, 786433, 1048577, 131071, 131073]))
        if variant != 0: n = min(n, 600000)
        frac = float(rng.choice([0.0, 0.1, 0.5, 1.0]))
        pk, sig, msg, exp, _ = wl.make_batch(bv, variant, n, frac, seed=1000 + it)
        ver = {0: bv.verify_single, 1: bv.verify_double, 2: bv.verify_vargen}[variant]
        st, c = ver(pk, sig, msg, True)
        ok_host = np.array_equal(st, exp)
        d = [torch.from_numpy(x).to(dev) for x in (pk, sig, msg)]
        d_st = torch.full((n,), 0xEE, dtype=torch.uint8, device=dev); d_c = torch.empty((n, 32), dtype=torch.uint8, device=dev)
        s = torch.cuda.Stream(dev)
        with torch.cuda.stream(s):
            bv.verify_device(variant, d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), n, d_st.data_ptr(), d_c.data_ptr(), stream=s.cuda_stream)
        s.synchronize()
        ok_dev = np.array_equal(d_st.cpu().numpy(), exp) and np.array_equal(d_c.cpu().numpy(), c)
        ok_bm = True
        if variant == 0:
            ok_bm = np.array_equal(bv.unpack_bitmap(bv.verify_batch(pk, sig, msg), n), exp == 0)
        print(it, variant, n, frac, ok_host, ok_dev, ok_bm, flush=True)
        assert ok_host and ok_dev and ok_bm
print("stress ok")
