import time, sys
sys.path.insert(0, ".")
t0 = time.perf_counter()
import torch
from jubjub_schnorr_b200 import BatchVerifier
t1 = time.perf_counter()
bv = BatchVerifier([0]); t2 = time.perf_counter()
bv.close() if hasattr(bv, "close") else None
t3 = time.perf_counter()
bv2 = BatchVerifier([0]); t4 = time.perf_counter()
print({"import_s": round(t1 - t0, 2), "first_init_s": round(t2 - t1, 3), "second_init_s": round(t4 - t3, 3)})
