"""Stage times of one 2^20-item single batch for whichever library JJS_B200_LIB names, with no result checks (for
timing diagnostics built with deliberately wrong arithmetic).  usage: JJS_B200_LIB=build/lib_x.so python tools/time_stages.py"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from jubjub_schnorr_b200 import BatchVerifier  # noqa: E402
from jubjub_schnorr_b200 import workload as wl  # noqa: E402

n = 1 << 20
with BatchVerifier([0]) as bv:
    pk, sig, msg, exp, _ = wl.make_batch(bv, 0, n, 0.10, seed=0xB200)
    dev = torch.device("cuda", 0)
    d = [torch.from_numpy(x).to(dev) for x in (pk, sig, msg)]
    st = torch.empty(n, dtype=torch.uint8, device=dev)
    s = torch.cuda.current_stream(dev)
    for _ in range(3):
        bv.verify_device(0, d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), n, st.data_ptr(), None, stream=s.cuda_stream)
    torch.cuda.synchronize()
    bv.profile(True)
    for _ in range(4):
        bv.verify_device(0, d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), n, st.data_ptr(), None, stream=s.cuda_stream)
    torch.cuda.synchronize()
    bv.profile(False)
    print(os.environ.get("JJS_B200_LIB", "default"), {k: round(v[0] / 4, 3) for k, v in bv.profile_collect().items()},
          "mismatches", int((st.cpu().numpy() != exp).sum()))
