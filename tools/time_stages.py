"""Stage times of one 2^20-item batch per variant for whichever library JJS_B200_LIB names, device-resident inputs
(CUDA events of the library's own stage timers, launches serialised).  Statuses are compared with the constructed
expectation, so an A/B build with wrong arithmetic shows up as mismatches.
usage: JJS_B200_LIB=build/lib_x.so python tools/time_stages.py [variants, default 0] [log2n, default 20]"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from jubjub_schnorr_b200 import BatchVerifier  # noqa: E402
from jubjub_schnorr_b200 import workload as wl  # noqa: E402

variants = [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "0").split(",")]
n = 1 << (int(sys.argv[2]) if len(sys.argv) > 2 else 20)
with BatchVerifier([0]) as bv:
    for variant in variants:
        pk, sig, msg, exp, cls = wl.make_batch(bv, variant, n, 0.10, seed=0xB200)
        # units of the kernels (what the ncu counts are divided by): decoded points, work-listed items, evaluated equations
        sig_only = np.array([nm.startswith("R_") and st == 2 for nm, st in wl.CLASSES], dtype=bool)
        listed = int(((exp <= 1) | ((cls >= 0) & sig_only[np.clip(cls, 0, len(wl.CLASSES) - 1)])).sum())
        units = {"k_decode": n * (2, 4, 3)[variant], "k_challenge": listed, "k_equation": listed * (1, 2, 1)[variant]}
        dev = torch.device("cuda", 0)
        d = [torch.from_numpy(x).to(dev) for x in (pk, sig, msg)]
        st = torch.empty(n, dtype=torch.uint8, device=dev)
        s = torch.cuda.current_stream(dev)
        for _ in range(3):
            bv.verify_device(variant, d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), n, st.data_ptr(), None, stream=s.cuda_stream)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(4):
            bv.verify_device(variant, d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), n, st.data_ptr(), None, stream=s.cuda_stream)
        e1.record(s)
        torch.cuda.synchronize()
        whole = e0.elapsed_time(e1) / 4
        bv.profile(True)
        for _ in range(4):
            bv.verify_device(variant, d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), n, st.data_ptr(), None, stream=s.cuda_stream)
        torch.cuda.synchronize()
        bv.profile(False)
        print(os.path.basename(os.environ.get("JJS_B200_LIB", "default")), "variant", variant, "step_ms", round(whole, 3),
              {k: round(v[0] / 4, 3) for k, v in bv.profile_collect().items()}, "mismatches", int((st.cpu().numpy() != exp).sum()), "units", units, flush=True)
