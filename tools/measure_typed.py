#!/usr/bin/env python3
"""Throughput of the typed-input path (jjs_verify_ext) on one GPU: 2^20 single items given as
JubJubExtended coordinates in pinned host memory, end to end.  Prints one JSON line (kept in profiles/)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from jubjub_schnorr_b200 import BatchVerifier  # noqa: E402
from jubjub_schnorr_b200 import workload as wl  # noqa: E402

bv = BatchVerifier([0])
n = 1 << 20
pts, u, msg, exp = wl.make_typed_single_batch(bv, n, 0.10)
hp = [torch.from_numpy(x).pin_memory() for x in (pts, u, msg)]
for _ in range(3):
    st = bv.verify_ext(0, hp[0].numpy(), hp[1].numpy(), hp[2].numpy())
assert np.array_equal(st, exp)
steps = 5
t0 = time.perf_counter()
for _ in range(steps):
    st = bv.verify_ext(0, hp[0].numpy(), hp[1].numpy(), hp[2].numpy())
dt = time.perf_counter() - t0
assert np.array_equal(st, exp)
# stage breakdown from a second pass: with the stage timers on, the library keeps its launches on one stream
bv.profile(True)
for _ in range(steps):
    bv.verify_ext(0, hp[0].numpy(), hp[1].numpy(), hp[2].numpy())
bv.profile(False)
stages = bv.profile_collect()
print(json.dumps({"metric": "schnorr_verifications_per_second", "path": "jjs_verify_ext (typed JubJubExtended inputs, host buffers)",
                  "value": n * steps / dt, "items": n, "ms_per_step": 1e3 * dt / steps, "h2d_bytes_per_step": int(pts.nbytes + u.nbytes + msg.nbytes),
                  "stage_ms_per_step": {k: v[0] / steps for k, v in stages.items()}, "statuses_match_expectation": True}))
