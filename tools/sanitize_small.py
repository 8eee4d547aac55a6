#!/usr/bin/env python3
"""Tiny pass over every entry point, meant to run under `compute-sanitizer --tool memcheck` (one tool per gpurun call)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np  # noqa: E402

from jubjub_schnorr_b200 import BatchVerifier  # noqa: E402
from jubjub_schnorr_b200 import workload as wl  # noqa: E402

with BatchVerifier([0]) as bv:
    for variant in (0, 1, 2):
        pk, sig, msg, exp, _ = wl.make_batch(bv, variant, 300, 0.3, seed=1)
        st, c = {0: bv.verify_single, 1: bv.verify_double, 2: bv.verify_vargen}[variant](pk, sig, msg, True)
        assert np.array_equal(st, exp), variant
        bv.challenge_only(variant, pk, sig, msg)
    pks, off, sig, msg, exp, _ = wl.make_aggregate_batch(bv, 200, 0.2, seed=2)
    st, c, agg = bv.verify_aggregate(pks, off, sig, msg, True, True)
    assert np.array_equal(st, exp)
    pts, u, msg, exp = wl.make_typed_single_batch(bv, 200, 0.2, seed=3)
    assert np.array_equal(bv.verify_ext(0, pts, u, msg), exp)
    bv.subgroup_check(pks[:64], 0)
    bv.subgroup_check(pks[:64], 1)
    assert bv.fb_table_check(0, np.arange(0, 12 << 21, 99991, dtype=np.uint32)) == 0
    assert bv.fb_table_check(1, np.arange(5, 12 << 21, 199999, dtype=np.uint32)) == 0
    # a multisig session built from the aggregate batch keys (shares are garbage: exercises the failing path)
    K = int(off[8])
    st, bad, sg, ok = bv.multisig_combine(pks[:K], pks[:K], pks[:K], np.zeros((K, 32), np.uint8), off[:9], msg[:8])
    assert (st == 5).all()
print("sanitize pass ok")
