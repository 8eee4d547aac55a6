"""One context over two devices (the single-process sharding of the host entry points): statuses, challenges and the\naccept bitmap against the expectation known by construction.  Run on a box with two GPUs: python tools/check_two_devices.py"""
import sys, numpy as np
import os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from jubjub_schnorr_b200 import BatchVerifier
from jubjub_schnorr_b200 import workload as wl
with BatchVerifier([0]) as g:
    data = {}
    for variant, n in ((0, 300001), (1, 70001), (2, 65537), (0, 33), (0, 1)):
        data[(variant, n)] = wl.make_batch(g, variant, n, 0.2, seed=variant * 7 + n)
    agg = wl.make_aggregate_batch(g, 50001, 0.05, seed=5)
    typed = wl.make_typed_single_batch(g, 150001, 0.10)
with BatchVerifier([0, 1]) as bv:
    assert bv._lib.jjs_device_count(bv._ctx) == 2
    for (variant, n), (pk, sig, msg, exp, _) in data.items():
        ver = {0: bv.verify_single, 1: bv.verify_double, 2: bv.verify_vargen}[variant]
        for rep in range(2):
            st, c = ver(pk, sig, msg, True)
            assert np.array_equal(st, exp), (variant, n)
        if variant == 0:
            assert np.array_equal(bv.unpack_bitmap(bv.verify_batch(pk, sig, msg), n), exp == 0), n
        print("ok", variant, n, flush=True)
    pks, off, sig, msg, exp, _ = agg
    st = bv.verify_aggregate(pks, off, sig, msg)
    assert np.array_equal(st, exp)
    print("ok aggregate")
    pts, u, msg, exp = typed
    for rep in range(2):
        assert np.array_equal(bv.verify_ext(0, pts, u, msg), exp)
    print("ok typed")
print("multi-device ok")
