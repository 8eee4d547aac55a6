"""Full-size (BASELINE.json) batches through the C ABI: size-independent properties instead of an item-by-item oracle
(the CPU oracle needs ~1 min per 2^20 items on 16 cores): statuses equal the expectation known by construction for
every item, a random permutation of the batch permutes the result, device-pointer and host-pointer entry points
agree, and an oracle spot check on a random sample."""
import numpy as np
import pytest

from oracle import c_oracle as co

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def bv():
    from jubjub_schnorr_b200 import BatchVerifier
    with BatchVerifier([0]) as v:
        yield v


SPOT = 1 << 14   # items per variant that go to the CPU oracle at full size


@pytest.mark.parametrize("variant,log2n", [(0, 20), (1, 20), (2, 19)])   # BASELINE.json configs[1], configs[2], the var-generator half of configs[3]
def test_full_size_batch_properties(bv, variant, log2n):
    from jubjub_schnorr_b200 import workload as wl
    n = 1 << log2n
    pk, sig, msg, expected, cls = wl.make_batch(bv, variant, n, 0.10, seed=0xB200 + variant)
    ver = {0: bv.verify_single, 1: bv.verify_double, 2: bv.verify_vargen}[variant]
    st, c = ver(pk, sig, msg, True)
    assert np.array_equal(st, expected)
    assert int((st != 0).sum()) == int(round(0.10 * n))
    assert not c[st >= 2].any() and c[st == 0].any(axis=1).all()
    # permutation equivariance (no cross-item state, ragged chunk boundaries exercised by an odd prefix)
    perm = np.random.default_rng(1).permutation(n)[: n - 12345]
    st_p, c_p = ver(pk[perm], sig[perm], msg[perm], True)
    assert np.array_equal(st_p, st[perm]) and np.array_equal(c_p, c[perm])
    # oracle spot check
    # every invalid item plus random valid ones, so that each tampering class reaches the oracle several hundred times
    rng = np.random.default_rng(2)
    bad = np.nonzero(cls >= 0)[0]
    idx = np.unique(np.concatenate([rng.choice(bad, size=SPOT // 2, replace=False), rng.choice(n, size=SPOT // 2, replace=False)]))
    over = {0: co.verify_single, 1: co.verify_double, 2: co.verify_vargen}[variant]
    st_o, c_o = over(pk[idx], sig[idx], msg[idx])
    assert np.array_equal(st_o, st[idx]) and np.array_equal(c_o, c[idx])
    assert set(np.unique(cls[idx]).tolist()) == set(range(-1, len(wl.CLASSES)))


def test_device_pointer_entry_matches_host_entry(bv):
    import torch
    from jubjub_schnorr_b200 import workload as wl
    n = 100_003
    pk, sig, msg, expected, _ = wl.make_batch(bv, 0, n, 0.2, seed=5)
    dev = torch.device("cuda", 0)
    d_pk, d_sig, d_msg = (torch.from_numpy(x).to(dev) for x in (pk, sig, msg))
    d_st = torch.empty(n, dtype=torch.uint8, device=dev)
    d_c = torch.empty((n, 32), dtype=torch.uint8, device=dev)
    stream = torch.cuda.Stream(dev)
    with torch.cuda.stream(stream):
        bv.verify_device(0, d_pk.data_ptr(), d_sig.data_ptr(), d_msg.data_ptr(), n, d_st.data_ptr(), d_c.data_ptr(), stream=stream.cuda_stream)
    stream.synchronize()
    st, c = bv.verify_single(pk, sig, msg, True)
    assert np.array_equal(d_st.cpu().numpy(), st) and np.array_equal(d_c.cpu().numpy(), c) and np.array_equal(st, expected)


@pytest.mark.parametrize("variant,n", [(0, 3 * (1 << 18) + 12345), (1, (1 << 19) + 777), (2, (1 << 18) + 1)])
def test_device_pointer_entry_overlapped_sub_chunks(bv, variant, n):
    """Device-pointer batches above one sub-chunk (2^18 items) are cut into sub-chunks that alternate between two
    internal streams and scratch halves; the caller's stream must still see the finished result.  Run twice back to
    back on the same stream (the second pass reuses both scratch halves while the first may still be draining)."""
    import torch
    from jubjub_schnorr_b200 import workload as wl
    pk, sig, msg, expected, _ = wl.make_batch(bv, variant, n, 0.1, seed=31 + variant)
    dev = torch.device("cuda", 0)
    d_pk, d_sig, d_msg = (torch.from_numpy(x).to(dev) for x in (pk, sig, msg))
    d_st = [torch.full((n,), 0xEE, dtype=torch.uint8, device=dev) for _ in range(2)]
    d_c = torch.empty((n, 32), dtype=torch.uint8, device=dev)
    stream = torch.cuda.Stream(dev)
    with torch.cuda.stream(stream):
        for k in range(2):
            bv.verify_device(variant, d_pk.data_ptr(), d_sig.data_ptr(), d_msg.data_ptr(), n, d_st[k].data_ptr(), d_c.data_ptr(), stream=stream.cuda_stream)
        first = d_st[0].clone()     # ordered after both passes on the caller's stream
    stream.synchronize()
    assert np.array_equal(first.cpu().numpy(), expected) and np.array_equal(d_st[1].cpu().numpy(), expected)
    st, c = {0: bv.verify_single, 1: bv.verify_double, 2: bv.verify_vargen}[variant](pk, sig, msg, True)
    assert np.array_equal(st, expected) and np.array_equal(d_c.cpu().numpy(), c)


def test_aggregate_workload_full_size(bv):
    """2^17 aggregate-key items signed on the GPU (signer counts 2..4), 5 % invalidated: expectation by construction,
    host and device entry points, and an oracle spot check."""
    import torch
    from jubjub_schnorr_b200 import workload as wl
    n = 1 << 17
    pks, off, sig, msg, expected, cls = wl.make_aggregate_batch(bv, n, 0.05, seed=0xA66)
    st, c, agg = bv.verify_aggregate(pks, off, sig, msg, True, True)
    assert np.array_equal(st, expected)
    sel = np.sort(np.random.default_rng(3).choice(n - 1, size=300, replace=False))
    sub_off = np.zeros(len(sel) + 1, dtype=np.uint32)
    sub_keys = []
    for j, i in enumerate(sel):
        sub_keys.append(pks[off[i]:off[i + 1]])
        sub_off[j + 1] = sub_off[j] + (off[i + 1] - off[i])
    st_o, c_o, agg_o = co.verify_aggregate(np.concatenate(sub_keys), sub_off, sig[sel], msg[sel])
    assert np.array_equal(st_o, st[sel]) and np.array_equal(c_o, c[sel])
    ok = st_o != 3
    assert np.array_equal(agg_o[ok], agg[sel][ok])
    dev = torch.device("cuda", 0)
    d = [torch.from_numpy(x).to(dev) for x in (pks, off, sig, msg)]
    d_st = torch.empty(n, dtype=torch.uint8, device=dev)
    bv.verify_aggregate_device(d[0].data_ptr(), d[1].data_ptr(), off, d[2].data_ptr(), d[3].data_ptr(), n, d_st.data_ptr(),
                               stream=torch.cuda.current_stream(dev).cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(d_st.cpu().numpy(), expected)


def _agg_subset(pks, off, sel):
    sub_off = np.zeros(len(sel) + 1, dtype=np.uint32)
    keys = []
    for j, i in enumerate(sel):
        keys.append(pks[off[i]:off[i + 1]])
        sub_off[j + 1] = sub_off[j] + (off[i + 1] - off[i])
    return np.concatenate(keys), sub_off


def test_mixed4_full_size_one_call(bv):
    """BASELINE.json configs[3]: 2^19 var-generator items (a distinct generator each) + 2^19 aggregate-key items (2-4 signers), 5 % invalid,
    through ONE jjs_verify_mixed call: expectation by construction for every item, an oracle spot check per kind, and equality with the
    per-kind entry points."""
    from jubjub_schnorr_b200 import batch as B
    from jubjub_schnorr_b200 import workload as wl
    n = 1 << 19
    pk, sig, msg, exp_v, cls_v = wl.make_batch(bv, B.VARGEN, n, 0.05, seed=0x44)
    pks, off, asig, amsg, exp_a, cls_a = wl.make_aggregate_batch(bv, n, 0.05, seed=0x45)
    res = bv.verify_mixed([(B.VARGEN, pk, sig, msg), (B.AGGREGATE, pks, asig, amsg, off)], want_challenge=True, want_bitmap=True)
    assert np.array_equal(res[0]["status"], exp_v) and np.array_equal(res[1]["status"], exp_a)
    assert np.array_equal(bv.unpack_bitmap(res[0]["bitmap"], n), exp_v == 0) and np.array_equal(bv.unpack_bitmap(res[1]["bitmap"], n), exp_a == 0)
    rng = np.random.default_rng(4)
    idx = np.unique(np.concatenate([rng.choice(np.nonzero(cls_v >= 0)[0], size=SPOT // 4, replace=False), rng.choice(n, size=SPOT // 4, replace=False)]))
    st_o, c_o = co.verify_vargen(pk[idx], sig[idx], msg[idx])
    assert np.array_equal(st_o, res[0]["status"][idx]) and np.array_equal(c_o, res[0]["c"][idx])
    sel = np.unique(np.concatenate([rng.choice(np.nonzero(cls_a >= 0)[0], size=SPOT // 8, replace=False), rng.choice(n, size=SPOT // 8, replace=False)]))
    keys, sub_off = _agg_subset(pks, off, sel)
    st_o, c_o, agg_o = co.verify_aggregate(keys, sub_off, asig[sel], amsg[sel])
    assert np.array_equal(st_o, res[1]["status"][sel]) and np.array_equal(c_o, res[1]["c"][sel])
    ok = st_o != 3
    assert np.array_equal(agg_o[ok], res[1]["aggpk"][sel][ok])
    st_v, c_v = bv.verify_vargen(pk, sig, msg, True)
    assert np.array_equal(st_v, res[0]["status"]) and np.array_equal(c_v, res[0]["c"])
    st_a, c_a, agg_a = bv.verify_aggregate(pks, off, asig, amsg, True, True)
    assert np.array_equal(st_a, res[1]["status"]) and np.array_equal(c_a, res[1]["c"]) and np.array_equal(agg_a, res[1]["aggpk"])


def test_mixed5_shape_one_call(bv):
    """BASELINE.json configs[4] at one GPU's share: a mixed single / double batch (2^20 + 2^20 items, 10 % invalid) through ONE context
    and ONE jjs_verify_mixed call, against the expectation by construction and an oracle spot check; pageable numpy buffers."""
    from jubjub_schnorr_b200 import batch as B
    from jubjub_schnorr_b200 import workload as wl
    n = 1 << 20
    pk_s, sig_s, msg_s, exp_s, cls_s = wl.make_batch(bv, B.SINGLE, n, 0.10, seed=0x51)
    pk_d, sig_d, msg_d, exp_d, cls_d = wl.make_batch(bv, B.DOUBLE, n, 0.10, seed=0x52)
    res = bv.verify_mixed([(B.SINGLE, pk_s, sig_s, msg_s), (B.DOUBLE, pk_d, sig_d, msg_d)], want_challenge=True)
    assert np.array_equal(res[0]["status"], exp_s) and np.array_equal(res[1]["status"], exp_d)
    rng = np.random.default_rng(6)
    for r, pk, sig, msg, over in ((res[0], pk_s, sig_s, msg_s, co.verify_single), (res[1], pk_d, sig_d, msg_d, co.verify_double)):
        idx = rng.choice(n, size=SPOT // 4, replace=False)
        st_o, c_o = over(pk[idx], sig[idx], msg[idx])
        assert np.array_equal(st_o, r["status"][idx]) and np.array_equal(c_o, r["c"][idx])


def test_chunk_boundaries(bv):
    """Batches that straddle the internal 2^20-item pass and the 2^18-item copy slices of the host path."""
    from jubjub_schnorr_b200 import workload as wl
    for variant, n in ((0, (1 << 20) + 3), (1, (1 << 19) + (1 << 18) + 5)):
        pk, sig, msg, expected, _ = wl.make_batch(bv, variant, n, 0.05, seed=77 + variant)
        st = {0: bv.verify_single, 1: bv.verify_double}[variant](pk, sig, msg)
        assert np.array_equal(st, expected)


def test_two_contexts_on_two_host_threads(bv):
    """Contexts are independent: two host threads, each with its own context on the same GPU, get correct results."""
    import threading
    from jubjub_schnorr_b200 import BatchVerifier
    from jubjub_schnorr_b200 import workload as wl
    results = {}

    def work(tag, variant):
        with BatchVerifier([0]) as v:
            pk, sig, msg, expected, _ = wl.make_batch(v, variant, 60_000, 0.2, seed=900 + tag)
            for _ in range(3):
                st = {0: v.verify_single, 2: v.verify_vargen}[variant](pk, sig, msg)
                results[(tag, _)] = np.array_equal(st, expected)

    threads = [threading.Thread(target=work, args=(i, (0, 2)[i % 2])) for i in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert len(results) == 6 and all(results.values())


def test_argument_errors_are_reported_not_fatal(bv):
    import ctypes as C
    from jubjub_schnorr_b200 import _native
    lib = _native.lib()
    st = (C.c_uint8 * 4)()
    buf = (C.c_uint8 * 256)()
    assert lib.jjs_verify_single(bv._ctx, None, buf, buf, 4, st, None) == -1
    assert b"null" in lib.jjs_last_error(bv._ctx)
    assert lib.jjs_verify_single(bv._ctx, buf, buf, buf, 0, st, None) == 0          # empty batch is fine
    assert lib.jjs_verify_single_device(bv._ctx, 7, buf, buf, buf, 4, st, None, None) == -1
    assert lib.jjs_challenge_only(bv._ctx, 9, buf, buf, buf, 1, buf) == -1
    ctx = C.c_void_p()
    dev = (C.c_int * 1)(99)
    assert lib.jjs_init(dev, 1, C.byref(ctx)) == -1 and b"not present" in lib.jjs_last_error(ctx)
    lib.jjs_destroy(ctx)


def test_one_context_over_two_devices():
    """Single-process sharding of the host entry points (contiguous shards, one set of streams per device, shards
    starting on a bitmap word).  Needs two GPUs; skipped on a one-GPU box (tools/check_two_devices.py is the same check)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from jubjub_schnorr_b200 import BatchVerifier
    from jubjub_schnorr_b200 import workload as wl
    with BatchVerifier([0]) as g:
        batches = [(v, n) + tuple(wl.make_batch(g, v, n, 0.2, seed=11 * v + n)[:4]) for v, n in ((0, 300_001), (1, 70_001), (2, 65_537), (0, 33), (0, 1))]
        pts, u, tmsg, texp = wl.make_typed_single_batch(g, 150_001, 0.10)
    with BatchVerifier([0]) as g:
        pks, off, asig, amsg, aexp, _ = wl.make_aggregate_batch(g, 90_001, 0.05, seed=0x77)
    with BatchVerifier([0, 1]) as bv:
        res = bv.verify_mixed([(0, batches[0][2], batches[0][3], batches[0][4]), (1, batches[1][2], batches[1][3], batches[1][4]), (3, pks, asig, amsg, off)])
        assert np.array_equal(res[0]["status"], batches[0][5]) and np.array_equal(res[1]["status"], batches[1][5]) and np.array_equal(res[2]["status"], aexp)
        assert np.array_equal(bv.verify_ext(0, pts, u, tmsg), texp)
        for v, n, pk, sig, msg, exp in batches:
            st, _ = {0: bv.verify_single, 1: bv.verify_double, 2: bv.verify_vargen}[v](pk, sig, msg, True)
            assert np.array_equal(st, exp), (v, n)
            if v == 0:
                assert np.array_equal(bv.unpack_bitmap(bv.verify_batch(pk, sig, msg), n), exp == 0), n
