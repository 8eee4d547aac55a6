"""CPU twins of the device algorithms (tests/hostsim, built from the very headers the CUDA kernels use) against the
pinned oracle: limb arithmetic and the Montgomery schedule, table square root, scaled Hades permutation, point
decoding, fixed/variable-base multiplication, Tate subgroup test, signing, and the whole verify pipeline on
adversarial batches.  The PTX blocks themselves are covered by the `-m gpu` parity tests."""
import ctypes as C
import random

import numpy as np
import pytest

import __graft_entry__ as entry
from oracle import c_oracle as co
from oracle import jjs_oracle as o
from tests import adversarial as adv

Q, R256 = o.Q, 1 << 256


@pytest.fixture(scope="module")
def L():
    return C.CDLL(entry.build_hostsim())


def arr(x, n=8):
    return (C.c_uint32 * n)(*[(x >> (32 * i)) & 0xFFFFFFFF for i in range(n)])


def val(a):
    return sum(int(v) << (32 * i) for i, v in enumerate(a))


def _cases(rnd, n):
    specials = [0, 1, Q - 1, Q - 2, 2**256 - 1, 2**255, (1 << 32) - 1, Q >> 1, R256 % Q]
    for a in specials:
        for b in specials:
            yield a, b
    for _ in range(n):
        yield rnd.getrandbits(256), rnd.getrandbits(256)


def test_wide_products_and_reduction(L):
    rnd = random.Random(1)
    rinv = pow(R256, -1, Q)
    for a, b in _cases(rnd, 4000):
        t = (C.c_uint32 * 16)()
        L.hs_mul_wide(arr(a), arr(b), t)
        assert val(t) == a * b
        L.hs_sqr_wide(arr(a), t)
        assert val(t) == a * a
    for t in [0, (Q << 256) - 1, (Q - 1) * (Q - 1), (1 << 256) - 1, Q * ((1 << 256) - 1)] + [rnd.randrange(Q << 256) for _ in range(4000)]:
        r = (C.c_uint32 * 8)()
        L.hs_redc(arr(t, 16), r)
        assert val(r) == t * rinv % Q


def test_field_ops(L):
    rnd = random.Random(2)
    rinv = pow(R256, -1, Q)
    r = (C.c_uint32 * 8)()
    for a, b in _cases(rnd, 3000):
        a %= Q
        b %= Q
        L.hs_fq_mul(arr(a), arr(b), r); assert val(r) == a * b * rinv % Q
        L.hs_fq_mul_fp(arr(a), arr(b), r); assert val(r) == a * b * rinv % Q, (hex(a), hex(b))   # FP64-pipe multiplier
        L.hs_fq_sqr(arr(a), r); assert val(r) == a * a * rinv % Q
        L.hs_fq_add(arr(a), arr(b), r); assert val(r) == (a + b) % Q
        L.hs_fq_sub(arr(a), arr(b), r); assert val(r) == (a - b) % Q
        L.hs_fq_to_mont(arr(a), r); assert val(r) == a * R256 % Q
        L.hs_fq_from_mont(arr(a), r); assert val(r) == a * rinv % Q
    for _ in range(10):
        a = rnd.randrange(1, Q)
        L.hs_fq_inv(arr(a * R256 % Q), r)
        assert val(r) == pow(a, -1, Q) * R256 % Q


def test_sqrt_ratio(L):
    rnd = random.Random(3)
    rinv = pow(R256, -1, Q)
    r = (C.c_uint32 * 8)()
    for i in range(120):
        num, den = rnd.randrange(Q), rnd.randrange(1, Q)
        if i < 5:
            num, den = [0, 1, Q - 1, 4, 9][i], [5, 1, 1, 1, 4][i]
        ok = L.hs_sqrt_ratio(arr(num * R256 % Q), arr(den * R256 % Q), r)
        ratio = num * pow(den, -1, Q) % Q
        assert bool(ok) == (ratio == 0 or pow(ratio, (Q - 1) // 2, Q) == 1)
        if ok:
            x = val(r) * rinv % Q
            assert x * x % Q == ratio


def test_hades_permutation(L):
    rnd = random.Random(4)
    for st in ([1, 2, 3, 4, 5], [0] * 5, [Q - 1] * 5, [rnd.randrange(Q) for _ in range(5)]):
        buf = (C.c_uint32 * 40)(*[(x >> (32 * i)) & 0xFFFFFFFF for x in st for i in range(8)])
        L.hs_hades_permute(buf)
        assert [val(buf[8 * i:8 * i + 8]) for i in range(5)] == o.hades_permute(st)


def test_point_decode(L):
    rnd = random.Random(5)
    encs = [rnd.randbytes(32) for _ in range(80)] + [o.point_to_bytes(o.IDENTITY), adv.ORDER2_ENC, o.le32(o.Q), o.le32(o.Q - 1),
                                                     bytes([1] + [0] * 30 + [0x80]), bytes(32), bytes([0] * 31 + [0x80])]
    for enc in encs:
        out = (C.c_uint32 * 16)()
        ok = L.hs_point_decode(enc, out)
        p = o.point_from_bytes(enc)
        assert bool(ok) == (p is not None)
        if ok:
            assert (val(out[:8]), val(out[8:])) == p


def test_fixed_base_table_entries(L):
    for which, w, j in [(0, 0, 1), (0, 0, 255), (0, 3, 17), (1, 20, 5), (1, 7, 200), (0, 20, 4095), (1, 0, 0), (0, 11, 2049)]:
        a, b = (C.c_uint32 * 24)(), (C.c_uint32 * 24)()
        L.hs_fb_entry(which, w, j, a, b)
        assert list(a) == list(b)


def test_scalar_multiplications(L):
    rnd = random.Random(6)
    out = (C.c_uint8 * 32)()
    for i in range(8):
        k = rnd.randrange(o.R_ORDER) if i > 2 else [0, 1, o.R_ORDER - 1][i]
        P = o.pmul(o.G, rnd.randrange(1, o.R_ORDER))
        enc = o.point_to_bytes(P)
        L.hs_varbase_mul(enc, o.le32(k), out, 0); assert bytes(out) == o.point_to_bytes(o.pmul(P, k))
        L.hs_varbase_mul(enc, o.le32(k), out, 1); assert bytes(out) == o.point_to_bytes(o.pmul(o.G, k))
        L.hs_varbase_mul(enc, o.le32(k), out, 2); assert bytes(out) == o.point_to_bytes(o.pmul(o.G_NUMS, k))


def test_tate_subgroup_test_equals_definition(L):
    rnd = random.Random(7)
    t8 = adv.torsion()[8]
    for _ in range(6):
        cur = o.point_to_bytes(o.pmul(o.G, rnd.randrange(1, o.R_ORDER)))
        for j in range(8):
            assert L.hs_subgroup(cur, 0) == L.hs_subgroup(cur, 1) == (1 if j == 0 else 0)
            cur = co.point_add(cur, t8)
    cur = o.point_to_bytes(o.IDENTITY)
    for j in range(8):
        assert L.hs_subgroup(cur, 0) == L.hs_subgroup(cur, 1) == (1 if j == 0 else 0)
        cur = co.point_add(cur, t8)
    for _ in range(60):
        enc = rnd.randbytes(32)
        assert L.hs_subgroup(enc, 0) == L.hs_subgroup(enc, 1)


def test_signing_twin_matches_oracle(L):
    rnd = random.Random(8)
    for variant in (0, 1, 2):
        sk, rr, m, g = rnd.randrange(o.R_ORDER), rnd.randrange(o.R_ORDER), rnd.randrange(o.Q), rnd.randrange(1, o.R_ORDER)
        pk, sig = (C.c_uint8 * 64)(), (C.c_uint8 * 96)()
        assert L.hs_sign(variant, o.le32(sk), o.le32(rr), o.le32(g), o.le32(m), pk, sig)
        if variant == 0:
            u, R = o.sign_single(sk, rr, m)
            epk, esig = o.point_to_bytes(o.pmul(o.G, sk)), o.le32(u) + o.point_to_bytes(R)
        elif variant == 1:
            u, R, Rp = o.sign_double(sk, rr, m)
            epk = o.point_to_bytes(o.pmul(o.G, sk)) + o.point_to_bytes(o.pmul(o.G_NUMS, sk))
            esig = o.le32(u) + o.point_to_bytes(R) + o.point_to_bytes(Rp)
        else:
            gen = o.pmul(o.G, g)
            u, R = o.sign_var_gen(sk, gen, rr, m)
            epk, esig = o.point_to_bytes(o.pmul(gen, sk)) + o.point_to_bytes(gen), o.le32(u) + o.point_to_bytes(R)
        assert bytes(pk)[:len(epk)] == epk and bytes(sig)[:len(esig)] == esig


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("impl", [2, 1])
@pytest.mark.parametrize("vi,kind", [(0, "single"), (1, "double"), (2, "vargen")])
def test_pipeline_twin_on_adversarial_batches(L, vi, kind, impl):
    """impl 1: the register-operand evaluation k_equation runs (var-gen: three short scalars); impl 2: the slot-based evaluation of
    csrc/fqs.cuh kept as a compile-time alternative (var-gen: two full-size scalars)."""
    L.hs_set_equation_impl(impl)
    gen = {"single": co.gen_single, "double": co.gen_double, "vargen": co.gen_vargen}[kind]
    ver = {"single": co.verify_single, "double": co.verify_double, "vargen": co.verify_vargen}[kind]
    n = 320
    pk, sig, msg = gen(0xB200, n)
    pk, sig, msg, exp, names = adv.make_adversarial(kind, pk, sig, msg, seed=7, frac=0.6)
    st, c = ver(pk, sig, msg)
    hst, hc = np.zeros(n, np.uint8), np.zeros((n, 32), np.uint8)
    L.hs_verify(vi, _p(pk), _p(sig), _p(msg), C.c_size_t(n), _p(hst), _p(hc))
    L.hs_set_equation_impl(1)
    bad = np.nonzero(hst != st)[0]
    assert bad.size == 0, [(int(i), names[i], int(hst[i]), int(st[i])) for i in bad[:5]]
    assert np.array_equal(hc, c) and np.array_equal(st, exp)


@pytest.mark.parametrize("vi,kind", [(0, "single"), (1, "double"), (2, "vargen")])
def test_special_response_scalars_twin(L, vi, kind):
    """u at the edges of the scalar decompositions in otherwise valid items (the GPU test of the same name, on the twin)."""
    special = [0, 1, 2, 7, 8, o.R_ORDER - 1, o.R_ORDER - 2, o.R_ORDER // 2, 1 << 84, 1 << 126, (1 << 126) - 1, 1 << 127, 1 << 168, (1 << 170) - 1, 1 << 250]
    gen = {"single": co.gen_single, "double": co.gen_double, "vargen": co.gen_vargen}[kind]
    ver = {"single": co.verify_single, "double": co.verify_double, "vargen": co.verify_vargen}[kind]
    n = 2 * len(special)
    pk, sig, msg = gen(0xB207, n)
    for j, u in enumerate(special):
        sig[2 * j, :32] = np.frombuffer(u.to_bytes(32, "little"), dtype=np.uint8)
    st, c = ver(pk, sig, msg)
    hst, hc = np.zeros(n, np.uint8), np.zeros((n, 32), np.uint8)
    L.hs_verify(vi, _p(pk), _p(sig), _p(msg), C.c_size_t(n), _p(hst), _p(hc))
    assert (st[0::2] == 1).all() and not st[1::2].any()
    assert np.array_equal(hst, st) and np.array_equal(hc, c)


def test_vargen_fallback_when_no_short_vector_is_found(L):
    """The var-generator equation falls back to the two-table evaluation when the lattice reduction reports no vector that fits
    its windows (unreachable with hash outputs, so the twin forces it): same statuses and challenges."""
    n = 96
    pk, sig, msg = co.gen_vargen(0xB200, n)
    pk, sig, msg, exp, names = adv.make_adversarial("vargen", pk, sig, msg, seed=9, frac=0.5)
    st, c = co.verify_vargen(pk, sig, msg)
    hst, hc = np.zeros(n, np.uint8), np.zeros((n, 32), np.uint8)
    L.hs_force_vargen_fallback(1)
    try:
        L.hs_verify(2, _p(pk), _p(sig), _p(msg), C.c_size_t(n), _p(hst), _p(hc))
    finally:
        L.hs_force_vargen_fallback(0)
    assert np.array_equal(hst, st) and np.array_equal(hc, c) and np.array_equal(st, exp)


def test_scalar_products_mod_r(L):
    """rho * u mod r by Barrett reduction (rho < 2^128) and the general bit-serial product, against big integers."""
    rnd = random.Random(9)
    r = o.R_ORDER
    out = (C.c_uint32 * 8)()
    cases = [(0, 0), (1, r - 1), ((1 << 126) - 1, r - 1), ((1 << 128) - 1, r - 1), ((1 << 127) + 12345, 1), (1 << 125, r // 2)]
    cases += [(rnd.getrandbits(126), rnd.randrange(r)) for _ in range(3000)]
    for a, b in cases:
        L.hs_fr_mul_short(arr(a, 4), arr(b), out)
        assert val(out) == a * b % r, (hex(a), hex(b))
    for a, b in [(0, 0), ((1 << 160) - 1, r - 1), (1 << 159, r - 1), (1, r - 1)] + [(rnd.getrandbits(160), rnd.randrange(r)) for _ in range(3000)]:
        L.hs_fr_mul_160(arr(a, 5), arr(b), out)
        assert val(out) == a * b % r, (hex(a), hex(b))
    for _ in range(200):
        a, b = rnd.randrange(r), rnd.randrange(r)
        L.hs_fr_mul(arr(a), arr(b), out)
        assert val(out) == a * b % r


def test_half_size_decomposition(L):
    """tau == rho * c (mod r) with |tau| < 2^130, 0 < |rho| < 2^126, the digits recompose them, and rho is odd for
    nearly every challenge (what lets the equation stand in for the subgroup test of R)."""
    rng = np.random.default_rng(5)
    cs = [0, 1, 2, o.R_ORDER - 1, o.R_ORDER - 2, 1 << 126, (1 << 126) - 1, (1 << 125) + 1, 1 << 249, (1 << 250) - 1, o.R_ORDER // 2, o.R_ORDER // 3]
    cs += [int.from_bytes(rng.bytes(32), "little") % (1 << 250) for _ in range(3000)]
    odd = 0
    for c in cs:
        tau_b, rho_b, dig = (C.c_uint8 * 20)(), (C.c_uint8 * 16)(), (C.c_int8 * 66)()
        fl = L.hs_half_gcd((C.c_uint8 * 32).from_buffer_copy(c.to_bytes(32, "little")), tau_b, rho_b, dig)
        tau, rho = int.from_bytes(bytes(tau_b), "little"), int.from_bytes(bytes(rho_b), "little")
        assert 0 < rho < (1 << 128) and tau < (1 << 130), hex(c)
        srho = -rho if fl & 1 else rho
        assert (srho * c - tau) % o.R_ORDER == 0, hex(c)
        assert bool(fl & 2) == bool(rho & 1)
        odd += (fl >> 1) & 1
        d = list(dig)
        assert all(-8 <= x <= 8 for x in d)
        assert sum(x << (4 * i) for i, x in enumerate(d[:33])) == (-tau if fl & 1 else tau)
        assert sum(x << (4 * i) for i, x in enumerate(d[33:])) == -rho
    assert odd > 0.93 * len(cs), odd


def test_three_short_scalars(L):
    """lattice3_reduce: x == z u and y == z c (mod r) with z != 0 and all three below 2^170, the 43 digits recompose them, a
    vector is found for every random challenge and z is odd for nearly all of them; degenerate inputs either give a valid
    triple or report 'none' (the kernel then evaluates the equation directly)."""
    rng = np.random.default_rng(11)
    r = o.R_ORDER
    special = [0, 1, 2, r - 1, r - 2, 1 << 126, (1 << 126) - 1, (1 << 125) + 1, 1 << 249, (1 << 250) - 1, r // 2, r // 3, 1 << 84, 1 << 168]
    cases = [(u, c) for u in special for c in special if c < (1 << 250) or c in (r - 1, r - 2, r // 2, r // 3)]
    n_special = len(cases)
    cases += [(int.from_bytes(rng.bytes(32), "little") % r, int.from_bytes(rng.bytes(32), "little") % (1 << 250)) for _ in range(3000)]
    found = odd = 0
    for k, (u, c) in enumerate(cases):
        xb, yb, zb, fl, dig = (C.c_uint8 * 32)(), (C.c_uint8 * 32)(), (C.c_uint8 * 32)(), C.c_int(0), (C.c_int8 * 192)()
        ok = L.hs_lattice3((C.c_uint8 * 32).from_buffer_copy(u.to_bytes(32, "little")), (C.c_uint8 * 32).from_buffer_copy(c.to_bytes(32, "little")),
                           xb, yb, zb, C.byref(fl), dig)
        if not ok:
            assert k < n_special, (hex(u), hex(c))     # only degenerate inputs may come back empty
            continue
        f = fl.value
        x, y, z = (int.from_bytes(bytes(b), "little") for b in (xb, yb, zb))
        assert max(x, y, z) < (1 << 170) and z != 0, (hex(u), hex(c))
        sx, sy, sz = (-x if f & 1 else x), (-y if f & 2 else y), (-z if f & 4 else z)
        assert (sx - sz * u) % r == 0 and (sy - sz * c) % r == 0, (hex(u), hex(c))
        assert bool(f & 8) == bool(z & 1)
        d = list(dig)
        assert all(-8 <= v <= 8 for v in d)
        for j, want in enumerate((sx, sy, -sz)):
            assert sum(v << (4 * i) for i, v in enumerate(d[64 * j:64 * j + 43])) == want
            assert not any(d[64 * j + 43:64 * j + 64])
        if k >= n_special:
            found += 1
            odd += (f >> 3) & 1
    assert found == 3000 and odd > 0.99 * found, (found, odd)


@pytest.mark.parametrize("vi,kind", [(0, "single"), (1, "double"), (2, "vargen")])
def test_torsion_shifted_signatures_are_invalid_points(L, vi, kind):
    """Forgeries R = r*B + T (T of small order) by the key holder: the equation holds up to torsion, so only the
    subgroup test of R rejects them.  The twin must run that test whenever the equation did not settle it."""
    ver = {"single": co.verify_single, "double": co.verify_double, "vargen": co.verify_vargen}[kind]
    n = 48
    pk, sig, msg, exp = adv.torsion_shifted_signatures(kind, n, seed=11)
    st, c = ver(pk, sig, msg)
    assert np.array_equal(st, exp)
    hst, hc = np.zeros(n, np.uint8), np.zeros((n, 32), np.uint8)
    eqs, rts = C.c_int(), C.c_int()
    L.hs_rtest_counters(C.byref(eqs), C.byref(rts))
    L.hs_verify(vi, _p(pk), _p(sig), _p(msg), C.c_size_t(n), _p(hst), _p(hc))
    L.hs_rtest_counters(C.byref(eqs), C.byref(rts))
    assert np.array_equal(hst, st) and np.array_equal(hc, c)
    # every shifted point was tested explicitly (the un-shifted R' of a double item may be settled by its equation)
    assert rts.value >= n


def test_deferred_subgroup_test_is_rare_on_valid_batches(L):
    n = 96
    pk, sig, msg = co.gen_single(0xB200, n)
    hst, hc = np.zeros(n, np.uint8), np.zeros((n, 32), np.uint8)
    eqs, rts = C.c_int(), C.c_int()
    L.hs_rtest_counters(C.byref(eqs), C.byref(rts))
    L.hs_verify(0, _p(pk), _p(sig), _p(msg), C.c_size_t(n), _p(hst), _p(hc))
    L.hs_rtest_counters(C.byref(eqs), C.byref(rts))
    assert not hst.any() and eqs.value == n and rts.value <= n // 8


@pytest.mark.parametrize("vi,kind", [(0, "single"), (1, "double"), (2, "vargen")])
def test_pipeline_twin_on_bitflip_fuzz(L, vi, kind):
    gen = {"single": co.gen_single, "double": co.gen_double, "vargen": co.gen_vargen}[kind]
    ver = {"single": co.verify_single, "double": co.verify_double, "vargen": co.verify_vargen}[kind]
    n = 160
    pk, sig, msg = gen(0xF022, n)
    pk, sig, msg = adv.bitflip_fuzz(pk, sig, msg, seed=5)
    st, c = ver(pk, sig, msg)
    hst, hc = np.zeros(n, np.uint8), np.zeros((n, 32), np.uint8)
    L.hs_verify(vi, _p(pk), _p(sig), _p(msg), C.c_size_t(n), _p(hst), _p(hc))
    assert np.array_equal(hst, st) and np.array_equal(hc, c)


def test_aggregate_twin(L):
    signers = [1, 2, 3, 4, 2, 3, 5, 2]
    pks, off, sig, msg = co.gen_aggregate(11, signers)
    sig[2, 0] ^= 1
    pks[off[3]] = np.frombuffer(adv.torsion()[4], np.uint8)
    pks[off[4]] = np.frombuffer(adv.off_curve_encoding(np.random.default_rng(1)), np.uint8)
    st, c, agg = co.verify_aggregate(pks, off, sig, msg)
    n = len(signers)
    hst, hc, ha = np.zeros(n, np.uint8), np.zeros((n, 32), np.uint8), np.zeros((n, 32), np.uint8)
    L.hs_verify_aggregate(_p(pks), _p(off), _p(sig), _p(msg), C.c_size_t(n), _p(hst), _p(hc), _p(ha))
    assert np.array_equal(hst, st) and np.array_equal(hc, c)
    ok = st != 3
    assert np.array_equal(ha[ok], agg[ok])


def test_aggregate_twin_has_no_signer_limit(L):
    """65 and 200 signers (transcripts of 132 and 402 elements: SAFE tags from the host-side BLAKE2b, csrc/safe_tag.h), an empty
    signer set (identity aggregate: InvalidPoint) and a tampered many-signer item."""
    signers = [65, 0, 200, 66]
    pks, off, sig, msg = co.gen_aggregate(29, signers)
    sig[3, 1] ^= 4
    st, c, agg = co.verify_aggregate(pks, off, sig, msg)
    assert st.tolist() == [0, 2, 0, 1]
    n = len(signers)
    hst, hc, ha = np.zeros(n, np.uint8), np.zeros((n, 32), np.uint8), np.zeros((n, 32), np.uint8)
    L.hs_verify_aggregate(_p(pks), _p(off), _p(sig), _p(msg), C.c_size_t(n), _p(hst), _p(hc), _p(ha))
    assert np.array_equal(hst, st) and np.array_equal(hc, c) and np.array_equal(ha, agg)


def test_safe_tags_match_hashlib(L):
    out = (C.c_uint32 * 8)()
    for n in (0, 1, 5, 7, 10, 130, 131, 402, 803, 1 << 20):
        L.hs_safe_tag(n, out)
        assert val(out) == o.safe_tag(n) * R256 % Q


@pytest.mark.parametrize("vi,kind", [(0, "single"), (1, "double"), (2, "vargen")])
def test_typed_input_twin(L, vi, kind):
    """jjs_verify_ext stages on the CPU twin: typed items built from wire items give the wire result; typed-only
    failure modes (z = 0, inconsistent t1 t2, off-curve, unreduced limbs, projective identity) match the oracle."""
    from tests import typed_inputs as ti
    gen = {"single": co.gen_single, "double": co.gen_double, "vargen": co.gen_vargen}[kind]
    ver = {"single": co.verify_single, "double": co.verify_double, "vargen": co.verify_vargen}[kind]
    n = 150
    pk, sig, msg = gen(0x7E, n)
    pk, sig, msg, _, _ = adv.make_adversarial(kind, pk, sig, msg, seed=3, frac=0.5)
    st_wire, c_wire = ver(pk, sig, msg)
    pts, u, keep = ti.to_typed(vi, pk, sig, msg, seed=4)
    pts, u, msg_k, st_wire, c_wire = pts[keep], u[keep], msg[keep], st_wire[keep], c_wire[keep]
    m = int(keep.sum())
    st_o, c_o = co.verify_ext(vi, pts, u, msg_k)
    assert np.array_equal(st_o, st_wire) and np.array_equal(c_o, c_wire)     # oracle: typed == wire
    hst, hc = np.zeros(m, np.uint8), np.zeros((m, 32), np.uint8)
    L.hs_verify_ext(vi, _p(pts), _p(u), _p(msg_k), C.c_size_t(m), _p(hst), _p(hc))
    assert np.array_equal(hst, st_o) and np.array_equal(hc, c_o)
    bad, expected = ti.corrupt_typed(pts, st_o, seed=5)
    st_b, c_b = co.verify_ext(vi, bad, u, msg_k)
    L.hs_verify_ext(vi, _p(bad), _p(u), _p(msg_k), C.c_size_t(m), _p(hst), _p(hc))
    assert np.array_equal(hst, st_b) and np.array_equal(hc, c_b)
    for i, e in enumerate(expected):
        if e is not None and st_o[i] != 3:
            assert st_b[i] == e, (i, e, st_b[i])


def test_multisig_combine_twin(L):
    """multisig::combine stages on the CPU twin against the oracle: valid sessions, a bad share, a swapped commitment,
    an undecodable key, an empty session, a non-canonical share and a small-order commitment."""
    signers = [1, 2, 3, 4, 2, 3, 5, 0, 2, 3]
    pks, Rs, Ss, zs, off, msg = co.gen_multisig(5, signers)
    zs[off[2] + 1, 0] ^= 1
    Rs[off[4]] = Rs[off[4] + 1]
    pks[off[5]] = np.frombuffer(adv.off_curve_encoding(np.random.default_rng(1)), np.uint8)
    zs[off[8]] = 0xFF
    Ss[off[9]] = np.frombuffer(adv.torsion()[4], np.uint8)
    st, bad, sig, ok = co.multisig_combine(pks, Rs, Ss, zs, off, msg)
    n, K = len(signers), int(off[-1])
    hst, hbad, hsig, hok = np.zeros(n, np.uint8), np.zeros(n, np.uint32), np.zeros((n, 64), np.uint8), np.zeros(K, np.uint8)
    L.hs_multisig_combine(_p(pks), _p(Rs), _p(Ss), _p(zs), _p(off), _p(msg), C.c_size_t(n), _p(hok), _p(hst), _p(hbad), _p(hsig))
    assert np.array_equal(hst, st) and np.array_equal(hbad, bad) and np.array_equal(hsig, sig) and np.array_equal(hok, ok)
    assert st.tolist() == [0, 0, 5, 0, 5, 3, 0, 4, 3, 5]
