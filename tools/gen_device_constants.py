#!/usr/bin/env python3
"""Generate jubjub_schnorr_b200/csrc/jjs_constants.h: every number the CUDA path needs, as 32-bit limbs.

    python tools/gen_device_constants.py

Derived from first principles with plain Python integers and hashlib; it does NOT
import oracle/.  tests/test_constants.py cross-checks the emitted values against the pinned oracle.
Contents: curve constants, subgroup order and its signed radix-16 digits, square-root machinery for
q - 1 = 2^32 t (fixed-exponent schedule, root-of-unity tables, perfect hash for the 2^8 discrete log),
the optimised Hades schedule (scaled round constants, S-box scale fixes), SAFE tags.
"""
import hashlib
import os
import struct

Q = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
R_ORDER = 0x0E7DB4EA6533AFA906673B0101343B00A6682093CCC81082D0970E5ED6F72CB7
MONT = 1 << 256
D = (-10240 * pow(10241, -1, Q)) % Q
G = (0x3FD2814C43AC65A6F1FBF02D0FD6CCE62E3EBB21FD6C54ED4DF7B7FFEC7BEACA, 0x12)
G_NUMS = (0x5E67B8F316F414F7BD9514C773FD4456931E316A39FE4541921710179DF76377,
          0x43D80EB3B2F3EB1B7B162DBEEB3B34FD9949BA0F82A5507A6705B707162E3EF8)
OUT = os.path.join(os.path.dirname(__file__), "..", "jubjub_schnorr_b200", "csrc", "jjs_constants.h")


def limbs(x, n=8):
    return "{" + ", ".join("0x%08xu" % ((x >> (32 * i)) & 0xFFFFFFFF) for i in range(n)) + "}"


def mont(x):
    return x * MONT % Q


# ------------------------------------------------------------------------------------------------
# square root machinery
# ------------------------------------------------------------------------------------------------
def window_schedule(e):
    """Sliding-window (width 4, odd powers 1..15) schedule for the fixed exponent e, MSB first: list of
    (squarings, odd power index); a trailing (squarings, 0xff) finishes with no multiply.  Self-checked."""
    bits = bin(e)[2:]
    sched, i, pending_sq, first = [], 0, 0, True
    while i < len(bits):
        if bits[i] == "0":
            pending_sq += 1
            i += 1
            continue
        j = min(i + 4, len(bits))
        while bits[j - 1] == "0":
            j -= 1
        val = int(bits[i:j], 2)
        sched.append((0 if first else pending_sq + (j - i), (val - 1) // 2))
        first = False
        pending_sq = 0
        i = j
    if pending_sq:
        sched.append((pending_sq, 0xFF))
    x = 3
    odd = [pow(x, 2 * k + 1, Q) for k in range(8)]
    acc = None
    for sq, idx in sched:
        if acc is None:
            acc = odd[idx]
            continue
        for _ in range(sq):
            acc = acc * acc % Q
        if idx != 0xFF:
            acc = acc * odd[idx] % Q
    assert acc == pow(x, e, Q)
    return sched


def sqrt_constants():
    t = (Q - 1) >> 32
    z = 2
    while pow(z, (Q - 1) // 2, Q) == 1:
        z += 1
    g = pow(z, t, Q)  # generator of the 2^32 torsion
    assert pow(g, 1 << 31, Q) == Q - 1
    e = (t - 1) // 2
    sched = window_schedule(e)
    tabs = {}
    for name, shift in (("T0", 0), ("T1", 8), ("T2", 16), ("H1", 7), ("H2", 15), ("H3", 23)):
        base = pow(g, -(1 << shift), Q)
        tabs[name] = [pow(base, j, Q) for j in range(256)]
    g3 = pow(g, 1 << 24, Q)
    mu256 = [pow(g3, j, Q) for j in range(256)]
    # perfect hash on the low Montgomery limb: slot = (limb0 * mult) >> (32 - bits)
    keys = [mont(v) & 0xFFFFFFFF for v in mu256]
    nbits = 13   # 256 keys into 8192 slots: a collision-free odd multiplier turns up within a few hundred tries
    mult = 1
    while mult < (1 << 20):
        slots = {}
        ok = True
        for j, kk in enumerate(keys):
            s = ((kk * mult) & 0xFFFFFFFF) >> (32 - nbits)
            if s in slots:
                ok = False
                break
            slots[s] = j
        if ok:
            break
        mult += 2
    assert ok, "no perfect hash multiplier found"
    table = [0] * (1 << nbits)
    for s, j in slots.items():
        table[s] = j
    return dict(t=t, g=g, e=e, sched=sched, tabs=tabs, hash_mult=mult, hash_bits=nbits, hash_table=table, mu256=mu256)


# ------------------------------------------------------------------------------------------------
# Poseidon / Hades, reference form and the scaled small-integer-MDS form
# ------------------------------------------------------------------------------------------------
WIDTH, FULL, PARTIAL = 5, 8, 60
LCM = 360360  # lcm(5..13): LCM / (i + j + 5) are the integer MDS entries


def round_constants():
    out, b, c = [], b"poseidon-for-plonk", 1
    for _ in range((FULL + PARTIAL) * WIDTH):
        b = hashlib.sha512(b).digest()
        c = (int.from_bytes(b, "little") % Q + c) % Q
        out.append(c * MONT % Q)
    return out


def mds_matrix():
    return [[MONT * pow(i + j + 5, -1, Q) % Q for j in range(WIDTH)] for i in range(WIDTH)]


def hades_reference(state):
    s, rc, M = list(state), iter(round_constants()), mds_matrix()
    for rnd in range(FULL + PARTIAL):
        s = [(x + next(rc)) % Q for x in s]
        if rnd < 4 or rnd >= 64:
            s = [pow(x, 5, Q) for x in s]
        else:
            s[4] = pow(s[4], 5, Q)
        s = [sum(M[i][k] * s[k] for k in range(WIDTH)) % Q for i in range(WIDTH)]
    return s


def safe_tag(n):
    data = struct.pack(">II", 0x80000000 | n, 1) + struct.pack(">Q", 0)
    return int.from_bytes(hashlib.blake2b(data, digest_size=64).digest(), "little") % Q


def hades_schedule():
    """Scale bookkeeping for the device permutation.

    The device keeps every lane as the integer  X = lam * x  (mod q), lam a per-round constant.
      * S-box via Montgomery products: X^5 / R^4 = lam^5 x^5 / R^4.
        full rounds: all lanes, so the common scale just becomes lam^5 / R^4.
        partial rounds: lane 4 only, followed by one Montgomery product with fix = R^5 / lam^4 so that
        the lane returns to scale lam.
      * MDS with the true matrix M[i][k] = R / (i+k+5) = (R / LCM) * A[i][k], A integer: the device
        computes  sum_k A[i][k] X_k + ark  and one 2^-32 Montgomery step, i.e. scale *= LCM' with
        LCM' = 2^-32 (mod q) relative to the integer combination.  The missing factor R/LCM of the true
        matrix is folded into lam as well.
      * the next round's constants are pre-scaled and folded into the MDS accumulator (before the 2^-32
        step, hence multiplied by 2^32).
    True state after the linear layer: y = (R/LCM) * A x.  Device: Y = 2^-32 * A X = 2^-32 lam' (A x)
      = [2^-32 lam' LCM / R] * y   ->  lam_next = lam' * LCM / (R * 2^32).
    """
    rc = round_constants()
    inv = lambda v: pow(v, -1, Q)
    two32 = 1 << 32
    lam = MONT % Q            # inputs enter in ordinary Montgomery form: X = R x
    first_ark = [lam * rc[i] % Q for i in range(WIDTH)]          # added as-is to the state before round 0
    folded_ark, fixes, lams = [], [], [lam]
    for rnd in range(FULL + PARTIAL):
        full = rnd < 4 or rnd >= 64
        if full:
            lam_s = pow(lam, 5, Q) * inv(pow(MONT, 4, Q)) % Q
        else:
            fixes.append(pow(MONT, 5, Q) * inv(pow(lam, 4, Q)) % Q)
            lam_s = lam
        lam_next = lam_s * LCM % Q * inv(MONT * two32 % Q) % Q
        if rnd + 1 < FULL + PARTIAL:
            nxt = [lam_next * rc[(rnd + 1) * WIDTH + i] % Q for i in range(WIDTH)]
            folded_ark.append([v * two32 % Q for v in nxt])          # sits in the accumulator before the 2^-32 step
        else:
            folded_ark.append([0] * WIDTH)
        lam = lam_next
        lams.append(lam)
    unscale = MONT * inv(lam) % Q        # Montgomery-multiply by (R^2 / lam) turns lam*x into R*x
    unscale_mont = unscale * MONT % Q
    return dict(first_ark=first_ark, folded_ark=folded_ark, fixes=fixes, lam_final=lam, unscale_mont=unscale_mont, lams=lams)


def hades_scaled_model(state, sch):
    """Integer model of exactly what the device does; must equal hades_reference."""
    A = [[LCM // (i + k + 5) for k in range(WIDTH)] for i in range(WIDTH)]
    Rinv = pow(MONT, -1, Q)
    mm = lambda a, b: a * b * Rinv % Q
    X = [(MONT * x + a) % Q for x, a in zip(state, sch["first_ark"])]
    p = 0
    for rnd in range(FULL + PARTIAL):
        full = rnd < 4 or rnd >= 64

        def sbox(v):
            v2 = mm(v, v)
            v4 = mm(v2, v2)
            return mm(v4, v)
        if full:
            X = [sbox(v) for v in X]
        else:
            X[4] = mm(sbox(X[4]), sch["fixes"][p])
            p += 1
        Y = []
        for i in range(WIDTH):
            acc = sch["folded_ark"][rnd][i] + sum(A[i][k] * X[k] for k in range(WIDTH))
            Y.append(acc * pow(1 << 32, -1, Q) % Q)
        X = Y
    return [mm(v, sch["unscale_mont"]) * Rinv % Q for v in X]   # back to canonical integers


# ------------------------------------------------------------------------------------------------
# subgroup membership by the order-8 Tate pairing (cofactor 8, cyclic 2-Sylow subgroup)
# ------------------------------------------------------------------------------------------------
def ed_add(p, q_):
    u1, v1 = p
    u2, v2 = q_
    k = D * u1 * u2 % Q * v1 * v2 % Q
    return ((u1 * v2 + v1 * u2) * pow(1 + k, -1, Q) % Q, (v1 * v2 + u1 * u2) * pow(1 - k, -1, Q) % Q)


def ed_mul(p, k):
    acc = (0, 1)
    for bit in bin(k)[2:]:
        acc = ed_add(acc, acc)
        if bit == "1":
            acc = ed_add(acc, p)
    return acc


def tonelli(a):
    if pow(a, (Q - 1) // 2, Q) != 1:
        return None
    s, t = 32, (Q - 1) >> 32
    z = 2
    while pow(z, (Q - 1) // 2, Q) == 1:
        z += 1
    c, x, b, m = pow(z, t, Q), pow(a, (t + 1) // 2, Q), pow(a, t, Q), s
    while b != 1:
        i, b2 = 0, b
        while b2 != 1:
            b2 = b2 * b2 % Q
            i += 1
        g = pow(c, 1 << (m - i - 1), Q)
        x, c = x * g % Q, g * g % Q
        b, m = b * c % Q, i
    return x


def tate_constants():
    """E(Fq) is cyclic of order 8 r, so P is in the prime-order subgroup iff P is in 8 E(Fq) iff the order-8 Tate
    pairing tau(T8, P) is an 8th power, T8 a point of exact order 8.  With the normalised Miller function
    f = l_T^4 l_2T^2 / (v_2T^4 x) on the Weierstrass model y^2 = x^3 + A B x^2 + B^2 x (x = B (1+v)/(1-v),
    y = B^2 (1+v)/((1-v) u)), cleared of denominators modulo 8th powers:
        g = (N_T V)^4 (N_2T u)^2 (B (1 - v^2))^7,   N_S = B^2 (1+v) - c_S (1-v) u - lam_S B (1+v) u,   V = B (1+v) - x_2T (1-v)
    and P is torsion free iff g^((q-1)/8) == 1.  Returns the constants and a checker used for self-test."""
    inv = lambda x: pow(x, -1, Q)
    a = Q - 1
    A, B = 2 * (a + D) * inv(a - D) % Q, 4 * inv(a - D) % Q
    a2, a4 = A * B % Q, B * B % Q
    v = 2
    while True:  # first curve point whose r-multiple has exact order 8
        v += 1
        u2 = (v * v - 1) * inv(1 + D * v * v) % Q
        u = tonelli(u2)
        if u is None:
            continue
        t8 = ed_mul((u, v), R_ORDER)
        if ed_mul(t8, 4) != (0, 1) and ed_mul(t8, 8) == (0, 1):
            break

    def to_w(p):
        pu, pv = p
        X = (1 + pv) * inv(1 - pv) % Q
        Y = (1 + pv) * inv((1 - pv) * pu) % Q
        return (B * X % Q, B * B % Q * Y % Q)

    def w_dbl(P):
        x1, y1 = P
        lam = (3 * x1 * x1 + 2 * a2 * x1 + a4) * inv(2 * y1) % Q
        x3 = (lam * lam - a2 - 2 * x1) % Q
        return (x3, (lam * (x1 - x3) - y1) % Q), lam

    Tw = to_w(t8)
    T2w, lamT = w_dbl(Tw)
    T4w, lam2T = w_dbl(T2w)
    assert T4w == (0, 0)
    cT, c2T = (Tw[1] - lamT * Tw[0]) % Q, (T2w[1] - lam2T * T2w[0]) % Q
    consts = dict(B=B, B2=B * B % Q, cT=cT, lamTB=lamT * B % Q, c2T=c2T, lam2TB=lam2T * B % Q, x2T=T2w[0])

    def torsion_free(p):
        pu, pv = p
        opv, omv = (1 + pv) % Q, (1 - pv) % Q
        nT = (consts["B2"] * opv - cT * omv % Q * pu - consts["lamTB"] * opv % Q * pu) % Q
        n2T = (consts["B2"] * opv - c2T * omv % Q * pu - consts["lam2TB"] * opv % Q * pu) % Q
        V = (B * opv - consts["x2T"] * omv) % Q
        h = B * (1 - pv * pv) % Q
        g = pow(nT * V, 4, Q) * pow(n2T * pu, 2, Q) % Q * pow(h, 7, Q) % Q
        return pow(g, (Q - 1) // 8, Q) == 1

    # self-test on every coset of the 8-torsion
    import random
    rnd = random.Random(7)
    for _ in range(6):
        p = ed_mul(G, rnd.randrange(1, R_ORDER))
        cur = p
        for j in range(8):
            assert torsion_free(cur) == (j == 0), "Tate subgroup test disagrees with the definition"
            cur = ed_add(cur, t8)
    cur = (0, 1)
    for j in range(8):  # the torsion points themselves (identity reported as not torsion free: it is rejected anyway)
        assert not torsion_free(cur)
        cur = ed_add(cur, t8)
    return consts


def signed_radix16(k, n=64):
    ds, carry = [], 0
    for i in range(n):
        d = ((k >> (4 * i)) & 15) + carry
        if d >= 8:
            d, carry = d - 16, 1
        else:
            carry = 0
        ds.append(d)
    assert carry == 0 and sum(d * 16 ** i for i, d in enumerate(ds)) == k
    return ds


def main():
    sq = sqrt_constants()
    sch = hades_schedule()
    tate = tate_constants()
    for st in ([1, 2, 3, 4, 5], [0] * 5, [Q - 1, 7, Q - 2, 12345678901234567890, 1 << 200]):
        assert hades_scaled_model(st, sch) == hades_reference(st), "scaled Hades model disagrees with the reference form"
    U, T = [], []
    u, t = U.append, T.append
    for out in (u, t):
        out("/* GENERATED by tools/gen_device_constants.py -- do not edit. */")
        out("#include <stdint.h>")
    u("/* Small, uniformly indexed constants.  Include inside a namespace with JJS_CONST_QUAL defined as")
    u("   `static const` (host) or `__constant__ const` (device constant bank). */")
    u("JJS_CONST_QUAL uint32_t R_ORDER[8] = %s;" % limbs(R_ORDER))
    u("JJS_CONST_QUAL int8_t R_ORDER_DIGITS[64] = {%s}; /* signed radix-16, little-endian */" % ", ".join(map(str, signed_radix16(R_ORDER))))
    u("JJS_CONST_QUAL uint32_t EDWARDS_D[8] = %s; /* Montgomery */" % limbs(mont(D)))
    u("JJS_CONST_QUAL uint32_t EDWARDS_2D[8] = %s; /* Montgomery */" % limbs(mont(2 * D % Q)))
    u("JJS_CONST_QUAL uint32_t GEN_UV[2][8] = {%s, %s}; /* GENERATOR_EXTENDED affine, Montgomery */" % (limbs(mont(G[0])), limbs(mont(G[1]))))
    u("JJS_CONST_QUAL uint32_t GEN_NUMS_UV[2][8] = {%s, %s}; /* GENERATOR_NUMS_EXTENDED */" % (limbs(mont(G_NUMS[0])), limbs(mont(G_NUMS[1]))))
    u("JJS_CONST_QUAL uint32_t Q_MINUS_2[8] = %s;" % limbs(Q - 2))
    u("/* order-8 Tate pairing subgroup test (see tate_constants in the generator), Montgomery: B, B^2, c_T, lam_T B, c_2T, lam_2T B, x_2T */")
    u("JJS_CONST_QUAL uint32_t TATE[7][8] = {%s};" % ", ".join(limbs(mont(tate[k])) for k in ("B", "B2", "cT", "lamTB", "c2T", "lam2TB", "x2T")))
    u("#define JJS_SQRT_SCHED_LEN %d" % len(sq["sched"]))
    u("JJS_CONST_QUAL uint8_t SQRT_SCHED[JJS_SQRT_SCHED_LEN][2] = {%s}; /* (squarings, odd-power index | 0xff) for a^((t-1)/2) */"
      % ", ".join("{%d, %d}" % s_ for s_ in sq["sched"]))
    u("#define JJS_DLOG_HASH_MULT 0x%08xu" % sq["hash_mult"])
    u("#define JJS_DLOG_HASH_BITS %d" % sq["hash_bits"])
    u("#define JJS_MDS_LCM %du" % LCM)
    u("JJS_CONST_QUAL uint32_t HADES_FIRST_ARK[5][8] = {%s};" % ", ".join(limbs(v) for v in sch["first_ark"]))
    # Emitted as ark + q 2^32 (nine limbs: the low eight here, the ninth in HADES_FOLDED_ARK_TOP at the end of the bank): the
    # single Montgomery step after a linear layer (csrc/fq.cuh, redc_one) then lands in (0, q + 2^244) without any correction.
    u("JJS_CONST_QUAL uint32_t HADES_FOLDED_ARK[68][5][8] = {")
    for row in sch["folded_ark"]:
        u(" {" + ", ".join(limbs(v + (Q << 32)) for v in row) + "},")
    u("};")
    u("JJS_CONST_QUAL uint32_t HADES_SBOX_FIX[60][8] = {")
    for v in sch["fixes"]:
        u(" %s," % limbs(v))
    u("};")
    u("JJS_CONST_QUAL uint32_t HADES_UNSCALE[8] = %s; /* Montgomery-multiply the scaled state by this to get R*x */" % limbs(sch["unscale_mont"]))
    u("#define JJS_MAX_ABSORB 130")
    u("JJS_CONST_QUAL uint32_t SAFE_TAG[JJS_MAX_ABSORB + 1][8] = { /* Montgomery; index = absorbed elements */")
    u(" %s," % limbs(0))
    for n in range(1, 131):
        u(" %s," % limbs(mont(safe_tag(n))))
    u("};")
    u("JJS_CONST_QUAL uint32_t DOUBLE_DOMAIN[8] = %s; /* BlsScalar::from(0x4a4a53434844424c), Montgomery */" % limbs(mont(0x4A4A53434844424C)))
    # appended last: the constant bank is laid out in declaration order, and the arrays above keep their offsets
    inv_sched = window_schedule(Q - 2)
    u("#define JJS_INV_SCHED_LEN %d" % len(inv_sched))
    u("JJS_CONST_QUAL uint8_t INV_SCHED[JJS_INV_SCHED_LEN][2] = {%s}; /* same encoding as SQRT_SCHED, for a^(q-2) */" % ", ".join("{%d, %d}" % s_ for s_ in inv_sched))
    u("/* the same limbs as doubles 2^52 + limb: the FP64 linear layer starts its exact column sums from them */")
    u("JJS_CONST_QUAL double HADES_FOLDED_ARK_D[68][5][8] = {")
    for row in sch["folded_ark"]:
        u(" {" + ", ".join("{" + ", ".join("%d.0" % ((1 << 52) + (((v + (Q << 32)) >> (32 * i)) & 0xFFFFFFFF)) for i in range(8)) + "}" for v in row) + "},")
    u("};")
    u("JJS_CONST_QUAL uint32_t HADES_FOLDED_ARK_TOP[68][5] = { /* limb 8 of ark + q 2^32 */")
    for row in sch["folded_ark"]:
        u(" {" + ", ".join("0x%08xu" % ((v + (Q << 32)) >> 256) for v in row) + "},")
    u("};")
    t("/* Large, data-dependently indexed tables: host arrays, uploaded to device global memory at context creation. */")
    t("static const uint8_t DLOG_HASH[1 << %d] = {%s};" % (sq["hash_bits"], ", ".join(map(str, sq["hash_table"]))))
    t("/* root-of-unity tables, Montgomery: T0,T1,T2 = g^(-j 2^(8i)); H1,H2,H3 = g^(-j 2^(8i-1)); order T0,T1,T2,H1,H2,H3 */")
    t("static const uint32_t ROOT_TABLES[6][256][8] = {")
    for name in ("T0", "T1", "T2", "H1", "H2", "H3"):
        t(" {" + ",\n  ".join(limbs(v_) for v_ in map(mont, sq["tabs"][name])) + "},")
    t("};")
    base = os.path.dirname(OUT)
    for name, lines in (("jjs_constants_uniform.h", U), ("jjs_constants_tables.h", T)):
        with open(os.path.join(base, name), "w") as f:
            f.write("\n".join(lines) + "\n")
    print("wrote constants: sched", len(sq["sched"]), "hash mult", hex(sq["hash_mult"]))


if __name__ == "__main__":
    main()
