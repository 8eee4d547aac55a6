// Scalar-side helpers of the verify path: arithmetic modulo the subgroup order r (off the multiply-heavy path) and
// the half-size decomposition of the challenge used by the equation kernel.
#pragma once
#include "consts.cuh"
#include "fq.cuh"
#include <cmath>

namespace jjs {

// x (16 limbs, only the low `nbits` bits may be non-zero) mod r, bit-serial
JJS_HD void fr_reduce_wide(uint32_t* out, const uint32_t* x16, int nbits = 512) {
    uint32_t acc[8], ord[8], s[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { acc[i] = 0; ord[i] = JJS_C(R_ORDER)[i]; }
#pragma unroll 1
    for (int bit = nbits - 1; bit >= 0; bit--) {
        uint32_t in = (x16[bit >> 5] >> (bit & 31)) & 1u;
#pragma unroll
        for (int i = 7; i > 0; i--) acc[i] = (acc[i] << 1) | (acc[i - 1] >> 31);
        acc[0] = (acc[0] << 1) | in;
        uint32_t borrow = sub8(s, acc, ord);
#pragma unroll
        for (int i = 0; i < 8; i++) acc[i] = borrow ? acc[i] : s[i];
    }
#pragma unroll
    for (int i = 0; i < 8; i++) out[i] = acc[i];
}
JJS_HD void fr_mul(uint32_t* out, const uint32_t* a, const uint32_t* b, int nbits = 512) {
    uint32_t t[16];
    mul_wide(t, a, b);
    fr_reduce_wide(out, t, nbits);
}
// out = a * b mod r for a < 2^128 (its four low limbs are used) and b < r, by Barrett reduction: with
// x = a b < 2^380,  q = ((x >> 248) * floor(2^380 / r)) >> 132  is floor(x / r) or up to two less, so x - q r < 3 r
// and two conditional subtractions finish.  Three wide products on the multiply pipe replace the 384 shift-and-subtract
// steps of fr_reduce_wide (the equation kernel computes rho * u this way, once per equation, with every warp of the
// SM in that phase at the same time).
JJS_HD void fr_mul_short(uint32_t* out, const uint32_t* a4, const uint32_t* b8) {
    constexpr uint32_t MU[5] = {0x83f0476au, 0x1c5aee5bu, 0xff6f1bbeu, 0x1aa84a76u, 0x1u};  // floor(2^380 / r)
    uint32_t a[8], x[16], x1[8], mu[8], t[16], q[8], ord[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] = i < 4 ? a4[i] : 0u; mu[i] = i < 5 ? MU[i] : 0u; ord[i] = JJS_C(R_ORDER)[i]; }
    mul_wide(x, a, b8);
    // x1 = x >> 248 (248 = 7 * 32 + 24): five limbs
#pragma unroll
    for (int i = 0; i < 8; i++) x1[i] = i < 5 ? ((x[7 + i] >> 24) | (x[8 + i] << 8)) : 0u;
    mul_wide(t, x1, mu);
    // q = t >> 132 (132 = 4 * 32 + 4): four limbs
#pragma unroll
    for (int i = 0; i < 8; i++) q[i] = i < 4 ? ((t[4 + i] >> 4) | (t[5 + i] << 28)) : 0u;
    mul_wide(t, q, ord);
    // rem = x - q r over nine limbs (the true value is below 3 r < 2^254)
    uint32_t rem[8], s[8];
    sub8(rem, x, t);
#pragma unroll 1
    for (int k = 0; k < 2; k++) {
        uint32_t borrow = sub8(s, rem, ord);
#pragma unroll
        for (int i = 0; i < 8; i++) rem[i] = borrow ? rem[i] : s[i];
    }
#pragma unroll
    for (int i = 0; i < 8; i++) out[i] = rem[i];
}
// The same for a < 2^160 (five limbs): x = a b < 2^412,  q = ((x >> 248) * floor(2^412 / r)) >> 164.  Used where the second
// vector of the half-size decomposition feeds the three-scalar reduction (its rho can exceed 128 bits by a few).
JJS_HD void fr_mul_160(uint32_t* out, const uint32_t* a5, const uint32_t* b8) {
    constexpr uint32_t MU[6] = {0x01b7c721u, 0x83f0476au, 0x1c5aee5bu, 0xff6f1bbeu, 0x1aa84a76u, 0x1u};  // floor(2^412 / r)
    uint32_t a[8], x[16], x1[8], mu[8], t[16], q[8], ord[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] = i < 5 ? a5[i] : 0u; mu[i] = i < 6 ? MU[i] : 0u; ord[i] = JJS_C(R_ORDER)[i]; }
    mul_wide(x, a, b8);
    // x1 = x >> 248 (248 = 7 * 32 + 24): six limbs
#pragma unroll
    for (int i = 0; i < 8; i++) x1[i] = i < 6 ? ((x[7 + i] >> 24) | (x[8 + i] << 8)) : 0u;
    mul_wide(t, x1, mu);
    // q = t >> 164 (164 = 5 * 32 + 4): six limbs
#pragma unroll
    for (int i = 0; i < 8; i++) q[i] = i < 6 ? ((t[5 + i] >> 4) | (t[6 + i] << 28)) : 0u;
    mul_wide(t, q, ord);
    uint32_t rem[8], sv[8];
    sub8(rem, x, t);   // the true value is below 3 r < 2^254
#pragma unroll 1
    for (int k = 0; k < 2; k++) {
        uint32_t borrow = sub8(sv, rem, ord);
#pragma unroll
        for (int i = 0; i < 8; i++) rem[i] = borrow ? rem[i] : sv[i];
    }
#pragma unroll
    for (int i = 0; i < 8; i++) out[i] = rem[i];
}
JJS_HD void fr_sub(uint32_t* out, const uint32_t* a, const uint32_t* b) {  // a, b < r
    uint32_t d[8], ord[8], e[8];
#pragma unroll
    for (int i = 0; i < 8; i++) ord[i] = JJS_C(R_ORDER)[i];
    uint32_t borrow = sub8(d, a, b);
    add8(e, d, ord);
#pragma unroll
    for (int i = 0; i < 8; i++) out[i] = borrow ? e[i] : d[i];
}

JJS_HD double limbs_to_double(const uint32_t* a) {
    double d = 0.0;
#pragma unroll
    for (int i = 7; i >= 0; i--) d = d * 4294967296.0 + (double)a[i];
    return d;
}

// The classical form of the half-size decomposition (see half_gcd below for what it is for): the extended Euclidean algorithm on
// (r, c), run until the remainder drops below 2^126.  The kernels use the warp-uniform reduction of half_gcd_vectors; this
// sequence is its fallback (half_gcd_classical) for a run that does not finish within its round cap, and the reference the
// twin tests it against.  Quotients are estimated from below in double precision (the margin 2^-40 dwarfs the 2^-49
// conversion error) and capped at 2^31 - 1; the exact comparison a >= b drives the loop, so an underestimate only costs
// another pass.  Huge quotients are consumed 32 bits at a time.  On return (a, ta) and (b, tb) are consecutive (remainder,
// cofactor) pairs with b < 2^126 <= a,
//     b == (neg ? -tb : tb) * c   and   a == (neg ? ta : -ta) * c   (mod r),        ta <= tb < 2^126.
// Consecutive cofactors are coprime, so when the stopping pair has an even cofactor the previous pair has an odd one
// (half_gcd_classical takes it when its remainder still fits the 33 signed radix-16 digits of the equation kernel).
JJS_HD void half_gcd_core(uint32_t* a, uint32_t* b, uint32_t* ta, uint32_t* tb, bool& neg, const uint32_t* c8) {
#pragma unroll
    for (int i = 0; i < 4; i++) { ta[i] = 0; tb[i] = 0; }
    tb[0] = 1;
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] = JJS_C(R_ORDER)[i]; b[i] = c8[i]; }
    neg = false;
#pragma unroll 1
    while ((b[7] | b[6] | b[5] | b[4]) != 0u || b[3] >= (1u << 30)) {
        uint32_t tmp[8];
#pragma unroll 1
        while (sub8(tmp, a, b) == 0u) {  // a >= b
            double qd = limbs_to_double(a) / limbs_to_double(b) * (1.0 - 9.094947017729282e-13);
            // quotients of 2^32 and more (probability 2^-32 per step for a hash output, but they must not stall the
            // loop) are taken one 32-bit digit at a time: q * 2^(32 k) <= a / b, applied to limb-shifted copies
            uint32_t bs[8], ts[4];
#pragma unroll
            for (int i = 0; i < 8; i++) bs[i] = b[i];
#pragma unroll
            for (int i = 0; i < 4; i++) ts[i] = tb[i];
#pragma unroll 1
            while (qd >= 4294967296.0) {
                qd *= 2.3283064365386963e-10;
#pragma unroll
                for (int i = 7; i > 0; i--) bs[i] = bs[i - 1];
                bs[0] = 0;
#pragma unroll
                for (int i = 3; i > 0; i--) ts[i] = ts[i - 1];
                ts[0] = 0;
            }
            uint32_t q = (uint32_t)qd;
            if (q == 0u) q = 1u;
            uint64_t carry = 0;
            int64_t borrow = 0;
#pragma unroll
            for (int i = 0; i < 8; i++) {  // a -= q * bs
                uint64_t p = (uint64_t)q * bs[i] + carry;
                carry = p >> 32;
                int64_t dlt = (int64_t)a[i] - (int64_t)(uint32_t)p + borrow;
                a[i] = (uint32_t)dlt;
                borrow = dlt >> 32;
            }
            carry = 0;
#pragma unroll
            for (int i = 0; i < 4; i++) {  // ta += q * ts   (stays below 2^126)
                uint64_t p = (uint64_t)q * ts[i] + ta[i] + carry;
                ta[i] = (uint32_t)p;
                carry = p >> 32;
            }
        }
#pragma unroll
        for (int i = 0; i < 8; i++) { uint32_t x = a[i]; a[i] = b[i]; b[i] = x; }
#pragma unroll
        for (int i = 0; i < 4; i++) { uint32_t x = ta[i]; ta[i] = tb[i]; tb[i] = x; }
        neg = !neg;
    }
}
// ---- three short scalars for an equation with two variable bases and a variable generator ---------------------------------
// u*Gen + c*PK == R has no fixed base, so the half-size trick above leaves a full-size multiplier on Gen.  Multiplying the
// equation by any z != 0 (mod r) gives the equivalent check  x*Gen + y*PK - z*R == O  with  x == z u,  y == z c  (mod r);
// the triples (x, y, z) form a lattice of rank 3 and determinant r^2, whose short vectors have ~168-bit coordinates
// (r^(2/3)), and a 3-table Straus interleave over 43 signed radix-16 windows then needs 168 doublings instead of 252.
//
// lattice3_reduce finds such a vector.  Basis: the two reduced vectors (tau_i, rho_i) of the half-size decomposition of c
// (half_gcd_vectors, ~126 bits each), extended by x_i = rho_i u mod r, and (r, 0, 0).  Reduction: greedy in the style of Semaev's rank-3 algorithm, with every
// quotient estimated in double precision from 53-bit approximations of the coordinates and applied EXACTLY to the 256-bit
// integers.  Every update is an integer row operation, so the vectors stay in the lattice whatever the rounding errors
// do: floating point only steers, it cannot make the result wrong, only longer -- and a result that is too long for the
// 43 windows (or a run that hits the round cap) makes the caller fall back to the two-table evaluation.
struct s256 {  // signed, two's complement, little-endian limbs
    uint32_t l[8];
};
JJS_HD bool s256_is_neg(const s256& a) { return (a.l[7] >> 31) != 0u; }
JJS_HD void s256_abs(uint32_t* m, const s256& a) {
    uint32_t z[8], n[8];
#pragma unroll
    for (int i = 0; i < 8; i++) z[i] = 0;
    sub8(n, z, a.l);
    bool neg = s256_is_neg(a);
#pragma unroll
    for (int i = 0; i < 8; i++) m[i] = neg ? n[i] : a.l[i];
}
JJS_HD void s256_negate(s256& a) {
    uint32_t z[8], n[8];
#pragma unroll
    for (int i = 0; i < 8; i++) z[i] = 0;
    sub8(n, z, a.l);
#pragma unroll
    for (int i = 0; i < 8; i++) a.l[i] = n[i];
}
// value * 2^-130 as a double (truncation error below 2^-50 relative)
JJS_HD double s256_to_double(const s256& a) {
    uint32_t m[8];
    s256_abs(m, a);
    double d = limbs_to_double(m) * 7.34683969347875e-40;  // 2^-130
    return s256_is_neg(a) ? -d : d;
}
// v -= k * (w << 32 limb_shift)   (mod 2^256: exact whenever the true result fits, which the callers' norms guarantee)
JJS_HD void s256_submul(s256& v, const s256& w, int32_t k, int limb_shift) {
    uint32_t ws[8], p[8];
#pragma unroll
    for (int i = 0; i < 8; i++) ws[i] = w.l[i];
#pragma unroll 1
    for (int s = 0; s < limb_shift; s++) {
#pragma unroll
        for (int i = 7; i > 0; i--) ws[i] = ws[i - 1];
        ws[0] = 0;
    }
    uint32_t mag = k < 0 ? (uint32_t)(-(int64_t)k) : (uint32_t)k;
    uint64_t carry = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint64_t t = (uint64_t)mag * ws[i] + carry;
        p[i] = (uint32_t)t;
        carry = t >> 32;
    }
    uint32_t r[8];
    if (k < 0) add8(r, v.l, p);
    else sub8(r, v.l, p);
#pragma unroll
    for (int i = 0; i < 8; i++) v.l[i] = r[i];
}
// One vector of the working basis in floating point: coordinates (scaled by 2^-130), squared length, and its row of the
// integer transformation that leads from the exact basis of the current round to it (entries are exact small integers).
struct lrow {
    double d[3], t[3], n;
};
JJS_HD void lrow_norm(lrow& v) { v.n = v.d[0] * v.d[0] + v.d[1] * v.d[1] + v.d[2] * v.d[2]; }
JJS_HD double lrow_dot(const lrow& a, const lrow& b) { return a.d[0] * b.d[0] + a.d[1] * b.d[1] + a.d[2] * b.d[2]; }
JJS_HD void lrow_swap_if(lrow& a, lrow& b, bool sw) {
#pragma unroll
    for (int k = 0; k < 3; k++) {
        double x = a.d[k], y = b.d[k];
        a.d[k] = sw ? y : x;
        b.d[k] = sw ? x : y;
        x = a.t[k], y = b.t[k];
        a.t[k] = sw ? y : x;
        b.t[k] = sw ? x : y;
    }
    double x = a.n, y = b.n;
    a.n = sw ? y : x;
    b.n = sw ? x : y;
}
JJS_HD double lattice_max3(double a, double b, double c) {
    a = fabs(a), b = fabs(b), c = fabs(c);
    double m = a > b ? a : b;
    return m > c ? m : c;
}
// a quotient as a 31-bit digit at a limb shift: q ~ k 2^(32 sh)
JJS_HD void lattice_digit(int32_t& k, int& sh, double q) {
    double m = fabs(q), sc = 1.0;
    sh = 0;
#pragma unroll 1
    while (m >= 2147483648.0 && sh < 8) {
        m *= 2.3283064365386963e-10;
        sc *= 2.3283064365386963e-10;
        sh++;
    }
    k = (int32_t)rint(q * sc);
}
JJS_HD bool s256_fits_170(const s256& a) {
    uint32_t m[8];
    s256_abs(m, a);
    return (m[7] | m[6]) == 0u && (m[5] >> 10) == 0u;
}
// t0 o0 + t1 o1 + t2 o2 for exact small integers t (as doubles)
JJS_HD void lattice3_comb(s256& r, const double* t, const s256& o0, const s256& o1, const s256& o2) {
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = 0;
    s256_submul(r, o0, -(int32_t)t[0], 0);
    s256_submul(r, o1, -(int32_t)t[1], 0);
    s256_submul(r, o2, -(int32_t)t[2], 0);
}
constexpr int LATTICE3_MAX_ROUNDS = 16, LATTICE3_INNER = 10;
constexpr double LATTICE3_CAP = 4194304.0;   // 2^22: bound on the entries of a round's transformation (and on every quotient)
constexpr int LATTICE3_WINDOWS = 43;         // signed radix-16 digits of a magnitude below 2^170
#ifndef JJS_LAT_RCP
#define JJS_LAT_RCP 1
#endif
// 1 / x for a positive normal x to ~40 bits (the quotient estimates need ~25): the hardware's approximate reciprocal and one
// Newton step in place of the ~40-instruction IEEE division subroutine; the host twin divides
JJS_HD double lattice_rcp(double x) {
#if defined(__CUDA_ARCH__) && JJS_LAT_RCP
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y * (2.0 - x * y);
#else
    return 1.0 / x;
#endif
}
JJS_HD double lattice_clamp(double k) {
    k = k > LATTICE3_CAP - 1.0 ? LATTICE3_CAP - 1.0 : k;
    return k < 1.0 - LATTICE3_CAP ? 1.0 - LATTICE3_CAP : k;
}
// every lane of the warp that is still reducing keeps the others in the loop (they run empty rounds), so that the exact
// update at the end of a round is executed once per warp, not once per lane
JJS_HD bool lattice_any(bool p) {
#if defined(__CUDA_ARCH__)
    return __any_sync(__activemask(), p) != 0;
#else
    return p;
#endif
}
// ---- the half-size decomposition, warp-uniform ------------------------------------------------------------------------------
// The two shortest vectors (tau, rho) of the lattice tau == rho c (mod r), i.e. a Gauss-reduced basis, by the same scheme as
// lattice3_reduce below (Lehmer-style rounds on doubles with an exact update per round, control flow independent of the data,
// a warp vote per round), started from (r, 0) and (c, 1).  Nearest-integer quotients make this the centred Euclidean
// algorithm: ~51 steps for a 250-bit challenge, 6-7 rounds.  It replaces the classical sequence of half_gcd_core above in the
// kernels (0.7 ms per 2^20 equations there, with every lane of a warp waiting for the slowest); the shortest vector is below
// 2^127 in both coordinates, the second one below 2^130 in all but 1 % of the cases, and one of the two has an odd rho
// (their determinant is +-r, which is odd).
constexpr int HGCD_MAX_ROUNDS = 24, HGCD_INNER = 14;
JJS_HD void half_gcd_vectors(s256 (&E)[2][2], const uint32_t* c8) {   // rows (tau, rho), E[0] the shorter on return
#pragma unroll
    for (int i = 0; i < 8; i++) {
        E[0][0].l[i] = c8[i];
        E[0][1].l[i] = i == 0 ? 1u : 0u;
        E[1][0].l[i] = JJS_C(R_ORDER)[i];
        E[1][1].l[i] = 0u;
    }
    bool converged = false;
#pragma unroll 1
    for (int round = 0; round < HGCD_MAX_ROUNDS && lattice_any(!converged); round++) {
        double ad[2], bd[2], at[2], bt[2], an, bn;
        ad[0] = s256_to_double(E[0][0]); ad[1] = s256_to_double(E[0][1]);
        bd[0] = s256_to_double(E[1][0]); bd[1] = s256_to_double(E[1][1]);
        an = ad[0] * ad[0] + ad[1] * ad[1];
        bn = bd[0] * bd[0] + bd[1] * bd[1];
        {   // a quotient of 2^20 or more against the first vector (tiny challenges; updates the cap kept out of a round) is applied
            // exactly, 31 bits at a time.  E[0] is the shorter vector here except before the first round, where it is (c, 1).
            double ratio = (ad[0] * bd[0] + ad[1] * bd[1]) * lattice_rcp(an);
            bool big = fabs(ratio) >= 1048576.0 && an <= bn;
            if (lattice_any(big)) {
                int32_t k;
                int sh;
                lattice_digit(k, sh, big ? ratio : 0.0);
                s256_submul(E[1][0], E[0][0], k, sh);
                s256_submul(E[1][1], E[0][1], k, sh);
                bd[0] = s256_to_double(E[1][0]);
                bd[1] = s256_to_double(E[1][1]);
                bn = bd[0] * bd[0] + bd[1] * bd[1];
            }
        }
        at[0] = 1.0; at[1] = 0.0;
        bt[0] = 0.0; bt[1] = 1.0;
#pragma unroll 1
        for (int it = 0; it <= HGCD_INNER; it++) {
            bool sw = an > bn;   // A: the shorter
#pragma unroll
            for (int j = 0; j < 2; j++) {
                double x = ad[j], y = bd[j];
                ad[j] = sw ? y : x;
                bd[j] = sw ? x : y;
                x = at[j], y = bt[j];
                at[j] = sw ? y : x;
                bt[j] = sw ? x : y;
            }
            { double x = an, y = bn; an = sw ? y : x; bn = sw ? x : y; }
            if (it == HGCD_INNER) break;   // the last pass only orders the pair
            double ra = (ad[0] * bd[0] + ad[1] * bd[1]) * lattice_rcp(an);
            bool want = fabs(ra) > 0.500001;
            double k = want ? lattice_clamp(rint(ra)) : 0.0;
            double t0 = bt[0] - k * at[0], t1 = bt[1] - k * at[1];
            k = (fabs(t0) <= LATTICE3_CAP && fabs(t1) <= LATTICE3_CAP) ? k : 0.0;
            bt[0] -= k * at[0]; bt[1] -= k * at[1];
            bd[0] -= k * ad[0]; bd[1] -= k * ad[1];
            bn = bd[0] * bd[0] + bd[1] * bd[1];
            converged = converged || !want;
        }
#pragma unroll
        for (int k = 0; k < 2; k++) {
            s256 o0 = E[0][k], o1 = E[1][k], z;
#pragma unroll
            for (int i = 0; i < 8; i++) z.l[i] = 0;
            double ta3[3] = {at[0], at[1], 0.0}, tb3[3] = {bt[0], bt[1], 0.0};
            lattice3_comb(E[0][k], ta3, o0, o1, z);
            lattice3_comb(E[1][k], tb3, o0, o1, z);
        }
    }
}
// magnitude below 2^bits
JJS_HD bool s256_fits(const s256& a, int bits) {
    uint32_t m[8];
    s256_abs(m, a);
    bool ok = true;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        int lo = 32 * i;
        if (lo >= bits) ok = ok && m[i] == 0u;
        else if (lo + 32 > bits) ok = ok && (m[i] >> (bits - lo)) == 0u;
    }
    return ok;
}
// For c < r returns tau < 2^130 (5 limbs) and rho != 0, |rho| < 2^128 (4 limbs), with  tau == rho * c (mod r),
// rho = rho_neg ? -|rho| : |rho|  (Antipa, Brown, Gallant, Lambert, Struik, Vanstone: "Accelerated verification of ECDSA
// signatures").  Then, for points of the prime-order subgroup,  u*G + c*PK == R   <=>   (rho*u)*G + tau*PK - rho*R == O,  which
// needs ~126-bit multipliers for the two variable bases.  Of the two reduced vectors the shorter one is taken if its rho is odd,
// else the other one if it fits; an odd rho is what lets the equation kernel skip the subgroup test of R (verify_core.cuh,
// stage_equation): rho_odd reports whether one was found (98 % of the challenges).
JJS_HD void half_gcd_classical(uint32_t* tau5, uint32_t* rho4, bool& rho_neg, bool& rho_odd, const uint32_t* c8) {
    uint32_t a[8], b[8], ta[4], tb[4];
    bool neg;
    half_gcd_core(a, b, ta, tb, neg, c8);
    bool prev = !(tb[0] & 1u) && (ta[0] & 1u) && (a[7] | a[6] | a[5]) == 0u && a[4] < 4u;
#pragma unroll
    for (int i = 0; i < 4; i++) { tau5[i] = prev ? a[i] : b[i]; rho4[i] = prev ? ta[i] : tb[i]; }
    tau5[4] = prev ? a[4] : 0u;
    rho_neg = prev ? !neg : neg;
    rho_odd = (rho4[0] & 1u) != 0u;
}
JJS_HD void half_gcd(uint32_t* tau5, uint32_t* rho4, bool& rho_neg, bool& rho_odd, const uint32_t* c8) {
    s256 E[2][2];
    half_gcd_vectors(E, c8);
    if (!(s256_fits(E[0][0], 130) && s256_fits(E[0][1], 128))) {
        // the reduction did not finish within its round cap (not seen; the floating-point steering gives no guarantee): the
        // classical sequence, which always does
        half_gcd_classical(tau5, rho4, rho_neg, rho_odd, c8);
        return;
    }
    bool odd0 = (E[0][1].l[0] & 1u) != 0u;
    bool use1 = !odd0 && (E[1][1].l[0] & 1u) != 0u && s256_fits(E[1][0], 130) && s256_fits(E[1][1], 128);
    uint32_t t[8], rr[8];
    s256 T, Rr;
#pragma unroll
    for (int i = 0; i < 8; i++) { T.l[i] = use1 ? E[1][0].l[i] : E[0][0].l[i]; Rr.l[i] = use1 ? E[1][1].l[i] : E[0][1].l[i]; }
    s256_abs(t, T);
    s256_abs(rr, Rr);
#pragma unroll
    for (int i = 0; i < 5; i++) tau5[i] = t[i];
#pragma unroll
    for (int i = 0; i < 4; i++) rho4[i] = rr[i];
    rho_neg = s256_is_neg(T) != s256_is_neg(Rr);   // tau is returned as a magnitude: the sign moves to rho
    rho_odd = (rho4[0] & 1u) != 0u;
}

// takes the vector (x, y, z) if it fits the windows, has z != 0 and (when asked) an odd z
JJS_HD bool lattice3_pick(const s256* V, bool want_odd, uint32_t* xm, uint32_t* ym, uint32_t* zm, bool& xneg, bool& yneg, bool& zneg) {
    uint32_t nz = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) nz |= V[2].l[i];
    if (nz == 0u || (want_odd && !(V[2].l[0] & 1u))) return false;
    if (!(s256_fits_170(V[0]) && s256_fits_170(V[1]) && s256_fits_170(V[2]))) return false;
    s256_abs(xm, V[0]);
    s256_abs(ym, V[1]);
    s256_abs(zm, V[2]);
    xneg = s256_is_neg(V[0]);
    yneg = s256_is_neg(V[1]);
    zneg = s256_is_neg(V[2]);
    return true;
}
// On success: |x|, |y|, |z| < 2^170 as magnitudes (six limbs used) with their signs, z != 0, x == z u and y == z c (mod r);
// z_odd tells whether z is odd (preferred: see verify_core.cuh, stage_equation).  Returns false if no basis vector fits.
//
// Organisation: Lehmer style, and uniform across a warp.  The exact basis E is touched once per ROUND: a round lifts E to
// doubles, runs LATTICE3_INNER greedy iterations on the doubles alone while accumulating the integer transformation T (its
// entries, and so every quotient, capped at 2^22, which leaves the estimates > 8 good bits after the cancellation that
// goes with it; an update that would exceed the cap waits for the next round), and then replaces E by T E exactly.  One
// iteration = order the three vectors by length; a Gauss step of the second against the shortest; then the longest against
// the plane of the other two (2 x 2 Gram system) when that pair is well conditioned, else against the shortest alone.
// Reducing the longest in EVERY iteration keeps all quotients small -- left alone, (r, 0, 0) would need 64-bit
// coefficients once the other two have finished their Gauss phase.  Control flow does not depend on the data: a step that
// does not apply has quotient 0, so the 32 lanes of a warp stay together (a first version with data-dependent branches and
// per-lane rounds cost 8.3 ms per 2^20 equations; this one 1.5 ms, plus 0.7 ms for the half-gcd).  Taking both steps of an
// iteration from the same state (half the dependency chain, 20 % more iterations) measured no faster; neither did replacing
// the IEEE divisions.  Model: tools/lattice3_model.py (4-6 rounds per challenge).
#if !defined(__CUDA_ARCH__)
inline bool& lattice3_host_force_none() {   // host twin only: lets a test drive the callers' fallback (no vector found)
    static bool force = false;
    return force;
}
#endif
JJS_HD bool lattice3_reduce(uint32_t* xm, uint32_t* ym, uint32_t* zm, bool& xneg, bool& yneg, bool& zneg, bool& z_odd, const uint32_t* u8,
                            const uint32_t* c8) {
#if !defined(__CUDA_ARCH__)
    if (lattice3_host_force_none()) return false;
#endif
    s256 E[3][3];   // rows: basis vectors (x, y, z)
    {
        // (rho_i u mod r, tau_i, rho_i) for the two reduced vectors of the half-size decomposition, and (r, 0, 0)
        s256 H[2][2];
        half_gcd_vectors(H, c8);
        if (!s256_fits(H[0][1], 160) || !s256_fits(H[1][1], 160)) return false;   // degenerate challenges (c tiny): direct evaluation
#pragma unroll
        for (int v = 0; v < 2; v++) {
            uint32_t rm[8], m[8];
            s256_abs(rm, H[v][1]);
            fr_mul_160(m, rm, u8);
#pragma unroll
            for (int i = 0; i < 8; i++) { E[v][0].l[i] = m[i]; E[v][1].l[i] = H[v][0].l[i]; E[v][2].l[i] = H[v][1].l[i]; }
            if (s256_is_neg(H[v][1])) s256_negate(E[v][0]);
        }
#pragma unroll
        for (int i = 0; i < 8; i++) { E[2][0].l[i] = JJS_C(R_ORDER)[i]; E[2][1].l[i] = 0; E[2][2].l[i] = 0; }
    }
    bool converged = false;
#pragma unroll 1
    for (int round = 0; round < LATTICE3_MAX_ROUNDS && lattice_any(!converged); round++) {
        lrow A, B, C;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            A.d[k] = s256_to_double(E[0][k]);
            B.d[k] = s256_to_double(E[1][k]);
            C.d[k] = s256_to_double(E[2][k]);
            A.t[k] = k == 0 ? 1.0 : 0.0;
            B.t[k] = k == 1 ? 1.0 : 0.0;
            C.t[k] = k == 2 ? 1.0 : 0.0;
        }
        lrow_norm(A);
        lrow_norm(B);
        lrow_norm(C);
        {   // laggards: a vector that needs a quotient of 2^20 or more against the first one (an update the cap kept out of the
            // previous round while the others went on shrinking) gets it exactly, 31 bits at a time; rare (3 in 10^4 challenges)
            double ian = lattice_rcp(A.n), rb = lrow_dot(A, B) * ian, rc = lrow_dot(A, C) * ian;
            bool bigb = fabs(rb) >= 1048576.0, bigc = fabs(rc) >= 1048576.0;
            if (lattice_any(bigb || bigc)) {
                int32_t kb, kc;
                int shb, shc;
                lattice_digit(kb, shb, bigb ? rb : 0.0);
                lattice_digit(kc, shc, bigc ? rc : 0.0);
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    s256_submul(E[1][k], E[0][k], kb, shb);
                    s256_submul(E[2][k], E[0][k], kc, shc);
                    B.d[k] = s256_to_double(E[1][k]);
                    C.d[k] = s256_to_double(E[2][k]);
                }
                lrow_norm(B);
                lrow_norm(C);
            }
        }
#pragma unroll 1
        for (int it = 0; it < LATTICE3_INNER; it++) {
            lrow_swap_if(A, B, A.n > B.n);
            lrow_swap_if(B, C, B.n > C.n);
            lrow_swap_if(A, B, A.n > B.n);
            double ab = lrow_dot(A, B), ac = lrow_dot(A, C), bc = lrow_dot(B, C);
            double ian = lattice_rcp(A.n), ra = ab * ian;
            bool want1 = fabs(ra) > 0.500001;
            // Gauss step of B against A
            double k = want1 ? lattice_clamp(rint(ra)) : 0.0;
            {
                double t0 = B.t[0] - k * A.t[0], t1 = B.t[1] - k * A.t[1], t2 = B.t[2] - k * A.t[2];
                k = lattice_max3(t0, t1, t2) <= LATTICE3_CAP ? k : 0.0;
            }
#pragma unroll
            for (int j = 0; j < 3; j++) { B.t[j] -= k * A.t[j]; B.d[j] -= k * A.d[j]; }
            lrow_norm(B);
            // C against the plane of A and B, or against A alone while A and B are still close to parallel
            ab = lrow_dot(A, B);
            bc = lrow_dot(B, C);
            double mn = A.n < B.n ? A.n : B.n;
            bool well = fabs(ab) <= 0.55 * mn;
            double idet = lattice_rcp(well ? A.n * B.n - ab * ab : 1.0);
            double fa = well ? (ac * B.n - bc * ab) * idet : ac * ian, fb = well ? (bc * A.n - ac * ab) * idet : 0.0;
            double qa = lattice_clamp(rint(fa)), qb = lattice_clamp(rint(fb));
            bool want2 = qa != 0.0 || qb != 0.0;
            {
                double t0 = C.t[0] - qa * A.t[0] - qb * B.t[0], t1 = C.t[1] - qa * A.t[1] - qb * B.t[1], t2 = C.t[2] - qa * A.t[2] - qb * B.t[2];
                bool fits = lattice_max3(t0, t1, t2) <= LATTICE3_CAP;
                qa = fits ? qa : 0.0;
                qb = fits ? qb : 0.0;
            }
#pragma unroll
            for (int j = 0; j < 3; j++) { C.t[j] -= qa * A.t[j] + qb * B.t[j]; C.d[j] -= qa * A.d[j] + qb * B.d[j]; }
            lrow_norm(C);
            converged = converged || (!want1 && !want2);
        }
        lrow_swap_if(A, B, A.n > B.n);   // E stays ordered by length
        lrow_swap_if(B, C, B.n > C.n);
        lrow_swap_if(A, B, A.n > B.n);
#pragma unroll
        for (int k = 0; k < 3; k++) {
            s256 o0 = E[0][k], o1 = E[1][k], o2 = E[2][k];
            lattice3_comb(E[0][k], A.t, o0, o1, o2);
            lattice3_comb(E[1][k], B.t, o0, o1, o2);
            lattice3_comb(E[2][k], C.t, o0, o1, o2);
        }
    }
    // the shortest vector that fits, one with an odd z first (E is ordered by length)
    z_odd = true;
    if (lattice3_pick(E[0], true, xm, ym, zm, xneg, yneg, zneg) || lattice3_pick(E[1], true, xm, ym, zm, xneg, yneg, zneg) ||
        lattice3_pick(E[2], true, xm, ym, zm, xneg, yneg, zneg))
        return true;
    z_odd = false;
    return lattice3_pick(E[0], false, xm, ym, zm, xneg, yneg, zneg) || lattice3_pick(E[1], false, xm, ym, zm, xneg, yneg, zneg) ||
           lattice3_pick(E[2], false, xm, ym, zm, xneg, yneg, zneg);
}

// signed radix-16 digits d_i in [-8, 8) of a little-endian scalar of NLIMBS limbs; writes 8 * NLIMBS + 1 digits
// (the last one is the final carry, 0 or 1).  For scalars below 2^(32 NLIMBS - 3) the last digit is 0.
template <int NLIMBS>
JJS_HD void recode_signed16_n(int8_t* digits, const uint32_t* k, bool negate) {
    uint32_t carry = 0;
#pragma unroll 1
    for (int i = 0; i < 8 * NLIMBS; i++) {
        int d = (int)((k[i >> 3] >> ((i & 7) * 4)) & 15u) + (int)carry;
        carry = d >= 8;
        d -= (int)(carry << 4);
        digits[i] = (int8_t)(negate ? -d : d);
    }
    digits[8 * NLIMBS] = (int8_t)(negate ? -(int)carry : (int)carry);
}

// 33 signed radix-16 digits of a 5-limb scalar below 2^130 (digit 32 takes bits 128..129 and the carry: at most 4)
JJS_HD void recode_signed16_33(int8_t* digits, const uint32_t* k5, bool negate) {
    recode_signed16_n<4>(digits, k5, negate);
    int top = (int)(k5[4] & 3u);
    digits[32] = (int8_t)(negate ? digits[32] - top : digits[32] + top);
}

}  // namespace jjs
