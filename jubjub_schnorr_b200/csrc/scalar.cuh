// Scalar-side helpers of the verify path: arithmetic modulo the subgroup order r (off the multiply-heavy path) and
// the half-size decomposition of the challenge used by the equation kernel.
#pragma once
#include "consts.cuh"
#include "fq.cuh"

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

// Half-size decomposition of a challenge (Antipa, Brown, Gallant, Lambert, Struik, Vanstone: "Accelerated
// verification of ECDSA signatures"): for c < r returns tau < 2^126 and rho != 0, |rho| < 2^126, with
//     tau == rho * c  (mod r)            (rho = rho_neg ? -rho_abs : rho_abs)
// by running the extended Euclidean algorithm on (r, c) until the remainder drops below 2^126.  Then, for points
// of the prime-order subgroup,  u*G + c*PK == R   <=>   (rho*u)*G + tau*PK - rho*R == O,  which needs 126-bit
// multipliers for the two variable bases.  Quotients are estimated from below in double precision (the margin
// 2^-40 dwarfs the 2^-49 conversion error) and capped at 2^31 - 1; the exact comparison a >= b drives the loop, so
// an underestimate only costs another pass.  Huge quotients are consumed 32 bits at a time.
//
// Consecutive cofactors of the Euclidean sequence are coprime, so when the stopping pair has an even rho the pair one
// step earlier (remainder a >= 2^126, cofactor ta < rho) has an odd one; it is returned instead when its remainder
// still fits the 33 signed radix-16 digits of the equation kernel (a < 2^130, the usual case: a * rho <= r).  An odd
// rho is what lets the equation kernel skip the subgroup test of R (verify_core.cuh, stage_equation): rho_odd reports
// whether one was found.  tau has 5 limbs (< 2^130), rho 4.
JJS_HD void half_gcd(uint32_t* tau5, uint32_t* rho4, bool& rho_neg, bool& rho_odd, const uint32_t* c8) {
    uint32_t a[8], b[8], ta[4] = {0, 0, 0, 0}, tb[4] = {1, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] = JJS_C(R_ORDER)[i]; b[i] = c8[i]; }
    bool neg = false;
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
    bool prev = !(tb[0] & 1u) && (ta[0] & 1u) && (a[7] | a[6] | a[5]) == 0u && a[4] < 4u;
#pragma unroll
    for (int i = 0; i < 4; i++) { tau5[i] = prev ? a[i] : b[i]; rho4[i] = prev ? ta[i] : tb[i]; }
    tau5[4] = prev ? a[4] : 0u;
    rho_neg = prev ? !neg : neg;
    rho_odd = (rho4[0] & 1u) != 0u;
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
