// Fq multiplication on the FP64 pipe (sm_100a).  B200 issues DFMA beside IMAD.WIDE (profiles/r01_microbench_fp64.json:
// ~1.35 cycles per warp instruction against ~4.2, the two overlapping almost completely), and the integer multiply
// pipe is what bounds every kernel of the verify path.  fq_mul_fp computes exactly what fq_mul computes -- a * b / 2^256
// mod q, fully reduced, same 8 x 32-bit limb interface -- without a single integer multiply, so a caller can send part
// of its multiplications down the otherwise idle pipe.
//
// Method (Emmart, Zheng, Weems: "Faster modular exponentiation using double precision floating point arithmetic on
// the GPU", ARITH 2018): operands as five 52-bit limbs held in doubles; a limb product p < 2^104 is split exactly by
//     hi = fma_rz(x, y, 2^104)              = 2^104 + floor(p / 2^52) 2^52     (ulp of [2^104, 2^105) is 2^52)
//     lo = fma_rz(x, y, (2^104 + 2^52) - hi) = 2^52 + (p mod 2^52)             (exact, in [2^52, 2^53))
// so the mantissa fields of hi and lo ARE the two 52-bit halves; their raw bit patterns are summed into 64-bit column
// accumulators that start at minus the sum of the exponent fields they will receive.  Montgomery reduction runs in
// five 52-bit steps (radix 2^260) with quotient digit m = -t (1 + 2^32) mod 2^52, because q = 1 - 2^32 (mod 2^52).
// One operand is taken as 16 b, which turns the radix-2^260 result a (16 b) / 2^260 into a b / 2^256.
#pragma once
#include <stdint.h>
#include <string.h>

#include "../jubjub_schnorr_b200/csrc/fq.cuh"
#if !defined(__CUDA_ARCH__)
#include <cfenv>
#include <cmath>
#endif

namespace jjs {

JJS_HD double fp_from_bits(uint64_t b) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)b);
#else
    double d;
    memcpy(&d, &b, 8);
    return d;
#endif
}
JJS_HD uint64_t fp_to_bits(double d) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(d);
#else
    uint64_t b;
    memcpy(&b, &d, 8);
    return b;
#endif
}
JJS_HD double fp_fma_rz(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
    return __fma_rz(a, b, c);
#else
    const int old = fegetround();
    fesetround(FE_TOWARDZERO);
    volatile double va = a, vb = b, vc = c;
    volatile double r = std::fma(va, vb, vc);
    fesetround(old);
    return r;
#endif
}

constexpr uint64_t FP_MASK52 = (uint64_t(1) << 52) - 1;
constexpr uint64_t FP_BITS_2P52 = 0x4330000000000000ull;   // bit pattern of 2^52
constexpr uint64_t FP_BITS_2P104 = 0x4670000000000000ull;  // bit pattern of 2^104

// limb i (52 bits) of (x << SHIFT), x = eight 32-bit words, SHIFT in {0, 4}; returned as a double
template <int SHIFT>
JJS_HD double fp_limb52(const uint32_t* w, int i) {
    const int o = 52 * i - SHIFT;  // first bit of the limb within x
    uint64_t v;
    if (o < 0) {
        v = (((uint64_t)w[1] << 32) | w[0]) << (-o);
    } else {
        const int k = o >> 5, r = o & 31;
        const uint64_t w0 = w[k], w1 = k + 1 < 8 ? w[k + 1] : 0u, w2 = k + 2 < 8 ? w[k + 2] : 0u;
        v = ((w1 << 32) | w0) >> r;
        if (r > 12) v |= w2 << (64 - r);
    }
    return fp_from_bits((v & FP_MASK52) | FP_BITS_2P52) - 4503599627370496.0;
}

// number of lo / hi halves each column receives: product terms (i, j) -> lo to i + j, hi to i + j + 1; reduction step k
// sends the hi of m q_0 to k + 1 and lo / hi of m q_j (j = 1..4) to k + j / k + j + 1
JJS_HD constexpr uint64_t fp_column_bias(int c) {
    int nlo = 0, nhi = 0;
    for (int i = 0; i < 5; i++)
        for (int j = 0; j < 5; j++) {
            if (i + j == c) nlo++;
            if (i + j + 1 == c) nhi++;
        }
    for (int k = 0; k < 5; k++)
        for (int j = 0; j < 5; j++) {
            if (j >= 1 && k + j == c) nlo++;
            if (k + j + 1 == c) nhi++;
        }
    return (uint64_t)0 - ((uint64_t)nlo * FP_BITS_2P52 + (uint64_t)nhi * FP_BITS_2P104);
}

JJS_HD void fq_mul_fp_inl(fq& r, const fq& a, const fq& b) {
    const double C1 = 20282409603651670423947251286016.0;                  // 2^104
    const double C2 = 20282409603651670423947251286016.0 + 4503599627370496.0;  // 2^104 + 2^52
    // q in 52-bit limbs
    const double Q[5] = {4503595332403201.0, 52776117727231.0, 2711223964777892.0, 2203984808738944.0, 127464551688605.0};
    double A[5], B[5];
#pragma unroll
    for (int i = 0; i < 5; i++) {
        A[i] = fp_limb52<0>(a.l, i);
        B[i] = fp_limb52<4>(b.l, i);
    }
    uint64_t c[10];
#pragma unroll
    for (int k = 0; k < 10; k++) c[k] = fp_column_bias(k);
#pragma unroll
    for (int i = 0; i < 5; i++)
#pragma unroll
        for (int j = 0; j < 5; j++) {
            const double hi = fp_fma_rz(A[i], B[j], C1);
            const double lo = fp_fma_rz(A[i], B[j], C2 - hi);
            c[i + j] += fp_to_bits(lo);
            c[i + j + 1] += fp_to_bits(hi);
        }
#pragma unroll
    for (int k = 0; k < 5; k++) {
        const uint64_t t = c[k];
        const uint64_t low = t & FP_MASK52;
        const uint64_t carry = (t >> 52) + (low != 0 ? 1u : 0u);
        const uint64_t m = ((uint64_t)0 - (low + (low << 32))) & FP_MASK52;
        const double M = fp_from_bits(m | FP_BITS_2P52) - 4503599627370496.0;
        const double h0 = fp_fma_rz(M, Q[0], C1);
        c[k + 1] += fp_to_bits(h0) + carry;
#pragma unroll
        for (int j = 1; j < 5; j++) {
            const double hi = fp_fma_rz(M, Q[j], C1);
            const double lo = fp_fma_rz(M, Q[j], C2 - hi);
            c[k + j] += fp_to_bits(lo);
            c[k + j + 1] += fp_to_bits(hi);
        }
    }
    // columns 5..9 -> normalised 52-bit limbs (value < 2q < 2^256)
    uint64_t l[5], cin = 0;
#pragma unroll
    for (int i = 0; i < 5; i++) {
        const uint64_t v = c[5 + i] + cin;
        l[i] = i < 4 ? (v & FP_MASK52) : v;
        cin = v >> 52;
    }
    uint32_t w[8];
    w[0] = (uint32_t)l[0];
    w[1] = (uint32_t)((l[0] >> 32) | (l[1] << 20));
    w[2] = (uint32_t)(l[1] >> 12);
    w[3] = (uint32_t)((l[1] >> 44) | (l[2] << 8));
    w[4] = (uint32_t)((l[2] >> 24) | (l[3] << 28));
    w[5] = (uint32_t)(l[3] >> 4);
    w[6] = (uint32_t)((l[3] >> 36) | (l[4] << 16));
    w[7] = (uint32_t)(l[4] >> 16);
    uint32_t s[8];
    const uint32_t borrow = sub_q(s, w);
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = borrow ? w[i] : s[i];
}

#if defined(__CUDA_ARCH__) && !defined(JJS_INLINE_FIELD)
__device__ __noinline__ fq fq_mul_fp_fn(fq a, fq b) {
    fq r;
    fq_mul_fp_inl(r, a, b);
    return r;
}
JJS_HD void fq_mul_fp(fq& r, const fq& a, const fq& b) { r = fq_mul_fp_fn(a, b); }
// Two independent products in one function body, one per pipe: the instruction streams of the integer and the
// floating-point multiplier have no dependences on each other, so ptxas interleaves them and a single warp keeps both
// pipes busy (separate calls do not overlap: the warps of an SM sub-partition move through the same code in step).
struct fq2 {
    fq x, y;
};
__device__ __noinline__ fq2 fq_mul_pair_fn(fq a0, fq b0, fq a1, fq b1) {
    fq2 r;
    fq_mul_inl(r.x, a0, b0);
    fq_mul_fp_inl(r.y, a1, b1);
    return r;
}
JJS_HD void fq_mul_pair(fq& r0, const fq& a0, const fq& b0, fq& r1, const fq& a1, const fq& b1) {
    fq2 r = fq_mul_pair_fn(a0, b0, a1, b1);
    r0 = r.x;
    r1 = r.y;
}
#else
JJS_HD void fq_mul_fp(fq& r, const fq& a, const fq& b) { fq_mul_fp_inl(r, a, b); }
JJS_HD void fq_mul_pair(fq& r0, const fq& a0, const fq& b0, fq& r1, const fq& a1, const fq& b1) {
    fq_mul_inl(r0, a0, b0);
    fq_mul_fp_inl(r1, a1, b1);
}
#endif

}  // namespace jjs
