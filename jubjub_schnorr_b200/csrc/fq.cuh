// BLS12-381 scalar field Fq (dusk-bls12_381 BlsScalar) on 8 x 32-bit Montgomery limbs for sm_100a.
//
// Every multi-limb carry chain is ONE inline-PTX block of mad.lo.cc / madc.hi.cc pairs, which ptxas
// fuses into IMAD.WIDE.U32(.X) with predicate carries (one integer-pipe issue per 32x32->64 product).
// Each block has a plain C++ twin under !__CUDA_ARCH__ so tests/hostsim can run the very same
// algorithms (reduction schedule, curve formulas, Poseidon rewrite) on the CPU.  That twin is test
// scaffolding: the shipped library never executes it.
//
// Replaces, for the verify direction, the Fq arithmetic the reference reaches through
// dusk_bls12_381::BlsScalar (call sites: reference src/signatures.rs:127-139, src/keys/public.rs:128).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define JJS_HD __host__ __device__ __forceinline__
#define JJS_HD_NOINLINE __host__ __device__ __noinline__
#else
#define JJS_HD inline
#define JJS_HD_NOINLINE
#endif

#if !defined(__CUDACC__)
struct uint4 {  // host twin of the CUDA vector type (tests/hostsim only)
    uint32_t x, y, z, w;
};
#endif

namespace jjs {

#if !defined(__CUDA_ARCH__)
// host twin only: a carry the device code drops as provably zero was not zero (tests/hostsim aborts loudly)
[[noreturn]] inline void jjs_host_bound_violation() { __builtin_trap(); }
#endif

// q = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001, little-endian 32-bit limbs.
// -q^-1 mod 2^32 = 0xffffffff, so the Montgomery quotient digit is just the negated low limb.
#define JJS_Q_LIMBS {0x00000001u, 0xffffffffu, 0xfffe5bfeu, 0x53bda402u, 0x09a1d805u, 0x3339d808u, 0x299d7d48u, 0x73eda753u}

struct alignas(16) fq {
    uint32_t l[8];
};

JJS_HD uint32_t q_limb(int i) {
    constexpr uint32_t Q[8] = JJS_Q_LIMBS;
    return Q[i];
}

// ---------------------------------------------------------------------------------------------
// carry-chain building blocks (device: one asm block each; host: exact C++ twin)
// ---------------------------------------------------------------------------------------------

// acc[0..2N) += a[k] * b with product k at limb pair (2k, 2k+1).  N in 1..4.  The carry out of limb 2N - 1 is
//   CAP == 2: added to acc[2N];   CAP == 1: stored to acc[2N] (the caller knows that limb is still untouched);
//   CAP == 0: dropped (the caller knows it is zero, see mul_wide / sqr_wide).
// ptxas turns every captured carry into a SEL plus, when it is accumulated, an IMAD.X and often an IMAD.MOV that zeroes
// the upper half of the 64-bit accumulator pair -- instructions on the multiply pipe -- so captures are only taken
// where a carry can exist.
template <int N, int CAP = 2>
JJS_HD void mad_row(uint32_t* acc, const uint32_t* a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    if (N == 1 && CAP == 0) {
        asm("mad.lo.cc.u32 %0, %2, %3, %0;\n\t"
            "madc.hi.u32 %1, %2, %3, %1;"
            : "+r"(acc[0]), "+r"(acc[1])
            : "r"(a[0]), "r"(b));
    }
    else if (N == 1 && CAP == 1) {
        asm("mad.lo.cc.u32 %0, %3, %4, %0;\n\t"
            "madc.hi.cc.u32 %1, %3, %4, %1;\n\t"
            "addc.u32 %2, 0, 0;"
            : "+r"(acc[0]), "+r"(acc[1]), "=r"(acc[2])
            : "r"(a[0]), "r"(b));
    }
    else if (N == 1 && CAP == 2) {
        asm("mad.lo.cc.u32 %0, %3, %4, %0;\n\t"
            "madc.hi.cc.u32 %1, %3, %4, %1;\n\t"
            "addc.u32 %2, %2, 0;"
            : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2])
            : "r"(a[0]), "r"(b));
    }
    else if (N == 2 && CAP == 0) {
        asm("mad.lo.cc.u32 %0, %4, %6, %0;\n\t"
            "madc.hi.cc.u32 %1, %4, %6, %1;\n\t"
            "madc.lo.cc.u32 %2, %5, %6, %2;\n\t"
            "madc.hi.u32 %3, %5, %6, %3;"
            : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3])
            : "r"(a[0]), "r"(a[1]), "r"(b));
    }
    else if (N == 2 && CAP == 1) {
        asm("mad.lo.cc.u32 %0, %5, %7, %0;\n\t"
            "madc.hi.cc.u32 %1, %5, %7, %1;\n\t"
            "madc.lo.cc.u32 %2, %6, %7, %2;\n\t"
            "madc.hi.cc.u32 %3, %6, %7, %3;\n\t"
            "addc.u32 %4, 0, 0;"
            : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "=r"(acc[4])
            : "r"(a[0]), "r"(a[1]), "r"(b));
    }
    else if (N == 2 && CAP == 2) {
        asm("mad.lo.cc.u32 %0, %5, %7, %0;\n\t"
            "madc.hi.cc.u32 %1, %5, %7, %1;\n\t"
            "madc.lo.cc.u32 %2, %6, %7, %2;\n\t"
            "madc.hi.cc.u32 %3, %6, %7, %3;\n\t"
            "addc.u32 %4, %4, 0;"
            : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4])
            : "r"(a[0]), "r"(a[1]), "r"(b));
    }
    else if (N == 3 && CAP == 0) {
        asm("mad.lo.cc.u32 %0, %6, %9, %0;\n\t"
            "madc.hi.cc.u32 %1, %6, %9, %1;\n\t"
            "madc.lo.cc.u32 %2, %7, %9, %2;\n\t"
            "madc.hi.cc.u32 %3, %7, %9, %3;\n\t"
            "madc.lo.cc.u32 %4, %8, %9, %4;\n\t"
            "madc.hi.u32 %5, %8, %9, %5;"
            : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5])
            : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(b));
    }
    else if (N == 3 && CAP == 1) {
        asm("mad.lo.cc.u32 %0, %7, %10, %0;\n\t"
            "madc.hi.cc.u32 %1, %7, %10, %1;\n\t"
            "madc.lo.cc.u32 %2, %8, %10, %2;\n\t"
            "madc.hi.cc.u32 %3, %8, %10, %3;\n\t"
            "madc.lo.cc.u32 %4, %9, %10, %4;\n\t"
            "madc.hi.cc.u32 %5, %9, %10, %5;\n\t"
            "addc.u32 %6, 0, 0;"
            : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "=r"(acc[6])
            : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(b));
    }
    else if (N == 3 && CAP == 2) {
        asm("mad.lo.cc.u32 %0, %7, %10, %0;\n\t"
            "madc.hi.cc.u32 %1, %7, %10, %1;\n\t"
            "madc.lo.cc.u32 %2, %8, %10, %2;\n\t"
            "madc.hi.cc.u32 %3, %8, %10, %3;\n\t"
            "madc.lo.cc.u32 %4, %9, %10, %4;\n\t"
            "madc.hi.cc.u32 %5, %9, %10, %5;\n\t"
            "addc.u32 %6, %6, 0;"
            : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6])
            : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(b));
    }
    else if (N == 4 && CAP == 0) {
        asm("mad.lo.cc.u32 %0, %8, %12, %0;\n\t"
            "madc.hi.cc.u32 %1, %8, %12, %1;\n\t"
            "madc.lo.cc.u32 %2, %9, %12, %2;\n\t"
            "madc.hi.cc.u32 %3, %9, %12, %3;\n\t"
            "madc.lo.cc.u32 %4, %10, %12, %4;\n\t"
            "madc.hi.cc.u32 %5, %10, %12, %5;\n\t"
            "madc.lo.cc.u32 %6, %11, %12, %6;\n\t"
            "madc.hi.u32 %7, %11, %12, %7;"
            : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7])
            : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b));
    }
    else if (N == 4 && CAP == 1) {
        asm("mad.lo.cc.u32 %0, %9, %13, %0;\n\t"
            "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
            "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
            "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
            "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
            "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
            "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
            "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
            "addc.u32 %8, 0, 0;"
            : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7]), "=r"(acc[8])
            : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b));
    }
    else if (N == 4 && CAP == 2) {
        asm("mad.lo.cc.u32 %0, %9, %13, %0;\n\t"
            "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
            "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
            "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
            "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
            "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
            "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
            "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
            "addc.u32 %8, %8, 0;"
            : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7]), "+r"(acc[8])
            : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b));
    }
#else
    uint64_t c = 0;
    for (int k = 0; k < N; k++) {
        uint64_t p = (uint64_t)a[k] * b;
        uint64_t lo = (uint64_t)acc[2 * k] + (uint32_t)p + c;
        acc[2 * k] = (uint32_t)lo;
        uint64_t hi = (uint64_t)acc[2 * k + 1] + (uint32_t)(p >> 32) + (lo >> 32);
        acc[2 * k + 1] = (uint32_t)hi;
        c = hi >> 32;
    }
    if (CAP == 2) acc[2 * N] += (uint32_t)c;
    else if (CAP == 1) { if (acc[2 * N] != 0) jjs_host_bound_violation(); acc[2 * N] = (uint32_t)c; }
    else if (c != 0) jjs_host_bound_violation();   // the twin checks the bounds the device code relies on
#endif
}

// t[0..15] += a_i^2 at limb pair (2i, 2i+1), one carry chain over all 16 limbs (no carry out by construction).
JJS_HD void mad_diag8(uint32_t* t, const uint32_t* a) {
#if defined(__CUDA_ARCH__)
    asm("mad.lo.cc.u32 %0, %16, %16, %0;\n\t"
        "madc.hi.cc.u32 %1, %16, %16, %1;\n\t"
        "madc.lo.cc.u32 %2, %17, %17, %2;\n\t"
        "madc.hi.cc.u32 %3, %17, %17, %3;\n\t"
        "madc.lo.cc.u32 %4, %18, %18, %4;\n\t"
        "madc.hi.cc.u32 %5, %18, %18, %5;\n\t"
        "madc.lo.cc.u32 %6, %19, %19, %6;\n\t"
        "madc.hi.cc.u32 %7, %19, %19, %7;\n\t"
        "madc.lo.cc.u32 %8, %20, %20, %8;\n\t"
        "madc.hi.cc.u32 %9, %20, %20, %9;\n\t"
        "madc.lo.cc.u32 %10, %21, %21, %10;\n\t"
        "madc.hi.cc.u32 %11, %21, %21, %11;\n\t"
        "madc.lo.cc.u32 %12, %22, %22, %12;\n\t"
        "madc.hi.cc.u32 %13, %22, %22, %13;\n\t"
        "madc.lo.cc.u32 %14, %23, %23, %14;\n\t"
        "madc.hi.u32 %15, %23, %23, %15;"
        : "+r"(t[0]), "+r"(t[1]), "+r"(t[2]), "+r"(t[3]), "+r"(t[4]), "+r"(t[5]), "+r"(t[6]), "+r"(t[7]), "+r"(t[8]),
          "+r"(t[9]), "+r"(t[10]), "+r"(t[11]), "+r"(t[12]), "+r"(t[13]), "+r"(t[14]), "+r"(t[15])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]));
#else
    uint64_t c = 0;
    for (int k = 0; k < 8; k++) {
        uint64_t p = (uint64_t)a[k] * a[k];
        uint64_t lo = (uint64_t)t[2 * k] + (uint32_t)p + c;
        t[2 * k] = (uint32_t)lo;
        uint64_t hi = (uint64_t)t[2 * k + 1] + (uint32_t)(p >> 32) + (lo >> 32);
        t[2 * k + 1] = (uint32_t)hi;
        c = hi >> 32;
    }
#endif
}

// r[0..7] = x[0..7] + y[0..7]; returns carry out in {0,1}.
JJS_HD uint32_t add8(uint32_t* r, const uint32_t* x, const uint32_t* y) {
    uint32_t cout;
#if defined(__CUDA_ARCH__)
    asm("add.cc.u32 %0, %9, %17;\n\t"
        "addc.cc.u32 %1, %10, %18;\n\t"
        "addc.cc.u32 %2, %11, %19;\n\t"
        "addc.cc.u32 %3, %12, %20;\n\t"
        "addc.cc.u32 %4, %13, %21;\n\t"
        "addc.cc.u32 %5, %14, %22;\n\t"
        "addc.cc.u32 %6, %15, %23;\n\t"
        "addc.cc.u32 %7, %16, %24;\n\t"
        "addc.u32 %8, 0, 0;"
        : "=&r"(r[0]), "=&r"(r[1]), "=&r"(r[2]), "=&r"(r[3]), "=&r"(r[4]), "=&r"(r[5]), "=&r"(r[6]), "=&r"(r[7]), "=&r"(cout)
        : "r"(x[0]), "r"(x[1]), "r"(x[2]), "r"(x[3]), "r"(x[4]), "r"(x[5]), "r"(x[6]), "r"(x[7]), "r"(y[0]), "r"(y[1]),
          "r"(y[2]), "r"(y[3]), "r"(y[4]), "r"(y[5]), "r"(y[6]), "r"(y[7]));
#else
    uint64_t c = 0;
    for (int i = 0; i < 8; i++) {
        c += (uint64_t)x[i] + y[i];
        r[i] = (uint32_t)c;
        c >>= 32;
    }
    cout = (uint32_t)c;
#endif
    return cout;
}
// same with a carry in (cin in {0,1})
JJS_HD uint32_t add8c(uint32_t* r, const uint32_t* x, const uint32_t* y, uint32_t cin) {
    uint32_t cout;
#if defined(__CUDA_ARCH__)
    asm("add.cc.u32 %8, %25, 0xffffffff;\n\t"  // CF = cin
        "addc.cc.u32 %0, %9, %17;\n\t"
        "addc.cc.u32 %1, %10, %18;\n\t"
        "addc.cc.u32 %2, %11, %19;\n\t"
        "addc.cc.u32 %3, %12, %20;\n\t"
        "addc.cc.u32 %4, %13, %21;\n\t"
        "addc.cc.u32 %5, %14, %22;\n\t"
        "addc.cc.u32 %6, %15, %23;\n\t"
        "addc.cc.u32 %7, %16, %24;\n\t"
        "addc.u32 %8, 0, 0;"
        : "=&r"(r[0]), "=&r"(r[1]), "=&r"(r[2]), "=&r"(r[3]), "=&r"(r[4]), "=&r"(r[5]), "=&r"(r[6]), "=&r"(r[7]), "=&r"(cout)
        : "r"(x[0]), "r"(x[1]), "r"(x[2]), "r"(x[3]), "r"(x[4]), "r"(x[5]), "r"(x[6]), "r"(x[7]), "r"(y[0]), "r"(y[1]),
          "r"(y[2]), "r"(y[3]), "r"(y[4]), "r"(y[5]), "r"(y[6]), "r"(y[7]), "r"(cin));
#else
    uint64_t c = cin;
    for (int i = 0; i < 8; i++) {
        c += (uint64_t)x[i] + y[i];
        r[i] = (uint32_t)c;
        c >>= 32;
    }
    cout = (uint32_t)c;
#endif
    return cout;
}

// r = x - y over 8 limbs; returns 0xffffffff if the subtraction borrowed, else 0.
JJS_HD uint32_t sub8(uint32_t* r, const uint32_t* x, const uint32_t* y) {
    uint32_t borrow;
#if defined(__CUDA_ARCH__)
    asm("sub.cc.u32 %0, %9, %17;\n\t"
        "subc.cc.u32 %1, %10, %18;\n\t"
        "subc.cc.u32 %2, %11, %19;\n\t"
        "subc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\t"
        "subc.cc.u32 %5, %14, %22;\n\t"
        "subc.cc.u32 %6, %15, %23;\n\t"
        "subc.cc.u32 %7, %16, %24;\n\t"
        "subc.u32 %8, 0, 0;"
        : "=&r"(r[0]), "=&r"(r[1]), "=&r"(r[2]), "=&r"(r[3]), "=&r"(r[4]), "=&r"(r[5]), "=&r"(r[6]), "=&r"(r[7]), "=&r"(borrow)
        : "r"(x[0]), "r"(x[1]), "r"(x[2]), "r"(x[3]), "r"(x[4]), "r"(x[5]), "r"(x[6]), "r"(x[7]), "r"(y[0]), "r"(y[1]),
          "r"(y[2]), "r"(y[3]), "r"(y[4]), "r"(y[5]), "r"(y[6]), "r"(y[7]));
#else
    int64_t c = 0;
    for (int i = 0; i < 8; i++) {
        c += (int64_t)x[i] - (int64_t)y[i];
        r[i] = (uint32_t)c;
        c >>= 32;
    }
    borrow = (uint32_t)c;
#endif
    return borrow;
}

// r = x - q; returns 0xffffffff if x < q (borrow), else 0.  q limbs are immediates.
JJS_HD uint32_t sub_q(uint32_t* r, const uint32_t* x) {
    uint32_t borrow;
#if defined(__CUDA_ARCH__)
    asm("sub.cc.u32 %0, %9, 0x00000001;\n\t"
        "subc.cc.u32 %1, %10, 0xffffffff;\n\t"
        "subc.cc.u32 %2, %11, 0xfffe5bfe;\n\t"
        "subc.cc.u32 %3, %12, 0x53bda402;\n\t"
        "subc.cc.u32 %4, %13, 0x09a1d805;\n\t"
        "subc.cc.u32 %5, %14, 0x3339d808;\n\t"
        "subc.cc.u32 %6, %15, 0x299d7d48;\n\t"
        "subc.cc.u32 %7, %16, 0x73eda753;\n\t"
        "subc.u32 %8, 0, 0;"
        : "=&r"(r[0]), "=&r"(r[1]), "=&r"(r[2]), "=&r"(r[3]), "=&r"(r[4]), "=&r"(r[5]), "=&r"(r[6]), "=&r"(r[7]), "=&r"(borrow)
        : "r"(x[0]), "r"(x[1]), "r"(x[2]), "r"(x[3]), "r"(x[4]), "r"(x[5]), "r"(x[6]), "r"(x[7]));
#else
    int64_t c = 0;
    for (int i = 0; i < 8; i++) {
        c += (int64_t)x[i] - (int64_t)q_limb(i);
        r[i] = (uint32_t)c;
        c >>= 32;
    }
    borrow = (uint32_t)c;
#endif
    return borrow;
}

#if !defined(__CUDA_ARCH__)
inline bool ge_q_host(const uint32_t* x) {  // host twin only: x >= q
    uint32_t s[8];
    return sub_q(s, x) == 0;
}
#endif
// r = x + (q & mask) over 8 limbs (mask is 0 or 0xffffffff); carry out dropped.
// Device: the eight additions are PREDICATED on the mask instead of adding a masked copy of q -- five instructions fewer
// in the tail of every product, squaring and subtraction (-DJJS_PRED_ADDQ=0 restores the masked form).
#ifndef JJS_PRED_ADDQ
#define JJS_PRED_ADDQ 1
#endif
JJS_HD void add_q_masked(uint32_t* r, const uint32_t* x, uint32_t mask) {
#if defined(__CUDA_ARCH__) && JJS_PRED_ADDQ
#pragma unroll
    for (int i = 0; i < 8; i++) r[i] = x[i];
    asm("{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.u32 p, %8, 0;\n\t"
        "@p add.cc.u32 %0, %0, 0x00000001;\n\t"
        "@p addc.cc.u32 %1, %1, 0xffffffff;\n\t"
        "@p addc.cc.u32 %2, %2, 0xfffe5bfe;\n\t"
        "@p addc.cc.u32 %3, %3, 0x53bda402;\n\t"
        "@p addc.cc.u32 %4, %4, 0x09a1d805;\n\t"
        "@p addc.cc.u32 %5, %5, 0x3339d808;\n\t"
        "@p addc.cc.u32 %6, %6, 0x299d7d48;\n\t"
        "@p addc.u32 %7, %7, 0x73eda753;\n\t"
        "}"
        : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7])
        : "r"(mask));
#else
    constexpr uint32_t Q[8] = JJS_Q_LIMBS;
    uint32_t y[8];
#pragma unroll
    for (int i = 0; i < 8; i++) y[i] = Q[i] & mask;
    add8(r, x, y);
#endif
}
// (ev, od) split representation of a multi-limb value:  V = sum ev[k] 2^(32k) + sum od[k] 2^(32(k+1)).  A reduction step
// expects ev[0] == 0 (its low limb was cancelled by the previous step; for the first step the caller presents T shifted up
// by one limb), shifts V right by 32 bits and appends `inject` as the new top limb.
// ---- additive reduction on the complement of q --------------------------------------------------------------------
// q == 1 (mod 2^32) makes the Montgomery quotient digit of a value with low limb e0 equal to -e0.  Instead of negating
// it (and patching the multiplier-free products of q's two low limbs with m - [m != 0]) the step below adds
//     e0 * qbar,  qbar = 2^256 - q = (0xffffffff, 0, 0x0001a401, 0xac425bfd, 0xf65e27fa, 0xccc627f7, 0xd66282b7, 0x8c1258ac),
// which cancels the low limb just the same (e0 + e0 * 0xffffffff = e0 * 2^32, exactly: limb 1 receives e0 and nothing
// else, qbar's limb 1 being zero) but is e0 * 2^256 too much.  Over the eight steps the excess adds up to M * 2^256 with
// M = sum e0_k 2^(32 k), i.e. the reduced value comes out as r + M, and one 8-limb subtraction at the end removes it:
//     r = (T + M qbar) / 2^256 - M = (T - M q) / 2^256  in (-q, q)  for T < q 2^256;  a borrow means "add q".
// Same six wide multiplies per step; the negation, the min and the extra subtraction of every step are gone (they
// compiled to IMAD.MOV / VIMNMX / IMAD.IADD, two of the three on the multiply pipe).
JJS_HD void redc_step2(const uint32_t* ev, uint32_t* od, uint32_t* n, uint32_t inject, uint32_t& e0_out) {
#if defined(__CUDA_ARCH__)
    uint32_t e0;
    asm("add.cc.u32 %9, %10, %11;\n\t"           // e0 = od0 + ev1, CF feeds limb 1
        "addc.cc.u32 %0, %12, %9;\n\t"           // n0 = ev2 + e0 + CF
        "addc.cc.u32 %1, %13, 0;\n\t"            // n1 = ev3 + CF
        "madc.lo.cc.u32 %2, %9, 0xac425bfd, %14;\n\t"
        "madc.hi.cc.u32 %3, %9, 0xac425bfd, %15;\n\t"
        "madc.lo.cc.u32 %4, %9, 0xccc627f7, %16;\n\t"
        "madc.hi.cc.u32 %5, %9, 0xccc627f7, %17;\n\t"
        "madc.lo.cc.u32 %6, %9, 0x8c1258ac, %18;\n\t"
        "madc.hi.cc.u32 %7, %9, 0x8c1258ac, %19;\n\t"
        "addc.u32 %8, 0, 0;"                       // the running value stays below 2^256 + qbar: this limb is 0 or 1
        : "=&r"(n[0]), "=&r"(n[1]), "=&r"(n[2]), "=&r"(n[3]), "=&r"(n[4]), "=&r"(n[5]), "=&r"(n[6]), "=&r"(n[7]), "=&r"(n[8]), "=&r"(e0)
        : "r"(od[0]), "r"(ev[1]), "r"(ev[2]), "r"(ev[3]), "r"(ev[4]), "r"(ev[5]), "r"(ev[6]), "r"(ev[7]), "r"(ev[8]), "r"(inject));
    asm("mad.lo.cc.u32 %0, %7, 0x0001a401, %0;\n\t"
        "madc.hi.cc.u32 %1, %7, 0x0001a401, %1;\n\t"
        "madc.lo.cc.u32 %2, %7, 0xf65e27fa, %2;\n\t"
        "madc.hi.cc.u32 %3, %7, 0xf65e27fa, %3;\n\t"
        "madc.lo.cc.u32 %4, %7, 0xd66282b7, %4;\n\t"
        "madc.hi.cc.u32 %5, %7, 0xd66282b7, %5;\n\t"
        "addc.u32 %6, %6, 0;"
        : "+r"(od[2]), "+r"(od[3]), "+r"(od[4]), "+r"(od[5]), "+r"(od[6]), "+r"(od[7]), "+r"(od[8])
        : "r"(e0));
    od[0] = 0;
    e0_out = e0;
#else
    constexpr uint32_t QB[8] = {0xffffffffu, 0u, 0x0001a401u, 0xac425bfdu, 0xf65e27fau, 0xccc627f7u, 0xd66282b7u, 0x8c1258acu};
    uint64_t s = (uint64_t)od[0] + ev[1];
    uint32_t e0 = (uint32_t)s;
    uint64_t c = s >> 32;
    uint64_t t = (uint64_t)ev[2] + e0 + c;
    n[0] = (uint32_t)t;
    t = (uint64_t)ev[3] + (t >> 32);
    n[1] = (uint32_t)t;
    c = t >> 32;
    const uint32_t addend[6] = {ev[4], ev[5], ev[6], ev[7], ev[8], inject};
    for (int k = 0; k < 3; k++) {
        uint64_t p = (uint64_t)e0 * QB[2 * k + 3];
        uint64_t lo = (uint64_t)addend[2 * k] + (uint32_t)p + c;
        n[2 * k + 2] = (uint32_t)lo;
        uint64_t hi = (uint64_t)addend[2 * k + 1] + (uint32_t)(p >> 32) + (lo >> 32);
        n[2 * k + 3] = (uint32_t)hi;
        c = hi >> 32;
    }
    n[8] = (uint32_t)c;
    c = 0;
    for (int k = 0; k < 3; k++) {
        uint64_t p = (uint64_t)e0 * QB[2 * k + 2];
        uint64_t lo = (uint64_t)od[2 * k + 2] + (uint32_t)p + c;
        od[2 * k + 2] = (uint32_t)lo;
        uint64_t hi = (uint64_t)od[2 * k + 3] + (uint32_t)(p >> 32) + (lo >> 32);
        od[2 * k + 3] = (uint32_t)hi;
        c = hi >> 32;
    }
    od[8] += (uint32_t)c;
    od[0] = 0;
    e0_out = e0;
#endif
}

JJS_HD uint32_t funnel_l1(uint32_t lo, uint32_t hi) {  // (hi:lo << 1) >> 32
#if defined(__CUDA_ARCH__)
    return __funnelshift_l(lo, hi, 1);
#else
    return (hi << 1) | (lo >> 31);
#endif
}

// ---------------------------------------------------------------------------------------------
// wide products and Montgomery reduction
// ---------------------------------------------------------------------------------------------

// a * b as an even and an odd accumulator (schoolbook, 64 wide multiplies in 16 carry chains):
//     a b = sum_k E[k] 2^(32 k) + sum_k O[k] 2^(32 (k + 1)),   E: 17 limbs, O: 15.
JJS_HD void mul_wide_eo(uint32_t* E, uint32_t* O, const uint32_t* a, const uint32_t* b) {
#pragma unroll
    for (int i = 0; i < 17; i++) E[i] = 0;
#pragma unroll
    for (int i = 0; i < 15; i++) O[i] = 0;
    const uint32_t ae[4] = {a[0], a[2], a[4], a[6]}, ao[4] = {a[1], a[3], a[5], a[7]};
    // Carry out of a row: all products accumulated so far lie on diagonals <= p + 6 for a row at offset p, so limbs
    // [p, p + 8) can only overflow when TWO products sit on diagonal p + 6, i.e. for the second row at an offset (the
    // lower diagonals add < 2^195 to a first product < 2^256 - 2^225).  The first row at an offset therefore drops its
    // carry, the second stores it into the limb above, which no earlier row has touched.
    mad_row<4, 0>(E + 0, ae, b[0]);
    mad_row<4, 0>(O + 0, ao, b[0]);
    mad_row<4, 1>(O + 0, ae, b[1]);
    mad_row<4, 0>(E + 2, ao, b[1]);
#pragma unroll
    for (int i = 2; i < 8; i += 2) {
        mad_row<4, 1>(E + i, ae, b[i]);
        mad_row<4, 0>(O + i, ao, b[i]);
        mad_row<4, 1>(O + i, ae, b[i + 1]);
        mad_row<4, 0>(E + i + 2, ao, b[i + 1]);
    }
}
// t[0..15] = a * b: the two accumulators merged
JJS_HD void mul_wide(uint32_t* t, const uint32_t* a, const uint32_t* b) {
    uint32_t E[17], O[15];  // E[k]: limb k;  O[k]: limb k+1
    mul_wide_eo(E, O, a, b);
    // t = E + (O << 32)
    t[0] = E[0];
    uint32_t c = add8(t + 1, E + 1, O);
    uint32_t x[8] = {E[9], E[10], E[11], E[12], E[13], E[14], E[15], 0};
    uint32_t y[8] = {O[8], O[9], O[10], O[11], O[12], O[13], O[14], 0};
    uint32_t hi[8];
    add8c(hi, x, y, c);
#pragma unroll
    for (int i = 0; i < 7; i++) t[9 + i] = hi[i];
}

// t[0..15] = a^2 (28 off-diagonal + 8 diagonal wide multiplies)
JJS_HD void sqr_wide(uint32_t* t, const uint32_t* a) {
    uint32_t E[16], O[16];  // E[k]: limb k;  O[k]: limb k+1
#pragma unroll
    for (int i = 0; i < 16; i++) { E[i] = 0; O[i] = 0; }
    // row i multiplies a[i] with a[j], j > i: j - i odd -> odd accumulator, j - i even -> even accumulator.
    // Carries out of a row (see mul_wide): only possible when the row's top product is the second one on its diagonal
    // in that accumulator; it is then stored into the limb above, which is still untouched.
    { const uint32_t m[4] = {a[1], a[3], a[5], a[7]}; mad_row<4, 0>(O + 0, m, a[0]); }    // limbs 1..8,   top a0 a7: first on diagonal 7
    { const uint32_t m[3] = {a[2], a[4], a[6]};       mad_row<3, 0>(E + 2, m, a[0]); }    // limbs 2..7,   top a0 a6: first on 6
    { const uint32_t m[3] = {a[2], a[4], a[6]};       mad_row<3, 1>(O + 2, m, a[1]); }    // limbs 3..8,   top a1 a6: second on 7 -> O[8]
    { const uint32_t m[3] = {a[3], a[5], a[7]};       mad_row<3, 0>(E + 4, m, a[1]); }    // limbs 4..9,   top a1 a7: first on 8
    { const uint32_t m[3] = {a[3], a[5], a[7]};       mad_row<3, 0>(O + 4, m, a[2]); }    // limbs 5..10,  top a2 a7: first on 9
    { const uint32_t m[2] = {a[4], a[6]};             mad_row<2, 1>(E + 6, m, a[2]); }    // limbs 6..9,   top a2 a6: second on 8 -> E[10]
    { const uint32_t m[2] = {a[4], a[6]};             mad_row<2, 1>(O + 6, m, a[3]); }    // limbs 7..10,  top a3 a6: second on 9 -> O[10]
    { const uint32_t m[2] = {a[5], a[7]};             mad_row<2, 0>(E + 8, m, a[3]); }    // limbs 8..11,  top a3 a7: first on 10
    { const uint32_t m[2] = {a[5], a[7]};             mad_row<2, 0>(O + 8, m, a[4]); }    // limbs 9..12,  top a4 a7: first on 11
    { const uint32_t m[1] = {a[6]};                   mad_row<1, 1>(E + 10, m, a[4]); }   // limbs 10..11, top a4 a6: second on 10 -> E[12]
    { const uint32_t m[1] = {a[6]};                   mad_row<1, 1>(O + 10, m, a[5]); }   // limbs 11..12, top a5 a6: second on 11 -> O[12]
    { const uint32_t m[1] = {a[7]};                   mad_row<1, 0>(E + 12, m, a[5]); }   // limbs 12..13, top a5 a7: first on 12
    { const uint32_t m[1] = {a[7]};                   mad_row<1, 0>(O + 12, m, a[6]); }   // limbs 13..14, top a6 a7: first on 13
    // s = E + (O << 32): limbs 1..15
    uint32_t s[16];
    s[0] = 0;
    uint32_t lo[8] = {0, E[2], E[3], E[4], E[5], E[6], E[7], E[8]};
    uint32_t c = add8(s + 1, lo, O);
    uint32_t x[8] = {E[9], E[10], E[11], E[12], E[13], E[14], 0, 0};
    uint32_t y[8] = {O[8], O[9], O[10], O[11], O[12], O[13], O[14], 0};
    uint32_t hi[8];
    add8c(hi, x, y, c);
#pragma unroll
    for (int i = 0; i < 7; i++) s[9 + i] = hi[i];
    // t = 2 s + diagonal
#pragma unroll
    for (int i = 15; i >= 1; i--) t[i] = funnel_l1(s[i - 1], s[i]);
    t[0] = 0;
    mad_diag8(t, a);
}

// The eight reduction steps and the final correction, from a state (ev, od) that holds T 2^32 (ev[0] == 0) and the eight high
// limbs of T still to be injected, hi[i] at weight 2^(32 (8 + i)): r = T / 2^256 mod q, fully reduced, for T < q 2^256.
JJS_HD void redc_run(uint32_t* r, uint32_t* ev, uint32_t* od, const uint32_t* hi) {
    uint32_t n[9], M[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        redc_step2(ev, od, n, hi[i], M[i]);
#pragma unroll
        for (int k = 0; k < 9; k++) { ev[k] = od[k]; od[k] = n[k]; }
    }
    // r + M = (ev >> 32) + od < 2^256;  r = that - M, plus q if it went negative
    uint32_t v[8], d[8];
    add8(v, ev + 1, od);
    uint32_t borrow = sub8(d, v, M);
    add_q_masked(r, d, borrow);
}
// r = t / 2^256 mod q for a 16-limb t < q * 2^256, fully reduced.
JJS_HD void redc(uint32_t* r, const uint32_t* t) {
    uint32_t ev[9], od[9];
    ev[0] = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) ev[i + 1] = t[i];
#pragma unroll
    for (int i = 0; i < 9; i++) od[i] = 0;
    redc_run(r, ev, od, t + 8);
}
// The same for a product still held as an even and an odd accumulator (mul_wide_eo).  The reduction state is itself a pair
// of limb vectors offset by one limb, and every step treats it as such (the quotient digit is formed from od[0] + ev[1], all
// carries stay inside the step's two chains and are captured in the top limbs), so the LOW halves of E and O go in
// unmerged -- T 2^32 = sum E[k] 2^(32 (k + 1)) + sum O[k] 2^(32 (k + 2)): od = (E[0..7], 0), ev = (0, 0, O[0..6]) -- and only the
// high halves are merged, into the eight limbs that the steps inject: hi[i] = E[8 + i] + O[7 + i] + carry.  The value the
// state stands for is the same as in redc at every step, so the bounds that let the last addition drop od[8] and its carry
// hold unchanged; what is saved is the merge of the low halves (nine instructions per product).
JJS_HD void redc_eo(uint32_t* r, const uint32_t* E, const uint32_t* O) {
    uint32_t ev[9], od[9], hi[8];
#pragma unroll
    for (int i = 0; i < 8; i++) od[i] = E[i];      // E[k] at 2^(32 (k + 1)) is od[k]: the accumulators keep their register pairs
    od[8] = 0;
    ev[0] = 0;
    ev[1] = 0;
#pragma unroll
    for (int i = 0; i < 7; i++) ev[i + 2] = O[i];  // O[k] at 2^(32 (k + 2)) is ev[k + 2]
    uint32_t x[8] = {E[8], E[9], E[10], E[11], E[12], E[13], E[14], E[15]};
    uint32_t y[8] = {O[7], O[8], O[9], O[10], O[11], O[12], O[13], O[14]};
    uint32_t c = add8(hi, x, y);   // limbs 8..15 of the product, less the carry the unmerged low halves still owe
#if !defined(__CUDA_ARCH__)
    if (c != 0u || E[16] != 0u) jjs_host_bound_violation();   // a b < 2^512: nothing reaches limb 16
#else
    (void)c;
#endif
    redc_run(r, ev, od, hi);
}
// One Montgomery step after the small-integer linear maps of the hash, without any correction: for a 9-limb
//     v' = v + q 2^32,   v < 2^20 q   (the caller's constants carry the q 2^32: HADES_FOLDED_ARK is emitted that way)
// returns r == v / 2^32 (mod q) with 0 < r < q + 2^244 -- ALMOST reduced.  With e0 the quotient digit (it only depends on the low
// limb, which q 2^32 does not touch), (v' + e0 qbar) / 2^32 = r0 + q + e0 2^224 with r0 = (v - e0 q) / 2^32 in (-q, 2^244), so
// taking e0 off limb 7 (mod 2^256) leaves r = r0 + q in (0, q + 2^244).  The subtract-test-add of a full reduction (and even
// the unconditional addition of q of this round's first version) is gone: 32 instructions per lane and round of the
// permutation.  Every consumer takes such a value: the multiplier and the squarer need a b < q 2^256 (and return fully
// reduced results), the next linear layer only uses the limbs as 32-bit integers and keeps its bound, and the permutation
// ends with a product per lane (HADES_UNSCALE), so nothing unreduced leaves it.  The twin checks the range.
JJS_HD void redc_one(uint32_t* r, const uint32_t* v) {
    uint32_t ev[9], od[9], n[9];
    ev[0] = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) ev[i + 1] = v[i];
#pragma unroll
    for (int i = 0; i < 9; i++) od[i] = 0;
    uint32_t e0;
    redc_step2(ev, od, n, v[8], e0);
    add8(r, od + 1, n);
    r[7] -= e0;   // mod 2^256; the true value is in (0, q + 2^244)
#if !defined(__CUDA_ARCH__)
    {   // host twin: 0 < r < q + 2^244
        uint32_t t[8], t2[8], lim[8] = {0, 0, 0, 0, 0, 0, 0, 1u << 20};
        uint32_t b = sub_q(t, r);
        if (b == 0 && sub8(t2, t, lim) == 0) jjs_host_bound_violation();
    }
#endif
}

// ---------------------------------------------------------------------------------------------
// field API (all values fully reduced, Montgomery form unless stated)
// ---------------------------------------------------------------------------------------------
#ifndef JJS_REDC_EO
#define JJS_REDC_EO 1   // reduce the product from its two accumulators without merging their low halves
#endif
JJS_HD void fq_mul_inl(fq& r, const fq& a, const fq& b) {
#if JJS_REDC_EO
    uint32_t E[17], O[15];
    mul_wide_eo(E, O, a.l, b.l);
    redc_eo(r.l, E, O);
#else
    uint32_t t[16];
    mul_wide(t, a.l, b.l);
    redc(r.l, t);
#endif
}
JJS_HD void fq_sqr_inl(fq& r, const fq& a) {
    uint32_t t[16];
    sqr_wide(t, a.l);
    redc(r.l, t);
}
// On the device the multiplier and squarer are real functions (operands travel in registers, no stack
// frame): ~20 moves per call buy a >10x smaller instruction footprint, so the hot loops of the curve and
// hash kernels stay resident in the SM instruction caches.  -DJJS_INLINE_FIELD inlines them instead.
#if defined(__CUDA_ARCH__) && !defined(JJS_INLINE_FIELD)
__device__ __noinline__ fq fq_mul_fn(fq a, fq b) {
    fq r;
    fq_mul_inl(r, a, b);
    return r;
}
__device__ __noinline__ fq fq_sqr_fn(fq a) {
    fq r;
    fq_sqr_inl(r, a);
    return r;
}
JJS_HD void fq_mul(fq& r, const fq& a, const fq& b) { r = fq_mul_fn(a, b); }
JJS_HD void fq_sqr(fq& r, const fq& a) { r = fq_sqr_fn(a); }
// a^(2^n): the loop lives INSIDE the callee.  A chain of calls `x = sqr(x)` costs ~16 moves per squaring, because the caller
// keeps x in other registers than the callee's arguments; the fixed exponentiations of the square root, the subgroup test and
// the inversion are ~90 % such chains (-DJJS_SQR_CHAIN=0: a loop of calls).
#ifndef JJS_SQR_CHAIN
#define JJS_SQR_CHAIN 1
#endif
__device__ __noinline__ fq fq_sqr_n_fn(fq a, int n) {
#pragma unroll 1
    for (int k = 0; k < n; k++) fq_sqr_inl(a, a);
    return a;
}
JJS_HD void fq_sqr_n(fq& r, const fq& a, int n) {
#if JJS_SQR_CHAIN
    r = fq_sqr_n_fn(a, n);
#else
    r = a;
#pragma unroll 1
    for (int k = 0; k < n; k++) r = fq_sqr_fn(r);
#endif
}
#else
JJS_HD void fq_mul(fq& r, const fq& a, const fq& b) { fq_mul_inl(r, a, b); }
JJS_HD void fq_sqr(fq& r, const fq& a) { fq_sqr_inl(r, a); }
JJS_HD void fq_sqr_n(fq& r, const fq& a, int n) {
    r = a;
    for (int k = 0; k < n; k++) fq_sqr_inl(r, r);
}
#endif
JJS_HD void fq_add(fq& r, const fq& a, const fq& b) {
    uint32_t v[8], s[8];
    add8(v, a.l, b.l);  // < 2q < 2^256
    uint32_t borrow = sub_q(s, v);
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = borrow ? v[i] : s[i];
}
// a + b WITHOUT the reduction: the sum of two reduced elements is below 2q < 2^256 and may be used as ONE operand of a
// product whose other operand is reduced (a b < 2 q^2 < q 2^256 is all the Montgomery reduction needs; its result is fully
// reduced again).  Never as an operand of a squaring, an addition or a comparison.  Saves the 17 instructions of the
// conditional subtraction (-DJJS_LAZY_ADD=0: same as fq_add).
#ifndef JJS_LAZY_ADD
#define JJS_LAZY_ADD 1
#endif
JJS_HD void fq_add_lazy(fq& r, const fq& a, const fq& b) {
#if JJS_LAZY_ADD
    uint32_t v[8];
    uint32_t c = add8(v, a.l, b.l);
#if !defined(__CUDA_ARCH__)
    if (c != 0 || ge_q_host(a.l) || ge_q_host(b.l)) jjs_host_bound_violation();   // the twin checks the contract
#else
    (void)c;
#endif
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = v[i];
#else
    fq_add(r, a, b);
#endif
}
JJS_HD void fq_sub(fq& r, const fq& a, const fq& b) {
    uint32_t v[8];
    uint32_t borrow = sub8(v, a.l, b.l);
    add_q_masked(r.l, v, borrow);
}
JJS_HD void fq_dbl(fq& r, const fq& a) { fq_add(r, a, a); }
JJS_HD void fq_neg(fq& r, const fq& a) {
    fq z;
#pragma unroll
    for (int i = 0; i < 8; i++) z.l[i] = 0;
    fq_sub(r, z, a);
}
JJS_HD bool fq_is_zero(const fq& a) {
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) x |= a.l[i];
    return x == 0;
}
JJS_HD bool fq_eq(const fq& a, const fq& b) {
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) x |= a.l[i] ^ b.l[i];
    return x == 0;
}
JJS_HD void fq_zero(fq& r) {
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = 0;
}
// Montgomery R = 2^256 mod q
JJS_HD void fq_one(fq& r) {
    constexpr uint32_t R1[8] = {0xfffffffeu, 0x00000001u, 0x00034802u, 0x5884b7fau, 0xecbc4ff5u, 0x998c4fefu, 0xacc5056fu, 0x1824b159u};
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = R1[i];
}
// canonical integer (< q) -> Montgomery: multiply by R^2
JJS_HD void fq_to_mont(fq& r, const fq& a) {
    constexpr uint32_t R2[8] = {0xf3f29c6du, 0xc999e990u, 0x87925c23u, 0x2b6cedcbu, 0x7254398fu, 0x05d31496u, 0x9f59ff11u, 0x0748d9d9u};
    fq r2;
#pragma unroll
    for (int i = 0; i < 8; i++) r2.l[i] = R2[i];
    fq_mul(r, a, r2);
}
// Montgomery -> canonical integer
JJS_HD void fq_from_mont(fq& r, const fq& a) {
    uint32_t t[16];
#pragma unroll
    for (int i = 0; i < 8; i++) { t[i] = a.l[i]; t[8 + i] = 0; }
    redc(r.l, t);
}
// x >= q ?  (x any 256-bit value)
JJS_HD bool ge_q(const uint32_t* x) {
    uint32_t s[8];
    return sub_q(s, x) == 0;
}

}  // namespace jjs
