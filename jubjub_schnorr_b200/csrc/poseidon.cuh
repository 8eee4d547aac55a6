// Poseidon challenge hash of the verify path: Hades permutation (width 5, x^5, 4 + 60 + 4 rounds) and the
// SAFE sponge (rate 4, capacity lane 0 = IO-pattern tag), one sponge per thread, for sm_100a.
//
// Replaces dusk_poseidon::Hash::digest_truncated(Domain::Other, ..) at reference src/signatures.rs:130,
// src/signatures/double.rs:162, src/signatures/var_gen.rs:130 and src/multisig.rs:408.
//
// The permutation is algebraically the reference's, restructured for the integer pipe:
//   * the Cauchy MDS matrix R/(i+j+5) is (R/360360) times an INTEGER matrix with 17-bit entries, so the
//     linear layer is 25 (8-limb x 32-bit) multiply-accumulates plus one 2^-32 Montgomery step per lane
//     instead of 25 full field multiplications;
//   * lanes are kept as lam * x with a per-round constant lam that absorbs the dropped factors (R/360360,
//     2^-32, the R^-4 of the Montgomery S-box); round constants are pre-scaled and folded into the
//     accumulator of the preceding linear layer; in partial rounds lane 4 is brought back to the common
//     scale with one extra Montgomery product (HADES_SBOX_FIX).
// tools/gen_device_constants.py derives the constants and proves the schedule against the reference form.
#pragma once
#include "consts.cuh"
#include "fq.cuh"
#include <cmath>
#include <cstring>

namespace jjs {

JJS_HD void sbox5(fq& x) {
    fq x2, x4;
    fq_sqr(x2, x);
    fq_sqr(x4, x2);
    fq_mul(x, x4, x);
}

// out = 2^-32 * (ark + sum_k coef[k] * s[k])  mod q
// (ark, ark_top): the nine limbs of the round constant plus q 2^32, see redc_one
JJS_HD void mds_lane(fq& out, const fq* s, const uint32_t* ark, uint32_t ark_top, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t c4) {
    uint32_t E[9], O[9];
#pragma unroll
    for (int i = 0; i < 8; i++) { E[i] = ark[i]; O[i] = 0; }
    E[8] = ark_top;
    O[8] = 0;
    const uint32_t coef[5] = {c0, c1, c2, c3, c4};
#pragma unroll
    for (int k = 0; k < 5; k++) {
        const uint32_t ae[4] = {s[k].l[0], s[k].l[2], s[k].l[4], s[k].l[6]};
        const uint32_t ao[4] = {s[k].l[1], s[k].l[3], s[k].l[5], s[k].l[7]};
        mad_row<4>(E, ae, coef[k]);
        mad_row<4>(O, ao, coef[k]);
    }
    // v = E + (O << 32), 9 limbs
    uint32_t v[9];
    v[0] = E[0];
    uint32_t c = add8(v + 1, E + 1, O);
    (void)c;  // E[8] + O[7] + carry cannot overflow: the sum is < 2^288
    redc_one(out.l, v);
}

// The same linear layer on the FP64 pipe, which B200 issues beside the integer multiply pipe
// (profiles/r01_microbench_fp64.json: DFMA at ~1.35 cycles per warp instruction per scheduler against ~4.2 for
// IMAD.WIDE, the two overlapping almost completely).  A 17-bit matrix entry times a 32-bit limb is below 2^49, so
// a column sum  ark_i + sum_k N[o][k] * s[k].l[i]  < 2^50.5 is EXACT in double precision: every limb is lifted to
// a double (2^52 + x, bit pattern 0x43300000:x, minus 2^52), one DFMA per (output lane, input lane, limb)
// accumulates on top of 2^52 + ark_i, and the mantissa of the result is the integer column sum.  Columns are
// then carried into nine 32-bit limbs and finished by the same single Montgomery step as mds_lane.  This takes
// 200 of the 235 wide multiplies of a linear layer off the integer pipe.
JJS_HD double f64_biased_u32(uint32_t x) {  // 2^52 + x
#if defined(__CUDA_ARCH__)
    return __hiloint2double(0x43300000, (int)x);
#else
    uint64_t b = 0x4330000000000000ull | x;
    double d;
    memcpy(&d, &b, 8);
    return d;
#endif
}
JJS_HD uint64_t f64_bits(double d) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(d);
#else
    uint64_t b;
    memcpy(&b, &d, 8);
    return b;
#endif
}
#ifndef JJS_MDS_CVT
#define JJS_MDS_CVT 1   // lift the limbs with the conversion instruction (0: the bias trick 2^52 + x, two more moves per limb)
#endif
JJS_HD void mds_all_fp(fq* o, const fq* s, const double (*arkd)[8], const uint32_t* ark_top) {
    // integer Cauchy matrix 360360 / (o + k + 5), indexed by o + k
    const double N[9] = {72072.0, 60060.0, 51480.0, 45045.0, 40040.0, 36036.0, 32760.0, 30030.0, 27720.0};
    uint32_t v[5][9];
    uint64_t carry[5] = {0, 0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < 8; i++) {
        double d[5];
#pragma unroll
#if JJS_MDS_CVT
        for (int k = 0; k < 5; k++) d[k] = (double)s[k].l[i];
#else
        for (int k = 0; k < 5; k++) d[k] = f64_biased_u32(s[k].l[i]) - 4503599627370496.0;
#endif
#pragma unroll
        for (int l = 0; l < 5; l++) {
            double acc = arkd[l][i];   // 2^52 + limb i of the round constant, straight from the constant bank
#pragma unroll
            for (int k = 0; k < 5; k++) acc = fma(N[l + k], d[k], acc);
            uint64_t t = f64_bits(acc) - 0x4330000000000000ull + carry[l];   // acc is in [2^52, 2^53): its exponent field is constant
            v[l][i] = (uint32_t)t;
            carry[l] = t >> 32;
        }
    }
#pragma unroll
    for (int l = 0; l < 5; l++) {
        v[l][8] = (uint32_t)carry[l] + ark_top[l];
        redc_one(o[l].l, v[l]);
    }
}

#ifndef JJS_MDS_FP64
#define JJS_MDS_FP64 1
#endif

// In/out: ordinary Montgomery form.
JJS_HD void hades_permute(fq* s) {
#pragma unroll
    for (int i = 0; i < 5; i++) {
        fq a;
#pragma unroll
        for (int j = 0; j < 8; j++) a.l[j] = JJS_C(HADES_FIRST_ARK)[i][j];
        fq_add(s[i], s[i], a);
    }
#pragma unroll 1
    for (int rnd = 0; rnd < 68; rnd++) {
        if (rnd < 4 || rnd >= 64) {
            // Deliberately not unrolled: the run-time lane index keeps the five lanes in a 160-byte local frame (14 M L1-resident
            // local stores per 2^20 hashes), and the unrolled form without that frame measured 0.8-1.7 % SLOWER (code size;
            // gpurun_out/ab3.log).
#pragma unroll 1
            for (int i = 0; i < 5; i++) sbox5(s[i]);
        } else {
            sbox5(s[4]);
            fq fix;
#pragma unroll
            for (int j = 0; j < 8; j++) fix.l[j] = JJS_C(HADES_SBOX_FIX)[rnd - 4][j];
            fq_mul(s[4], s[4], fix);
        }
        fq o[5];
#if JJS_MDS_FP64
        mds_all_fp(o, s, JJS_C(HADES_FOLDED_ARK_D)[rnd], JJS_C(HADES_FOLDED_ARK_TOP)[rnd]);
#else
        // integer Cauchy matrix 360360 / (i + k + 5)
        mds_lane(o[0], s, JJS_C(HADES_FOLDED_ARK)[rnd][0], JJS_C(HADES_FOLDED_ARK_TOP)[rnd][0], 72072u, 60060u, 51480u, 45045u, 40040u);
        mds_lane(o[1], s, JJS_C(HADES_FOLDED_ARK)[rnd][1], JJS_C(HADES_FOLDED_ARK_TOP)[rnd][1], 60060u, 51480u, 45045u, 40040u, 36036u);
        mds_lane(o[2], s, JJS_C(HADES_FOLDED_ARK)[rnd][2], JJS_C(HADES_FOLDED_ARK_TOP)[rnd][2], 51480u, 45045u, 40040u, 36036u, 32760u);
        mds_lane(o[3], s, JJS_C(HADES_FOLDED_ARK)[rnd][3], JJS_C(HADES_FOLDED_ARK_TOP)[rnd][3], 45045u, 40040u, 36036u, 32760u, 30030u);
        mds_lane(o[4], s, JJS_C(HADES_FOLDED_ARK)[rnd][4], JJS_C(HADES_FOLDED_ARK_TOP)[rnd][4], 40040u, 36036u, 32760u, 30030u, 27720u);
#endif
#pragma unroll
        for (int i = 0; i < 5; i++) s[i] = o[i];
    }
    fq un;
#pragma unroll
    for (int j = 0; j < 8; j++) un.l[j] = JJS_C(HADES_UNSCALE)[j];
#pragma unroll 1
    for (int i = 0; i < 5; i++) fq_mul(s[i], s[i], un);
}

// SAFE sponge, IO pattern [Absorb(n), Squeeze(1)], Domain::Other
struct Sponge {
    fq s[5];
    int pos;
};
JJS_HD void sponge_start(Sponge& sp, int n_absorb) {
#pragma unroll
    for (int j = 0; j < 8; j++) sp.s[0].l[j] = JJS_C(SAFE_TAG)[n_absorb][j];
#pragma unroll
    for (int i = 1; i < 5; i++) fq_zero(sp.s[i]);
    sp.pos = 0;
}
// transcripts of run-time length (multisig: 2 + 2 n and 3 + 4 n elements): the tag comes from a table the host computes
// for the lengths a call needs (safe_tag.h), indexed by the number of absorbed elements
JJS_HD void sponge_start_tag(Sponge& sp, const fq& tag) {
    sp.s[0] = tag;
#pragma unroll
    for (int i = 1; i < 5; i++) fq_zero(sp.s[i]);
    sp.pos = 0;
}
JJS_HD void sponge_absorb(Sponge& sp, const fq& x) {  // x in Montgomery form
    if (sp.pos == 4) {
        hades_permute(sp.s);
        sp.pos = 0;
    }
    // lanes are addressed with compile-time indices to keep the state in registers
    if (sp.pos == 0) fq_add(sp.s[1], sp.s[1], x);
    else if (sp.pos == 1) fq_add(sp.s[2], sp.s[2], x);
    else if (sp.pos == 2) fq_add(sp.s[3], sp.s[3], x);
    else fq_add(sp.s[4], sp.s[4], x);
    sp.pos++;
}
// digest_truncated: canonical value of lane 1 after the final permutation, masked to 250 bits.
// Output: 8 little-endian words (a JubJub scalar < 2^250 < r).
JJS_HD void sponge_squeeze_truncated(uint32_t* c, Sponge& sp) {
    hades_permute(sp.s);
    fq canon;
    fq_from_mont(canon, sp.s[1]);
#pragma unroll
    for (int i = 0; i < 8; i++) c[i] = canon.l[i];
    c[7] &= 0x03ffffffu;
}

}  // namespace jjs
