// Field elements addressed by shared-memory slot ("handles") and the verification equation built on them, for sm_100a.
//
// Why: with field elements in registers every call of the (deliberately non-inlined) multiplier needs its operands in
// the callee's fixed registers; ptxas marshals them with IMAD.MOV, which issues on the SAME integer-multiply pipe the
// multiplier saturates (r01f profile of k_equation: 12.7 % of all executed instructions, ~24 moves around each of the
// 1 731 products of an equation).  Here a field element lives in a per-thread shared-memory slot, the multiplier is
// called with three slot handles (three integer registers) and loads its operands straight into the registers it
// wants (four LDS.128), so no field element is ever copied register to register.  The per-thread register footprint
// drops with it (the live state between calls is handles, counters and pointers), which buys resident warps.
//
// Layout: dynamic shared memory of the CTA as uint4 [slot][half][thread]; lane i of a warp touches byte 16 i of a 512-byte
// run, so every LDS.128 / STS.128 is conflict free.  A handle is the uint4 index of the thread's low half.
//
// The equation algorithm is the one of verify_core.cuh (stage_equation): half-size scalars, Straus interleave over two
// per-thread tables, 12-bit fixed-base windows; differences are noted at eq2_equation.  Host twin: the same code over a
// thread-local array (tests/hostsim), test scaffolding only.
#pragma once
#include "verify_core.cuh"

namespace jjs {

#ifndef JJS_EQ_BLOCK
#define JJS_EQ_BLOCK 128
#endif
constexpr int EQ2_SLOTS = 8;  // X Y Z T + four entry / temporary slots

#if defined(__CUDA_ARCH__)
extern __shared__ uint4 jjs_fqs_mem[];
constexpr uint32_t FQS_HALF = JJS_EQ_BLOCK, FQS_SLOT = 2 * JJS_EQ_BLOCK;
#define JJS_FQS_MEM jjs_fqs_mem
#else
constexpr uint32_t FQS_HALF = 1, FQS_SLOT = 2;
static thread_local uint4 jjs_fqs_host[2 * EQ2_SLOTS];
#define JJS_FQS_MEM jjs_fqs_host
#endif

typedef uint32_t fqh;

JJS_HD void fqs_ld(fq& r, fqh h) {
    uint4 a = JJS_FQS_MEM[h], b = JJS_FQS_MEM[h + FQS_HALF];
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
}
JJS_HD void fqs_st(fqh h, const fq& r) {
    uint4 a, b;
    a.x = r.l[0]; a.y = r.l[1]; a.z = r.l[2]; a.w = r.l[3];
    b.x = r.l[4]; b.y = r.l[5]; b.z = r.l[6]; b.w = r.l[7];
    JJS_FQS_MEM[h] = a;
    JJS_FQS_MEM[h + FQS_HALF] = b;
}

#if defined(__CUDA_ARCH__)
__device__ __noinline__ void fqs_mul_fn(fqh d, fqh a, fqh b) {
    fq x, y, r;
    fqs_ld(x, a);
    fqs_ld(y, b);
    fq_mul_inl(r, x, y);
    fqs_st(d, r);
}
__device__ __noinline__ void fqs_sqr_fn(fqh d, fqh a) {
    fq x, r;
    fqs_ld(x, a);
    fq_sqr_inl(r, x);
    fqs_st(d, r);
}
JJS_HD void fqs_mul(fqh d, fqh a, fqh b) { fqs_mul_fn(d, a, b); }
JJS_HD void fqs_sqr(fqh d, fqh a) { fqs_sqr_fn(d, a); }
#else
JJS_HD void fqs_mul(fqh d, fqh a, fqh b) {
    fq x, y, r;
    fqs_ld(x, a);
    fqs_ld(y, b);
    fq_mul_inl(r, x, y);
    fqs_st(d, r);
}
JJS_HD void fqs_sqr(fqh d, fqh a) {
    fq x, r;
    fqs_ld(x, a);
    fq_sqr_inl(r, x);
    fqs_st(d, r);
}
#endif

// The slots of one thread.  X Y Z T: the accumulator in extended coordinates.  e0..e3: the table entry being added
// (ypx, ymx, z2, t2d), reused as temporaries once its four products are taken (and throughout a doubling).
struct Eq2Slots {
    fqh X, Y, Z, T, e0, e1, e2, e3;
};
JJS_HD Eq2Slots eq2_slots(uint32_t thread_in_block) {
    Eq2Slots s;
    s.X = thread_in_block;
    s.Y = s.X + FQS_SLOT;
    s.Z = s.Y + FQS_SLOT;
    s.T = s.Z + FQS_SLOT;
    s.e0 = s.T + FQS_SLOT;
    s.e1 = s.e0 + FQS_SLOT;
    s.e2 = s.e1 + FQS_SLOT;
    s.e3 = s.e2 + FQS_SLOT;
    return s;
}

// acc = 2 acc (dbl-2008-hwcd, a = -1, on the negated quantities of ext_dbl_inl): 4 S + 3 M (+ 1 M for T)
JJS_HD void eq2_dbl_inl(const Eq2Slots& s, bool want_t) {
    fqs_sqr(s.e0, s.X);  // A
    fqs_sqr(s.e1, s.Y);  // B
    fqs_sqr(s.e2, s.Z);  // Z^2
    {
        fq x, y, t;
        fqs_ld(x, s.X);
        fqs_ld(y, s.Y);
        fq_add(t, x, y);
        fqs_st(s.e3, t);
    }
    fqs_sqr(s.e3, s.e3);  // (X + Y)^2
    {
        fq a, b, c, e, f, g, h;
        fqs_ld(a, s.e0);
        fqs_ld(b, s.e1);
        fqs_ld(c, s.e2);
        fqs_ld(e, s.e3);
        fq_dbl(c, c);
        fq_add(h, a, b);  // H'
        fq_sub(e, h, e);  // E'
        fq_sub(g, a, b);  // G'
        fq_add(f, g, c);  // F'
        fqs_st(s.e0, e);
        fqs_st(s.e1, f);
        fqs_st(s.e2, g);
        fqs_st(s.e3, h);
    }
    fqs_mul(s.X, s.e0, s.e1);
    fqs_mul(s.Y, s.e2, s.e3);
    fqs_mul(s.Z, s.e1, s.e2);
    if (want_t) fqs_mul(s.T, s.e0, s.e3);
}

// acc += (neg ? -entry : entry), entry = (e0, e1, e2, e3) = (ypx, ymx, z2, t2d) (add-2008-hwcd-3).  -(x, y) = (-x, y) swaps ypx / ymx
// and negates t2d: the swap is a choice of handles, the negation a swap of f and g.  affine: the entry has Z = 1 (z2 == 2, e2 unused).
JJS_HD void eq2_add_inl(const Eq2Slots& s, bool neg, bool affine, bool want_t) {
    {
        fq x, y, d, m;
        fqs_ld(x, s.X);
        fqs_ld(y, s.Y);
        fq_sub(d, y, x);
        fq_add(m, y, x);
        fqs_st(s.X, d);
        fqs_st(s.Y, m);
    }
    fqs_mul(s.X, s.X, neg ? s.e0 : s.e1);  // a = (Y - X) ymx
    fqs_mul(s.Y, s.Y, neg ? s.e1 : s.e0);  // b = (Y + X) ypx
    fqs_mul(s.T, s.T, s.e3);               // c (sign applied below)
    if (affine) {
        fq z;
        fqs_ld(z, s.Z);
        fq_dbl(z, z);
        fqs_st(s.Z, z);
    } else {
        fqs_mul(s.Z, s.Z, s.e2);           // d
    }
    {
        fq a, b, c, d, e, f, g, h;
        fqs_ld(a, s.X);
        fqs_ld(b, s.Y);
        fqs_ld(c, s.T);
        fqs_ld(d, s.Z);
        fq_sub(e, b, a);
        fq_add(h, b, a);
        fq_sub(f, d, c);
        fq_add(g, d, c);
        fqs_st(s.e0, e);
        fqs_st(s.e3, h);
        fqs_st(neg ? s.e2 : s.e1, f);
        fqs_st(neg ? s.e1 : s.e2, g);
    }
    // e0 = e, e1 = f, e2 = g, e3 = h
    fqs_mul(s.X, s.e0, s.e1);
    fqs_mul(s.Y, s.e2, s.e3);
    fqs_mul(s.Z, s.e1, s.e2);
    if (want_t) fqs_mul(s.T, s.e0, s.e3);
}

#if defined(__CUDA_ARCH__) && !defined(JJS_EQ2_INLINE_POINT_OPS)
__device__ __noinline__ void eq2_dbl_fn(Eq2Slots s, bool want_t) { eq2_dbl_inl(s, want_t); }
__device__ __noinline__ void eq2_add_fn(Eq2Slots s, bool neg, bool affine, bool want_t) { eq2_add_inl(s, neg, affine, want_t); }
JJS_HD void eq2_dbl(const Eq2Slots& s, bool want_t) { eq2_dbl_fn(s, want_t); }
JJS_HD void eq2_add(const Eq2Slots& s, bool neg, bool affine, bool want_t) { eq2_add_fn(s, neg, affine, want_t); }
#else
JJS_HD void eq2_dbl(const Eq2Slots& s, bool want_t) { eq2_dbl_inl(s, want_t); }
JJS_HD void eq2_add(const Eq2Slots& s, bool neg, bool affine, bool want_t) { eq2_add_inl(s, neg, affine, want_t); }
#endif

JJS_HD void eq2_set_identity(const Eq2Slots& s) {
    fq zero, one;
    fq_zero(zero);
    fq_one(one);
    fqs_st(s.X, zero);
    fqs_st(s.Y, one);
    fqs_st(s.Z, one);
    fqs_st(s.T, zero);
}

// Per-thread table of 1..8 times a variable base in projective Niels form, in global memory: `tab` points at this thread's
// first element, consecutive fq of the table are `stride` elements apart (device: the number of resident threads of the
// persistent kernel, so a warp reading one coordinate of one entry touches one contiguous kilobyte and the whole scratch is
// small enough to stay in L2; host: 1).  Entry k - 1 holds k * P; the identity is not stored (eq2_load_entry).
constexpr int EQ2_TAB_FQ = 32;  // 8 entries x 4 coordinates
JJS_HD void eq2_store_entry(const Eq2Slots& s, fq* tab, size_t stride, int k, bool keep_in_slots) {
    // to_pniels of the accumulator: (Y + X, Y - X, 2 Z, 2 d T)
    fq x, y, z, a, b, c, d2;
    fqs_ld(x, s.X);
    fqs_ld(y, s.Y);
    fqs_ld(z, s.Z);
    fq_add(a, y, x);
    fq_sub(b, y, x);
    fq_dbl(c, z);
    fq_load_const(d2, JJS_C(EDWARDS_2D));
    fqs_st(s.e0, d2);
    fqs_mul(s.e3, s.T, s.e0);
    fq* p = tab + (size_t)(k - 1) * 4 * stride;
    p[0] = a;
    p[stride] = b;
    p[2 * stride] = c;
    fq t;
    fqs_ld(t, s.e3);
    p[3 * stride] = t;
    if (keep_in_slots) {
        fqs_st(s.e0, a);
        fqs_st(s.e1, b);
        fqs_st(s.e2, c);
    }
}
// entry slots <- |digit| * P (the identity entry (1, 1, 2, 0) for digit 0, from the constant bank)
JJS_HD void eq2_load_entry(const Eq2Slots& s, const fq* tab, size_t stride, int mag) {
    fq a, b, c, d;
    if (mag == 0) {
        fq_one(a);
        b = a;
        fq_dbl(c, a);
        fq_zero(d);
    } else {
        const fq* p = tab + (size_t)(mag - 1) * 4 * stride;
        a = p[0];
        b = p[stride];
        c = p[2 * stride];
        d = p[3 * stride];
    }
    fqs_st(s.e0, a);
    fqs_st(s.e1, b);
    fqs_st(s.e2, c);
    fqs_st(s.e3, d);
}
JJS_HD void eq2_table_build(const Eq2Slots& s, fq* tab, size_t stride, const fq& u, const fq& v) {
    fq one;
    fq_one(one);
    fqs_st(s.X, u);
    fqs_st(s.Y, v);
    fqs_st(s.Z, one);
    fqs_mul(s.T, s.X, s.Y);
    eq2_store_entry(s, tab, stride, 1, false);
#pragma unroll 1
    for (int k = 2; k <= 8; k++) {
        eq2_load_entry(s, tab, stride, 1);   // the thread's own entry 1, written above
        eq2_add(s, false, false, true);
        eq2_store_entry(s, tab, stride, k, false);
    }
}

// acc = sum_i 16^i (dA[i] A + dB[i] B) over n signed radix-16 digits each; the tables are built by eq2_table_build.
// acc.T is defined on return iff want_t_last.
JJS_HD void eq2_straus2(const Eq2Slots& s, int n, const fq* tabA, const fq* tabB, size_t stride, const int8_t* dA, const int8_t* dB, bool want_t_last) {
    eq2_set_identity(s);
#pragma unroll 1
    for (int i = n - 1; i >= 0; i--) {
        if (i != n - 1) {
#pragma unroll 1
            for (int k = 0; k < 4; k++) eq2_dbl(s, k == 3);
        }
        int da = dA[i], db = dB[i];
        eq2_load_entry(s, tabA, stride, da < 0 ? -da : da);
        eq2_add(s, da < 0, false, true);
        eq2_load_entry(s, tabB, stride, db < 0 ? -db : db);
        eq2_add(s, db < 0, false, i != 0 || want_t_last);   // inside the loop the next operation is a doubling, which ignores T
    }
}

// acc += k * B for a fixed base with precomputed window tables (one mixed addition per 12-bit window); acc.T must be defined
JJS_HD void eq2_fixedbase_acc(const Eq2Slots& s, const niels* table, const uint32_t* k) {
#pragma unroll 1
    for (int w = 0; w < FB_WINDOWS; w++) {
        int bit = w * FB_W;
        uint32_t lo = k[bit >> 5] >> (bit & 31);
        if ((bit & 31) + FB_W > 32 && (bit >> 5) + 1 < 8) lo |= k[(bit >> 5) + 1] << (32 - (bit & 31));
        uint32_t idx = lo & (FB_ENTRIES - 1);
        const niels* e = table + (size_t)w * FB_ENTRIES + idx;
        fq a = e->ypx, b = e->ymx, d = e->t2d;
        fqs_st(s.e0, a);
        fqs_st(s.e1, b);
        fqs_st(s.e3, d);
        eq2_add(s, false, true, w != FB_WINDOWS - 1);
    }
}

// One verification equation, same contract as stage_equation (verify_core.cuh).  Differences in the evaluation order only:
// the Straus part runs first and the 21 fixed-base windows are then added onto the same accumulator (no second accumulator,
// no conversion of the fixed-base sum), and the second addition of every window skips its T product.
JJS_HD bool eq2_equation(const Eq2Slots& s, const fq* pts_u, const fq* pts_v, size_t n, size_t item, int pk_slot, int r_slot, int base_slot,
                         const niels* fb, const WireField& usc, const uint32_t* c_words, fq* tabA, fq* tabB, size_t stride, bool* r_implied) {
    uint32_t u[8], c[8];
    wire_load(u, usc, item);
#pragma unroll
    for (int i = 0; i < 8; i++) c[i] = c_words[item * 8 + i];
    if (base_slot < 0) {
        uint32_t tau[5], rho[8];
        bool rho_neg, rho_odd;
        half_gcd(tau, rho, rho_neg, rho_odd, c);
#pragma unroll
        for (int i = 4; i < 8; i++) rho[i] = 0;
        int8_t dT[33], dR[33];
        recode_signed16_33(dT, tau, rho_neg);
        recode_signed16_n<4>(dR, rho, true);
        eq2_table_build(s, tabA, stride, pts_u[pk_slot * n + item], pts_v[pk_slot * n + item]);
        eq2_table_build(s, tabB, stride, pts_u[r_slot * n + item], pts_v[r_slot * n + item]);
        eq2_straus2(s, 33, tabA, tabB, stride, dT, dR, true);
        uint32_t ru[8];
        fr_mul_short(ru, rho, u);
        eq2_fixedbase_acc(s, fb, ru);
        fq x, y, z;
        fqs_ld(x, s.X);
        fqs_ld(y, s.Y);
        fqs_ld(z, s.Z);
        bool ok = fq_is_zero(x) && fq_eq(y, z);
        *r_implied = ok && rho_odd;
        return ok;
    }
    int8_t dU[64], dC[64];
    recode_signed16(dU, u);
    recode_signed16(dC, c);
    eq2_table_build(s, tabA, stride, pts_u[base_slot * n + item], pts_v[base_slot * n + item]);
    eq2_table_build(s, tabB, stride, pts_u[pk_slot * n + item], pts_v[pk_slot * n + item]);
    eq2_straus2(s, 64, tabA, tabB, stride, dU, dC, false);
    // projective comparison with R: X == u_R Z and Y == v_R Z
    fqs_st(s.e0, pts_u[r_slot * n + item]);
    fqs_st(s.e1, pts_v[r_slot * n + item]);
    fqs_mul(s.e0, s.e0, s.Z);
    fqs_mul(s.e1, s.e1, s.Z);
    fq a, b, x, y;
    fqs_ld(a, s.e0);
    fqs_ld(b, s.e1);
    fqs_ld(x, s.X);
    fqs_ld(y, s.Y);
    bool ok = fq_eq(a, x) && fq_eq(b, y);
    *r_implied = ok;
    return ok;
}

// stage_equation_item (verify_core.cuh) on the slot-based evaluation: same flag handling, same results
JJS_HD bool eq2_equation_item(const Eq2Slots& s, int variant, int eq, const fq* pts_u, const fq* pts_v, uint8_t* pflags, size_t n, size_t item,
                              const niels* fb, const WireField& usc, const uint32_t* c_words, fq* tabA, fq* tabB, size_t stride, bool* need_r_test) {
    int pk_slot, r_slot, base_slot;
    equation_slots(variant, eq, pk_slot, r_slot, base_slot);
    *need_r_test = false;
    if (!point_flags_valid(pflags[pk_slot * n + item])) return false;
    if (base_slot >= 0 && !point_flags_valid(pflags[base_slot * n + item])) return false;
    bool implied = false;
    bool ok = eq2_equation(s, pts_u, pts_v, n, item, pk_slot, r_slot, base_slot, fb, usc, c_words, tabA, tabB, stride, &implied);
    uint8_t rf = pflags[r_slot * n + item];
    if (rf & PF_TORSION_PENDING) {
        if (implied) pflags[r_slot * n + item] = (uint8_t)((rf & ~PF_TORSION_PENDING) | PF_TORSION_FREE);
        else *need_r_test = true;
    }
    return ok;
}

}  // namespace jjs
