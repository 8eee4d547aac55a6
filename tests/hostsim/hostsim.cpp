// Host build of the device headers (their C++ twins of the PTX blocks) so the algorithms can be checked
// on a CPU-only machine.  Test scaffolding: never linked into libjjschnorr_b200.so.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../jubjub_schnorr_b200/csrc/multisig_core.cuh"
#include "../../jubjub_schnorr_b200/csrc/sign_core.cuh"
#include "../../jubjub_schnorr_b200/csrc/verify_core.cuh"
#include "../../jubjub_schnorr_b200/csrc/fqs.cuh"
#include "../../tools/fq_fp.cuh"
#include "../../jubjub_schnorr_b200/csrc/safe_tag.h"
namespace tables {
#include "../../jubjub_schnorr_b200/csrc/jjs_constants_tables.h"
}

using namespace jjs;

static std::vector<niels> g_fb_g, g_fb_gn;
static Tables g_tables;
static bool g_ready = false;

static void build_fb(std::vector<niels>& out, const uint32_t uv[2][8]) {
    fq u, v;
    memcpy(u.l, uv[0], 32);
    memcpy(v.l, uv[1], 32);
    out.resize((size_t)FB_WINDOWS * FB_ENTRIES);
    // incremental construction with one batched inversion per window (fb_table_entry itself is exercised by hs_fb_entry)
    ext base;
    ext_from_affine(base, u, v);
    std::vector<ext> pts(FB_ENTRIES);
    std::vector<fq> prefix(FB_ENTRIES);
    fq d2;
    fq_load_const(d2, JJS_C(EDWARDS_2D));
    for (int w = 0; w < FB_WINDOWS; w++) {
        pniels nb;
        ext_to_pniels(nb, base);
        ext acc;
        ext_identity(acc);
        for (int j = 0; j < FB_ENTRIES; j++) {
            pts[j] = acc;
            ext t;
            ext_add_pniels<true>(t, acc, nb);
            acc = t;
        }
        base = acc;  // 2^FB_W * previous base
        fq run;
        fq_one(run);
        for (int j = 0; j < FB_ENTRIES; j++) {
            prefix[j] = run;
            fq_mul(run, run, pts[j].Z);
        }
        fq inv_all;
        fq_inv(inv_all, run);
        for (int j = FB_ENTRIES - 1; j >= 0; j--) {
            fq zi, au, av;
            fq_mul(zi, inv_all, prefix[j]);
            fq_mul(inv_all, inv_all, pts[j].Z);
            fq_mul(au, pts[j].X, zi);
            fq_mul(av, pts[j].Y, zi);
            niels& e = out[(size_t)w * FB_ENTRIES + j];
            fq_add(e.ypx, av, au);
            fq_sub(e.ymx, av, au);
            fq_mul(e.t2d, au, av);
            fq_mul(e.t2d, e.t2d, d2);
        }
    }
}

// SAFE tags for transcripts of run-time length, as the library's host side computes them (safe_tag.h)
static std::vector<fq> g_tags;
static const fq* tags_upto(size_t n_absorb) {
    while (g_tags.size() <= n_absorb) {
        fq t;
        safe_tag_mont(t.l, (uint32_t)g_tags.size());
        g_tags.push_back(t);
    }
    return g_tags.data();
}

static void ensure_ready() {
    if (g_ready) return;
    g_tables.root_tables = reinterpret_cast<const fq*>(tables::ROOT_TABLES);
    g_tables.dlog_hash = tables::DLOG_HASH;
    build_fb(g_fb_g, hconsts::GEN_UV);
    build_fb(g_fb_gn, hconsts::GEN_NUMS_UV);
    g_tables.fb_g = g_fb_g.data();
    g_tables.fb_gn = g_fb_gn.data();
    g_tables.safe_tags = nullptr;   // the twins pass the tag table explicitly (tags_upto)
    g_ready = true;
}

// equation stage as the kernels run it: k_equation (stage_equation_item) followed by k_rtest for the points it queued
static int g_rtests = 0, g_equations = 0;
static int g_eq_impl = 1;  // 1: register-operand evaluation (stage_equation_item), what k_equation runs; 2: slot-based evaluation (fqs.cuh, the -DJJS_EQ_V2 build)
static bool equation_with_rtest(int variant, int eq, const fq* pu, const fq* pv, uint8_t* pf, size_t n, size_t i, const WireField& fu, const uint32_t* cw,
                                fq* tab) {
    bool need;
    const niels* fb = (variant == VAR_DOUBLE && eq == 1) ? g_tables.fb_gn : g_tables.fb_g;
    bool ok = g_eq_impl == 2 ? eq2_equation_item(eq2_slots(0), variant, eq, pu, pv, pf, n, i, fb, fu, cw, tab, tab + 36, 1, &need)
                             : stage_equation_item(variant, eq, pu, pv, pf, n, i, fb, fu, cw, tab, tab + 36, 1, &need);
    g_equations++;
    if (need) {
        int pk_slot, r_slot, base_slot;
        equation_slots(variant, eq, pk_slot, r_slot, base_slot);
        stage_rtest(pu, pv, pf, (size_t)r_slot * n + i);
        g_rtests++;
    }
    return ok;
}

extern "C" {
void hs_set_equation_impl(int impl) { g_eq_impl = impl; }
void hs_force_vargen_fallback(int on) { lattice3_host_force_none() = on != 0; }
void hs_safe_tag(uint32_t n_absorb, uint32_t* out8) { safe_tag_mont(out8, n_absorb); }
// counters of the deferred subgroup tests since the last call (equations evaluated, tests run)
void hs_rtest_counters(int* equations, int* rtests) { *equations = g_equations; *rtests = g_rtests; g_equations = g_rtests = 0; }
void hs_mul_wide(const uint32_t* a, const uint32_t* b, uint32_t* t16) { mul_wide(t16, a, b); }
void hs_sqr_wide(const uint32_t* a, uint32_t* t16) { sqr_wide(t16, a); }
void hs_redc(const uint32_t* t16, uint32_t* r8) { redc(r8, t16); }
void hs_fq_mul(const uint32_t* a, const uint32_t* b, uint32_t* r) { fq x, y, z; memcpy(x.l, a, 32); memcpy(y.l, b, 32); fq_mul(z, x, y); memcpy(r, z.l, 32); }
void hs_fq_mul_fp(const uint32_t* a, const uint32_t* b, uint32_t* r) { fq x, y, z; memcpy(x.l, a, 32); memcpy(y.l, b, 32); fq_mul_fp_inl(z, x, y); memcpy(r, z.l, 32); }
void hs_fr_mul_short(const uint32_t* a4, const uint32_t* b8, uint32_t* r) { fr_mul_short(r, a4, b8); }
void hs_fr_mul_160(const uint32_t* a5, const uint32_t* b8, uint32_t* r) { fr_mul_160(r, a5, b8); }
void hs_fr_mul(const uint32_t* a8, const uint32_t* b8, uint32_t* r) { fr_mul(r, a8, b8); }
void hs_fq_sqr(const uint32_t* a, uint32_t* r) { fq x, z; memcpy(x.l, a, 32); fq_sqr(z, x); memcpy(r, z.l, 32); }
void hs_fq_add(const uint32_t* a, const uint32_t* b, uint32_t* r) { fq x, y, z; memcpy(x.l, a, 32); memcpy(y.l, b, 32); fq_add(z, x, y); memcpy(r, z.l, 32); }
void hs_fq_sub(const uint32_t* a, const uint32_t* b, uint32_t* r) { fq x, y, z; memcpy(x.l, a, 32); memcpy(y.l, b, 32); fq_sub(z, x, y); memcpy(r, z.l, 32); }
void hs_fq_to_mont(const uint32_t* a, uint32_t* r) { fq x, z; memcpy(x.l, a, 32); fq_to_mont(z, x); memcpy(r, z.l, 32); }
void hs_fq_from_mont(const uint32_t* a, uint32_t* r) { fq x, z; memcpy(x.l, a, 32); fq_from_mont(z, x); memcpy(r, z.l, 32); }
void hs_fq_inv(const uint32_t* a, uint32_t* r) { fq x, z; memcpy(x.l, a, 32); fq_inv(z, x); memcpy(r, z.l, 32); }
// sqrt(num/den), Montgomery in/out; returns 0 for non-residue
int hs_sqrt_ratio(const uint32_t* num, const uint32_t* den, uint32_t* r) {
    ensure_ready();
    fq x, y, z;
    memcpy(x.l, num, 32); memcpy(y.l, den, 32);
    bool ok = fq_sqrt_ratio(z, x, y, g_tables);
    memcpy(r, z.l, 32);
    return ok;
}
// canonical 5-lane state in/out
void hs_hades_permute(uint32_t* state40) {
    fq s[5];
    for (int i = 0; i < 5; i++) { fq c; memcpy(c.l, state40 + 8 * i, 32); fq_to_mont(s[i], c); }
    hades_permute(s);
    for (int i = 0; i < 5; i++) { fq c; fq_from_mont(c, s[i]); memcpy(state40 + 8 * i, c.l, 32); }
}
int hs_point_decode(const uint8_t* in32, uint32_t* uv16) {  // canonical (u, v) out
    ensure_ready();
    uint32_t w[8];
    memcpy(w, in32, 32);
    fq u, v, cu, cv;
    if (!point_from_wire(u, v, w, g_tables)) return 0;
    fq_from_mont(cu, u); fq_from_mont(cv, v);
    memcpy(uv16, cu.l, 32); memcpy(uv16 + 8, cv.l, 32);
    return 1;
}
void hs_fb_entry(int which, int w, int j, uint32_t* out24, uint32_t* ref24) {  // fb_table_entry vs incremental table
    ensure_ready();
    fq u, v;
    const uint32_t (*uv)[8] = which ? hconsts::GEN_NUMS_UV : hconsts::GEN_UV;
    memcpy(u.l, uv[0], 32); memcpy(v.l, uv[1], 32);
    niels e;
    fb_table_entry(e, u, v, w, j);
    memcpy(out24, &e, 96);
    memcpy(ref24, &(which ? g_fb_gn : g_fb_g)[(size_t)w * FB_ENTRIES + j], 96);
}
// k * P for a compressed point, k a 256-bit LE scalar < 2^253 ; returns compressed result (uses varbase path)
int hs_varbase_mul(const uint8_t* p32, const uint8_t* k32, uint8_t* out32, int fixed_base /*0 none, 1 G, 2 G'*/) {
    ensure_ready();
    uint32_t w[8], k[8];
    memcpy(k, k32, 32);
    ext r;
    if (fixed_base) {
        fixedbase_mul(r, fixed_base == 1 ? g_tables.fb_g : g_tables.fb_gn, k);
    } else {
        memcpy(w, p32, 32);
        fq u, v;
        if (!point_from_wire(u, v, w, g_tables)) return 0;
        std::vector<fq> tab(36);
        varbase_table_build(tab.data(), 1, u, v);
        int8_t digits[64];
        recode_signed16(digits, k);
        varbase_mul<true>(r, tab.data(), 1, digits);
    }
    fq zi, au, av;
    fq_inv(zi, r.Z);
    fq_mul(au, r.X, zi);
    fq_mul(av, r.Y, zi);
    uint32_t o[8];
    point_to_wire(o, au, av);
    memcpy(out32, o, 32);
    return 1;
}
// aggregate-key verification through the stage functions (ragged keys)
void hs_verify_aggregate(const uint8_t* pks, const uint32_t* offsets, const uint8_t* sig, const uint8_t* msg, size_t n, uint8_t* status,
                         uint8_t* c_out, uint8_t* agg_out) {
    ensure_ready();
    size_t K = offsets[n];
    std::vector<fq> ku(K), kv(K), pu(2 * n), pv(2 * n), tab(R32_TAB_FQ * AGG_GROUP);
    std::vector<uint8_t> kf(K), pf(2 * n), itf(n);
    std::vector<uint32_t> cw(8 * n), kc(8 * K + 8);
    WireField fk{pks, 32}, fR{sig + 32, 64}, fmsg{msg, 32}, fu{sig, 64};
    for (size_t k = 0; k < K; k++) stage_decode(fk, k, ku.data(), kv.data(), kf.data(), k, g_tables, false);
    size_t max_cnt = 0;
    for (size_t i = 0; i < n; i++) max_cnt = std::max<size_t>(max_cnt, offsets[i + 1] - offsets[i]);
    const fq* tags = tags_upto(2 + 2 * max_cnt);
    for (size_t i = 0; i < n; i++) {
        uint32_t w[8];
        // the per-key form is what k_agg_coeffs runs; odd items go through the per-item form so both stay covered
        if (i & 1) stage_aggregate_coeffs(ku.data(), kv.data(), kf.data(), offsets[i], offsets[i + 1], kc.data(), tags);
        else
            for (uint32_t j = offsets[i]; j < offsets[i + 1]; j++)
                stage_aggregate_coeff_key(ku.data(), kv.data(), kf.data(), offsets[i], offsets[i + 1], j, kc.data(), tags);
        stage_aggregate(ku.data(), kv.data(), kf.data(), offsets[i], offsets[i + 1], pu.data(), pv.data(), pf.data(), i, w, tab.data(), 1, tags, kc.data());
        memcpy(agg_out + 32 * i, w, 32);
        stage_decode(fR, i, pu.data(), pv.data(), pf.data(), n + i, g_tables, false);
        bool all = stage_challenge(VAR_SINGLE, pu.data(), pv.data(), pf.data(), n, i, fmsg, fu, cw.data(), itf.data());
        if (all && equation_with_rtest(VAR_SINGLE, 0, pu.data(), pv.data(), pf.data(), n, i, fu, cw.data(), tab.data())) itf[i] |= IF_EQ0_OK;
        status[i] = stage_status(VAR_SINGLE, pf.data(), itf[i], n, i);
        if (status[i] <= 1) memcpy(c_out + 32 * i, &cw[8 * i], 32); else memset(c_out + 32 * i, 0, 32);
    }
}
// half-size decomposition: outputs tau (20 bytes), |rho| (16 bytes), returns sign of rho (bit 0: negative) and bit 1: rho odd;
// digits33 (optional, 2 x 33 signed bytes): the radix-16 digits the equation kernel uses for sign(rho) tau and -|rho|
int hs_half_gcd(const uint8_t* c32, uint8_t* tau20, uint8_t* rho16, int8_t* digits33) {
    uint32_t c[8], tau[5], rho[4];
    memcpy(c, c32, 32);
    bool neg, odd;
    half_gcd(tau, rho, neg, odd, c);
    memcpy(tau20, tau, 20);
    memcpy(rho16, rho, 16);
    if (digits33) {
        recode_signed16_33(digits33, tau, neg);
        recode_signed16_n<4>(digits33 + 33, rho, true);
    }
    return (neg ? 1 : 0) | (odd ? 2 : 0);
}
// three short scalars of the var-generator equation: magnitudes (32 bytes each) and flags (bit 0/1/2: x/y/z negative, bit 3: z odd);
// returns 0 if no vector fits the 43 windows; digits (optional, 3 x 64 signed bytes): what the equation kernel uses for x, y, -z
int hs_lattice3(const uint8_t* u32, const uint8_t* c32, uint8_t* x32, uint8_t* y32, uint8_t* z32, int* flags, int8_t* digits) {
    uint32_t u[8], c[8], xm[8], ym[8], zm[8];
    memcpy(u, u32, 32);
    memcpy(c, c32, 32);
    bool xn, yn, zn, odd;
    if (!lattice3_reduce(xm, ym, zm, xn, yn, zn, odd, u, c)) return 0;
    memcpy(x32, xm, 32);
    memcpy(y32, ym, 32);
    memcpy(z32, zm, 32);
    *flags = (xn ? 1 : 0) | (yn ? 2 : 0) | (zn ? 4 : 0) | (odd ? 8 : 0);
    if (digits) {
        memset(digits, 0, 192);
        recode_signed16_n<6>(digits, xm, xn);
        recode_signed16_n<6>(digits + 64, ym, yn);
        recode_signed16_n<6>(digits + 128, zm, !zn);
    }
    return 1;
}
int hs_subgroup(const uint8_t* p32, int method) {
    ensure_ready();
    std::vector<fq> tab(36);
    return subgroup_check(WireField{p32, 32}, 0, method, tab.data(), 1, g_tables);
}
// typed inputs through the stage functions (single variant): pts = n x 2 x 160 bytes
void hs_verify_ext(int variant, const uint8_t* pts, const uint8_t* u32, const uint8_t* msg, size_t n, uint8_t* status, uint8_t* c_out) {
    ensure_ready();
    const int slots = variant_slots(variant);
    std::vector<fq> pu(slots * n), pv(slots * n), tab(108);
    std::vector<uint8_t> pf(slots * n), itf(n);
    std::vector<uint32_t> cw(8 * n);
    WireField fmsg{msg, 32}, fu{u32, 32};
    for (int s = 0; s < slots; s++)
        for (size_t i = 0; i < n; i++)
            stage_decode_ext(WireField{pts + 160 * s, (uint32_t)(160 * slots)}, i, pu.data(), pv.data(), pf.data(), s * n + i,
                             s < (variant == VAR_SINGLE ? 1 : 2));
    for (size_t i = 0; i < n; i++) {
        bool all = stage_challenge(variant, pu.data(), pv.data(), pf.data(), n, i, fmsg, fu, cw.data(), itf.data());
        if (all) {
            for (int eq = 0; eq < (variant == VAR_DOUBLE ? 2 : 1); eq++)
                if (equation_with_rtest(variant, eq, pu.data(), pv.data(), pf.data(), n, i, fu, cw.data(), tab.data())) itf[i] |= eq ? IF_EQ1_OK : IF_EQ0_OK;
        }
        status[i] = stage_status(variant, pf.data(), itf[i], n, i);
        if (status[i] <= 1) memcpy(c_out + 32 * i, &cw[8 * i], 32); else memset(c_out + 32 * i, 0, 32);
    }
}
// multisig::combine through the stage functions
void hs_multisig_combine(const uint8_t* pks, const uint8_t* Rs, const uint8_t* Ss, const uint8_t* zs, const uint32_t* offsets, const uint8_t* msg, size_t n,
                         uint8_t* share_ok, uint8_t* status, uint32_t* bad, uint8_t* sig) {
    ensure_ready();
    size_t K = offsets[n];
    std::vector<fq> pu(3 * K + 1), pv(3 * K + 1), tab(R32_TAB_FQ * AGG_GROUP), ru(n), rv(n);
    std::vector<uint8_t> pf(3 * K + 1), sf(n);
    std::vector<uint32_t> dw(8 * K + 8), cdw(8 * K + 8), aw(8 * n);
    WireField f[3] = {{pks, 32}, {Rs, 32}, {Ss, 32}}, fmsg{msg, 32}, fz{zs, 32};
    for (int s = 0; s < 3; s++)
        for (size_t j = 0; j < K; j++) stage_decode(f[s], j, pu.data(), pv.data(), pf.data(), s * K + j, g_tables, false);
    size_t max_cnt = 0;
    for (size_t s = 0; s < n; s++) max_cnt = std::max<size_t>(max_cnt, offsets[s + 1] - offsets[s]);
    const fq* tags = tags_upto(3 + 4 * max_cnt);
    for (size_t s = 0; s < n; s++)
        stage_msig_session(pu.data(), pv.data(), pf.data(), K, offsets[s], offsets[s + 1], fmsg, fz, s, dw.data(), cdw.data(), aw.data(), ru.data(), rv.data(),
                           sf.data(), tab.data(), 1, tags);
    for (size_t s = 0; s < n; s++)
        for (uint32_t j = offsets[s]; j < offsets[s + 1]; j++)
            share_ok[j] = sf[s] == (SF_DECODED | SF_NONEMPTY) &&
                          stage_msig_share(pu.data(), pv.data(), K, j, fz, cdw.data(), aw.data() + 8 * s, g_tables.fb_g, tab.data(), tab.data() + 36, 1);
    for (size_t s = 0; s < n; s++) {
        uint32_t sg[16];
        status[s] = stage_msig_finalize(sf.data(), share_ok, offsets[s], offsets[s + 1], s, fz, ru.data(), rv.data(), &bad[s], sg);
        memcpy(sig + 64 * s, sg, 64);
    }
}
int hs_sign(int variant, const uint8_t* sk, const uint8_t* rnd, const uint8_t* gsc, const uint8_t* msg, uint8_t* pk_out, uint8_t* sig_out) {
    ensure_ready();
    uint32_t a[8], b[8], g[8] = {0}, m[8], pk[16], sig[24];
    memcpy(a, sk, 32); memcpy(b, rnd, 32); memcpy(m, msg, 32);
    if (gsc) memcpy(g, gsc, 32);
    if (!sign_item(variant, a, b, g, m, pk, sig, g_tables)) return 0;
    memcpy(pk_out, pk, variant == VAR_SINGLE ? 32 : 64);
    memcpy(sig_out, sig, variant == VAR_DOUBLE ? 96 : 64);
    return 1;
}
// Whole pipeline through the same stage functions the kernels call.
void hs_verify(int variant, const uint8_t* pk, const uint8_t* sig, const uint8_t* msg, size_t n, uint8_t* status, uint8_t* c_out) {
    ensure_ready();
    const int slots = variant_slots(variant);
    uint32_t pk_stride = variant == VAR_SINGLE ? 32 : 64, sig_stride = variant == VAR_DOUBLE ? 96 : 64;
    WireField f[4], fmsg{msg, 32}, fu{sig, sig_stride};
    if (variant == VAR_SINGLE) { f[0] = {pk, 32}; f[1] = {sig + 32, 64}; }
    else if (variant == VAR_DOUBLE) { f[0] = {pk, 64}; f[1] = {pk + 32, 64}; f[2] = {sig + 32, 96}; f[3] = {sig + 64, 96}; }
    else { f[0] = {pk, 64}; f[1] = {pk + 32, 64}; f[2] = {sig + 32, 64}; }
    (void)pk_stride;
    std::vector<fq> pu(slots * n), pv(slots * n), tab(108);
    std::vector<uint8_t> pf(slots * n), itf(n);
    std::vector<uint32_t> cw(8 * n);
    for (int s = 0; s < slots; s++)
        for (size_t i = 0; i < n; i++) stage_decode(f[s], i, pu.data(), pv.data(), pf.data(), s * n + i, g_tables, s < (variant == VAR_SINGLE ? 1 : 2));
    std::vector<uint8_t> ready(n);
    for (size_t i = 0; i < n; i++) ready[i] = stage_challenge(variant, pu.data(), pv.data(), pf.data(), n, i, fmsg, fu, cw.data(), itf.data());
    for (size_t i = 0; i < n; i++) {
        bool all = ready[i];
        if (all) {
            for (int eq = 0; eq < (variant == VAR_DOUBLE ? 2 : 1); eq++)
                if (equation_with_rtest(variant, eq, pu.data(), pv.data(), pf.data(), n, i, fu, cw.data(), tab.data())) itf[i] |= eq ? IF_EQ1_OK : IF_EQ0_OK;
        }
        status[i] = stage_status(variant, pf.data(), itf[i], n, i);
        if (status[i] <= 1) memcpy(c_out + 32 * i, &cw[8 * i], 32); else memset(c_out + 32 * i, 0, 32);
    }
}
}
