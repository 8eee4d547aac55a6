// CUDA kernels (sm_100a) and the C ABI of libjjschnorr_b200.so.  See include/jjschnorr_b200.h for the
// contract and verify_core.cuh for the per-item stages.  One thread owns one point (decode, subgroup
// check), one item (challenge sponge, final status) or one verification equation (Straus-style
// u*B + c*PK); intermediates travel between the stages as SoA arrays of 32-byte field elements in HBM,
// so every warp access is a run of 128-bit vector loads.  There is deliberately no CPU path.
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <vector>

#include "../../include/jjschnorr_b200.h"
#include "multisig_core.cuh"
#include "sign_core.cuh"
#include "verify_core.cuh"

namespace tables {
#include "jjs_constants_tables.h"
}

using namespace jjs;

namespace {

#ifndef JJS_BLOCK
#define JJS_BLOCK 128
#endif
constexpr int BLOCK = JJS_BLOCK;
#ifndef JJS_EQ_MINBLOCKS
#define JJS_EQ_MINBLOCKS 3
#endif
#ifndef JJS_DEC_MINBLOCKS
#define JJS_DEC_MINBLOCKS 4
#endif
constexpr size_t CHUNK_ITEMS = size_t(1) << 20;   // items per pipeline pass
constexpr size_t TAB_THREADS = size_t(1) << 20;   // threads served by the per-thread table scratch (four tables of 1152 B each)

struct Fields {
    WireField f[4];
};

// thread t < slots * n decodes point field t / n of item t % n into index (slot_base + t / n) * n + t % n;
// bit s of subgroup_mask: run the subgroup test for slot s now (keys), else leave it to the equation stage (signature points)
__global__ void __launch_bounds__(BLOCK, JJS_DEC_MINBLOCKS) k_decode(Fields fields, int slots, int slot_base, size_t n, fq* pts_u, fq* pts_v,
                                                                      uint8_t* pflags, Tables T, uint32_t subgroup_mask) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)slots * n) return;
    int slot = (int)(t / n);
    size_t item = t - (size_t)slot * n;
    stage_decode(fields.f[slot], item, pts_u, pts_v, pflags, (size_t)(slot_base + slot) * n + item, T, (subgroup_mask >> slot) & 1u);
}

// typed inputs: thread t < slots * n normalises point t % slots of item t / slots (item-major, 160 bytes per point)
__global__ void __launch_bounds__(BLOCK, JJS_DEC_MINBLOCKS) k_decode_ext(const uint8_t* pts, int slots, size_t n, fq* pts_u, fq* pts_v, uint8_t* pflags,
                                                                          uint32_t subgroup_mask) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)slots * n) return;
    int slot = (int)(t / n);
    size_t item = t - (size_t)slot * n;
    stage_decode_ext(WireField{pts + 160 * (size_t)slot, (uint32_t)(160 * slots)}, item, pts_u, pts_v, pflags, (size_t)slot * n + item,
                     (subgroup_mask >> slot) & 1u);
}

// test-data utility: wire point -> JubJubExtended coordinates (u z, v z, z, u z, v) for a caller-chosen Montgomery z
__global__ void __launch_bounds__(BLOCK, JJS_DEC_MINBLOCKS) k_points_to_ext(const uint8_t* pts32, const uint8_t* z32, size_t n, uint8_t* out160, Tables T) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t w[8];
    fq u, v, z, c[5];
    wire_load(w, WireField{pts32, 32}, i);
    bool ok = point_from_wire(u, v, w, T);
    wire_load(z.l, WireField{z32, 32}, i);
    ok = ok && !ge_q(z.l);
    fq_mul(c[0], u, z);
    fq_mul(c[1], v, z);
    c[2] = z;
    c[3] = c[0];
    c[4] = v;
    uint4* o = reinterpret_cast<uint4*>(out160 + 160 * i);
#pragma unroll
    for (int k = 0; k < 5; k++) {
        o[2 * k] = ok ? make_uint4(c[k].l[0], c[k].l[1], c[k].l[2], c[k].l[3]) : make_uint4(0, 0, 0, 0);
        o[2 * k + 1] = ok ? make_uint4(c[k].l[4], c[k].l[5], c[k].l[6], c[k].l[7]) : make_uint4(0, 0, 0, 0);
    }
}

// one thread hashes the delinearisation coefficients of one item's signer keys (same item order as k_aggregate)
__global__ void __launch_bounds__(BLOCK, JJS_DEC_MINBLOCKS) k_agg_coeffs(const fq* keys_u, const fq* keys_v, const uint8_t* kflags,
                                                                          const uint32_t* offsets, const uint32_t* order, uint32_t key_base, size_t n,
                                                                          uint32_t* d_words) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    size_t item = order[t];
    stage_aggregate_coeffs(keys_u, keys_v, kflags, offsets[item] - key_base, offsets[item + 1] - key_base, d_words);
}

// one thread folds the signer keys of one item into its aggregate key (slot 0 of the single-variant point arrays)
// `order` lists the items sorted by signer count, so the lanes of a warp loop over the same number of signers
__global__ void __launch_bounds__(BLOCK) k_aggregate(const fq* keys_u, const fq* keys_v, const uint8_t* kflags, const uint32_t* offsets,
                                                     const uint32_t* order, uint32_t key_base, size_t n, fq* pts_u, fq* pts_v, uint8_t* pflags,
                                                     uint8_t* agg_out, fq* tab, size_t stride, uint32_t* d_words) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    size_t item = order[t];
    uint32_t w[8];
    stage_aggregate(keys_u, keys_v, kflags, offsets[item] - key_base, offsets[item + 1] - key_base, pts_u, pts_v, pflags, item, w, tab + t, stride,
                    d_words);
    if (agg_out) {
        uint4* o = reinterpret_cast<uint4*>(agg_out + item * 32);
        o[0] = make_uint4(w[0], w[1], w[2], w[3]);
        o[1] = make_uint4(w[4], w[5], w[6], w[7]);
    }
}

// Work list of the hash and equation stages: the items that go on after decoding (verify_core.cuh, stage_item_ready).
// The others are settled by their flags, and leaving them out keeps every lane of the two heavy kernels busy on
// batches with many invalid items.  One warp ballot and one atomic per warp; also writes the scalar-range flag of every
// item, clears its equation results and zeroes the challenge words of the items that are left out.
__global__ void __launch_bounds__(BLOCK) k_work_list(int variant, const uint8_t* pflags, size_t n, WireField msg, WireField usc, bool require_valid_keys,
                                                     uint8_t* iflags, uint8_t* eqflags, uint32_t* cwords, uint32_t* list, uint32_t* count) {
    size_t item = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool ready = false;
    if (item < n) {
        bool ok = stage_scalars_ok(msg, usc, item);
        iflags[item] = ok ? IF_SCALARS_OK : 0;
        ready = stage_item_ready(variant, pflags, n, item, ok, require_valid_keys);
        eqflags[item] = 0;
        if (variant == VAR_DOUBLE) eqflags[n + item] = 0;
        if (!ready) {
            uint4* c = reinterpret_cast<uint4*>(cwords + item * 8);
            c[0] = make_uint4(0, 0, 0, 0);
            c[1] = make_uint4(0, 0, 0, 0);
        }
    }
    uint32_t mask = __ballot_sync(0xffffffffu, ready);
    if (mask == 0) return;
    int lane = threadIdx.x & 31, leader = __ffs(mask) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(count, (uint32_t)__popc(mask));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (ready) list[base + __popc(mask & ((1u << lane) - 1u))] = (uint32_t)item;
}

// thread t < *count hashes the challenge of item list[t]
__global__ void __launch_bounds__(BLOCK, JJS_DEC_MINBLOCKS) k_challenge(int variant, const fq* pts_u, const fq* pts_v, size_t n, WireField msg,
                                                                         const uint32_t* list, const uint32_t* count, uint32_t* cwords) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= *count) return;
    stage_challenge_hash(variant, pts_u, pts_v, n, list[t], msg, cwords);
}

__global__ void __launch_bounds__(BLOCK) k_subgroup_check(WireField pts, size_t first, size_t count, int method, uint8_t* out, fq* tab,
                                                          size_t stride, Tables T) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    out[first + t] = subgroup_check(pts, first + t, method, tab + t, stride, T);
}

// thread T = first + t < neq * *count : equation T % neq of item list[T / neq].  Signature points whose subgroup
// membership the equation did not establish are appended to `rlist` (indices into the point arrays) for k_rtest.
__global__ void __launch_bounds__(BLOCK, JJS_EQ_MINBLOCKS) k_equation(int variant, const fq* pts_u, const fq* pts_v, uint8_t* pflags, size_t n,
                                                    size_t first, size_t count, const uint32_t* list, const uint32_t* lcount, WireField usc,
                                                    const uint32_t* cwords, uint8_t* eqflags, fq* tab, size_t stride, Tables T,
                                                    uint32_t* rlist, uint32_t* rcount) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int neq = variant == VAR_DOUBLE ? 2 : 1;
    size_t g = first + t;
    if (t >= count || g >= (size_t)neq * *lcount) return;
    int eq = (int)(g % neq);
    size_t item = list[g / neq];
    bool need_r_test;
    bool ok = stage_equation_item(variant, eq, pts_u, pts_v, pflags, n, item, (variant == VAR_DOUBLE && eq == 1) ? T.fb_gn : T.fb_g, usc, cwords,
                                  tab + t, tab + 36 * stride + t, stride, &need_r_test);
    if (need_r_test) {
        int pk_slot, r_slot, base_slot;
        equation_slots(variant, eq, pk_slot, r_slot, base_slot);
        rlist[atomicAdd(rcount, 1u)] = (uint32_t)((size_t)r_slot * n + item);
    }
    eqflags[(size_t)eq * n + item] = ok ? 1 : 0;
}

// deferred subgroup tests: thread t < *rcount tests point rlist[t]
__global__ void __launch_bounds__(BLOCK, JJS_DEC_MINBLOCKS) k_rtest(const fq* pts_u, const fq* pts_v, uint8_t* pflags, const uint32_t* rlist,
                                                                     const uint32_t* rcount) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= *rcount) return;
    stage_rtest(pts_u, pts_v, pflags, rlist[t]);
}

__global__ void __launch_bounds__(BLOCK) k_finalize(int variant, const uint8_t* pflags, const uint8_t* iflags, const uint8_t* eqflags,
                                                    const uint32_t* cwords, size_t n, uint8_t* status, uint8_t* c_out) {
    size_t item = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= n) return;
    uint8_t fl = iflags[item];
    if (eqflags[item]) fl |= IF_EQ0_OK;
    if (variant == VAR_DOUBLE && eqflags[n + item]) fl |= IF_EQ1_OK;
    uint8_t st = stage_status(variant, pflags, fl, n, item);
    status[item] = st;
    if (c_out) {
        uint4 a = make_uint4(0, 0, 0, 0), b = a;
        if (st <= 1) {
            const uint4* p = reinterpret_cast<const uint4*>(cwords + item * 8);
            a = p[0];
            b = p[1];
        }
        uint4* o = reinterpret_cast<uint4*>(c_out + item * 32);
        o[0] = a;
        o[1] = b;
    }
}

// accept bitmap: bit (i % 32) of word i / 32 is set iff status[i] == 0; one warp ballot per word
__global__ void __launch_bounds__(BLOCK) k_bitmap(const uint8_t* status, size_t n, uint32_t* words) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool ok = i < n && status[i] == 0;
    uint32_t w = __ballot_sync(0xffffffffu, ok);
    if ((threadIdx.x & 31) == 0 && i < n) words[i >> 5] = w;
}

__global__ void __launch_bounds__(BLOCK) k_fb_table(niels* out, int which) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= FB_WINDOWS * FB_ENTRIES) return;
    fq u, v;
    fq_load_const(u, which ? JJS_C(GEN_NUMS_UV)[0] : JJS_C(GEN_UV)[0]);
    fq_load_const(v, which ? JJS_C(GEN_NUMS_UV)[1] : JJS_C(GEN_UV)[1]);
    niels e;
    fb_table_entry(e, u, v, t / FB_ENTRIES, t % FB_ENTRIES);
    out[t] = e;
}

// one thread signs one item; inputs are 32-byte little-endian scalars, outputs the reference's wire bytes.
// The variant is a template parameter so that every array index below is static.
template <int VARIANT>
__global__ void __launch_bounds__(BLOCK) k_sign(const uint8_t* sk, const uint8_t* rnd, const uint8_t* gsc, const uint8_t* msg, size_t n,
                                                uint8_t* pk_out, uint8_t* sig_out, Tables T) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    constexpr int PKW = VARIANT == VAR_SINGLE ? 8 : 16, SGW = VARIANT == VAR_DOUBLE ? 24 : 16;
    uint32_t wsk[8], wrnd[8], wg[8], wm[8], pk[16], sig[24];
    wire_load(wsk, WireField{sk, 32}, i);
    wire_load(wrnd, WireField{rnd, 32}, i);
    wire_load(wm, WireField{msg, 32}, i);
    if (VARIANT == VAR_VARGEN) wire_load(wg, WireField{gsc, 32}, i);
    bool ok = sign_item(VARIANT, wsk, wrnd, wg, wm, pk, sig, T);
    uint4* po = reinterpret_cast<uint4*>(pk_out + i * PKW * 4);
    uint4* so = reinterpret_cast<uint4*>(sig_out + i * SGW * 4);
#pragma unroll
    for (int k = 0; k < PKW / 4; k++) po[k] = ok ? make_uint4(pk[4 * k], pk[4 * k + 1], pk[4 * k + 2], pk[4 * k + 3]) : make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int k = 0; k < SGW / 4; k++) so[k] = ok ? make_uint4(sig[4 * k], sig[4 * k + 1], sig[4 * k + 2], sig[4 * k + 3]) : make_uint4(0, 0, 0, 0);
}

// ---- multisig::combine ---------------------------------------------------------------------------------------------
struct MsigBuffers {
    fq *pu, *pv;           // [3K] decoded pk | R | S
    uint8_t* pf;           // [3K]
    uint32_t *d_words, *cd_words;  // [K][8]
    uint32_t* a_words;     // [n][8]
    fq *rsa_u, *rsa_v;     // [n]
    uint8_t* sflags;       // [n]
    uint8_t* share_ok;     // [K]
    const uint32_t *offsets, *owner, *order;
};

__global__ void __launch_bounds__(BLOCK) k_msig_session(MsigBuffers b, size_t K, size_t n, WireField msg, WireField zf, fq* tab, size_t stride) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    size_t s = b.order[t];
    stage_msig_session(b.pu, b.pv, b.pf, K, b.offsets[s], b.offsets[s + 1], msg, zf, s, b.d_words, b.cd_words, b.a_words, b.rsa_u, b.rsa_v, b.sflags,
                       tab + t, stride);
}
__global__ void __launch_bounds__(BLOCK, JJS_EQ_MINBLOCKS) k_msig_share(MsigBuffers b, size_t K, WireField zf, fq* tab, size_t stride, Tables T) {
    size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= K) return;
    size_t s = b.owner[j];
    bool ok = false;
    if (b.sflags[s] == (SF_DECODED | SF_NONEMPTY))
        ok = stage_msig_share(b.pu, b.pv, K, j, zf, b.cd_words, b.a_words + 8 * s, T.fb_g, tab + j, tab + 36 * stride + j, stride);
    b.share_ok[j] = ok ? 1 : 0;
}
__global__ void __launch_bounds__(BLOCK) k_msig_finalize(MsigBuffers b, size_t n, WireField zf, uint8_t* status, uint32_t* bad, uint8_t* sig64) {
    size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    uint32_t sig[16], bi;
    uint8_t st = stage_msig_finalize(b.sflags, b.share_ok, b.offsets[s], b.offsets[s + 1], s, zf, b.rsa_u, b.rsa_v, &bi, sig);
    status[s] = st;
    bad[s] = bi;
    uint4* o = reinterpret_cast<uint4*>(sig64 + 64 * s);
#pragma unroll
    for (int k = 0; k < 4; k++) o[k] = make_uint4(sig[4 * k], sig[4 * k + 1], sig[4 * k + 2], sig[4 * k + 3]);
}

// Aggregate-key items for synthetic batches: signer keys pk_j = sk_j * G, the aggregate secret sum_j d_j sk_j with the
// reference's delinearisation coefficients, and an ordinary hedged signature under it (what a completed SpeedyMuSig
// session verifies as).  At most JJS_MAX_GEN_SIGNERS signers per item.
#define JJS_MAX_GEN_SIGNERS 8
__global__ void __launch_bounds__(BLOCK) k_sign_aggregate(const uint8_t* sk, const uint32_t* offsets, const uint8_t* rnd, const uint8_t* msg, size_t n,
                                                          uint8_t* pks_out, uint8_t* sig_out, Tables T) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t lo = offsets[i], hi = offsets[i + 1];
    uint32_t cnt = hi - lo;
    fq ku[JJS_MAX_GEN_SIGNERS], kv[JJS_MAX_GEN_SIGNERS];
    uint32_t agg[8] = {0, 0, 0, 0, 0, 0, 0, 0}, ord[8], w[8];
    bool ok = cnt <= JJS_MAX_GEN_SIGNERS;
#pragma unroll
    for (int k = 0; k < 8; k++) ord[k] = JJS_C(R_ORDER)[k];
    if (ok) {
#pragma unroll 1
        for (uint32_t j = 0; j < cnt; j++) {
            wire_load(w, WireField{sk, 32}, lo + j);
            ok = ok && fr_wire_is_canonical(w);
            ext P;
            fixedbase_mul(P, T.fb_g, w);
            ext_to_affine(ku[j], kv[j], P);
            uint32_t enc[8];
            point_to_wire(enc, ku[j], kv[j]);
            uint4* o = reinterpret_cast<uint4*>(pks_out + 32 * (size_t)(lo + j));
            o[0] = make_uint4(enc[0], enc[1], enc[2], enc[3]);
            o[1] = make_uint4(enc[4], enc[5], enc[6], enc[7]);
        }
#pragma unroll 1
        for (uint32_t j = 0; j < cnt; j++) {
            Sponge sp;
            sponge_start(sp, (int)(2 + 2 * cnt));
            sponge_absorb(sp, ku[j]);
            sponge_absorb(sp, kv[j]);
#pragma unroll 1
            for (uint32_t k = 0; k < cnt; k++) {
                sponge_absorb(sp, ku[k]);
                sponge_absorb(sp, kv[k]);
            }
            uint32_t dj[8], t[8], s[8];
            sponge_squeeze_truncated(dj, sp);
            wire_load(w, WireField{sk, 32}, lo + j);
            fr_mul(t, dj, w);
            add8(s, agg, t);  // < 2r < 2^256
            uint32_t borrow = sub8(t, s, ord);
#pragma unroll
            for (int k = 0; k < 8; k++) agg[k] = borrow ? s[k] : t[k];
        }
    }
    uint32_t wr[8], wm[8], pk[16], sig[24];
    wire_load(wr, WireField{rnd, 32}, i);
    wire_load(wm, WireField{msg, 32}, i);
    ok = ok && sign_item(VAR_SINGLE, agg, wr, wr, wm, pk, sig, T);
    uint4* so = reinterpret_cast<uint4*>(sig_out + 64 * i);
#pragma unroll
    for (int k = 0; k < 4; k++) so[k] = ok ? make_uint4(sig[4 * k], sig[4 * k + 1], sig[4 * k + 2], sig[4 * k + 3]) : make_uint4(0, 0, 0, 0);
}

struct DeviceState {
    int device = -1;
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    cudaEvent_t copied = nullptr;
    // two internal compute streams: consecutive sub-chunks of a batch alternate between them (each on its own half of
    // the scratch), so the last, partially filled wave of one kernel overlaps the first blocks of the next chunk's
    cudaStream_t sub[2] = {nullptr, nullptr};
    cudaEvent_t fork = nullptr, join[2] = {nullptr, nullptr};
    fq* root_tables = nullptr;
    uint8_t* dlog_hash = nullptr;
    niels* fb_g = nullptr;
    niels* fb_gn = nullptr;
    // pipeline scratch, sized for CHUNK_ITEMS items of the widest variant (4 points, 2 equations)
    fq *pts_u = nullptr, *pts_v = nullptr, *tab = nullptr;
    uint8_t *pflags = nullptr, *iflags = nullptr, *eqflags = nullptr;
    uint32_t* cwords = nullptr;
    uint32_t *rlist = nullptr, *rcount = nullptr;  // signature points awaiting the deferred subgroup test
    uint32_t* eqlist = nullptr;                    // items that go on to the hash and equation kernels (k_work_list)
    // decoded signer keys of the aggregate-key path (grown on demand)
    fq *keys_u = nullptr, *keys_v = nullptr;
    uint8_t* kflags = nullptr;
    uint32_t* kcoef = nullptr;     // delinearisation coefficients, 8 words per key
    size_t cap_keys = 0;
    uint32_t* d_order = nullptr;   // items of a chunk sorted by signer count
    std::vector<uint32_t> h_order;
    uint8_t* agg_stage = nullptr;  // grow-only staging of the aggregate-key host path
    size_t agg_stage_bytes = 0;
    // staging for the host-buffer entry points
    uint8_t *s_pk = nullptr, *s_sig = nullptr, *s_msg = nullptr, *s_status = nullptr, *s_c = nullptr;
    uint32_t* s_bitmap = nullptr;
    size_t stage_items = 0;
    Tables tables() const { return Tables{root_tables, dlog_hash, fb_g, fb_gn}; }
};

// A slice of the pipeline scratch able to hold `cap` items of the widest variant; `tab` serves `cap` threads (the table
// scratch keeps its global stride TAB_THREADS).
struct Region {
    fq *pts_u, *pts_v, *tab;
    uint8_t *pflags, *iflags, *eqflags;
    uint32_t *cwords, *rlist, *rcount, *eqlist;   // rcount[0]: deferred subgroup tests, rcount[2]: length of the work list
    size_t cap;
};
inline Region region_of(const DeviceState& d, size_t first_item, size_t cap, int counter) {
    return Region{d.pts_u + 4 * first_item, d.pts_v + 4 * first_item, d.tab + first_item, d.pflags + 4 * first_item, d.iflags + first_item,
                  d.eqflags + 2 * first_item, d.cwords + 8 * first_item, d.rlist + 2 * first_item, d.rcount + counter, d.eqlist + 2 * first_item, cap};
}
inline Region region_whole(const DeviceState& d) { return region_of(d, 0, CHUNK_ITEMS, 0); }

}  // namespace

struct StageRecord {
    int stage;
    int device;
    cudaEvent_t e0, e1;
};

struct jjs_ctx {
    std::vector<DeviceState> dev;
    char err[512];
    uint64_t launches;
    bool profile;
    std::vector<StageRecord> records;
};

namespace {

int fail(jjs_ctx* ctx, int code, const char* fmt, ...) {
    if (ctx) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(ctx->err, sizeof(ctx->err), fmt, ap);
        va_end(ap);
    }
    return code;
}

#define JJS_CUDA(ctx, call)                                                                                   \
    do {                                                                                                      \
        cudaError_t e_ = (call);                                                                              \
        if (e_ != cudaSuccess) return fail(ctx, e_ == cudaErrorMemoryAllocation ? JJS_ERR_NOMEM : JJS_ERR_CUDA, \
                                           "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

// Optional per-stage timing: CUDA events recorded on the launching stream around each stage's launches.
struct StageTimer {
    jjs_ctx* ctx;
    StageRecord rec;
    bool on;
    StageTimer(jjs_ctx* c, int device, int stage, cudaStream_t s) : ctx(c), on(c->profile) {
        if (!on) return;
        rec.stage = stage;
        rec.device = device;
        cudaEventCreate(&rec.e0);
        cudaEventCreate(&rec.e1);
        cudaEventRecord(rec.e0, s);
    }
    void stop(cudaStream_t s) {
        if (!on) return;
        cudaEventRecord(rec.e1, s);
        ctx->records.push_back(rec);
    }
};

inline unsigned blocks_for(size_t threads) { return (unsigned)((threads + BLOCK - 1) / BLOCK); }

int ensure_scratch(jjs_ctx* ctx, DeviceState& d) {
    if (d.pts_u) return JJS_SUCCESS;
    JJS_CUDA(ctx, cudaSetDevice(d.device));
    JJS_CUDA(ctx, cudaMalloc(&d.pts_u, sizeof(fq) * 4 * CHUNK_ITEMS));
    JJS_CUDA(ctx, cudaMalloc(&d.pts_v, sizeof(fq) * 4 * CHUNK_ITEMS));
    JJS_CUDA(ctx, cudaMalloc(&d.pflags, 4 * CHUNK_ITEMS));
    JJS_CUDA(ctx, cudaMalloc(&d.iflags, CHUNK_ITEMS));
    JJS_CUDA(ctx, cudaMalloc(&d.eqflags, 2 * CHUNK_ITEMS));
    JJS_CUDA(ctx, cudaMalloc(&d.cwords, 32 * CHUNK_ITEMS));
    JJS_CUDA(ctx, cudaMalloc(&d.rlist, sizeof(uint32_t) * 2 * CHUNK_ITEMS));
    JJS_CUDA(ctx, cudaMalloc(&d.rcount, 4 * sizeof(uint32_t)));
    JJS_CUDA(ctx, cudaMalloc(&d.eqlist, sizeof(uint32_t) * 2 * CHUNK_ITEMS));
    JJS_CUDA(ctx, cudaMalloc(&d.tab, sizeof(fq) * AGG_GROUP * 36 * TAB_THREADS));
    return JJS_SUCCESS;
}

int ensure_staging(jjs_ctx* ctx, DeviceState& d, size_t items) {
    if (items <= d.stage_items) return JJS_SUCCESS;
    JJS_CUDA(ctx, cudaSetDevice(d.device));
    cudaFree(d.s_pk); cudaFree(d.s_sig); cudaFree(d.s_msg); cudaFree(d.s_status); cudaFree(d.s_c); cudaFree(d.s_bitmap);
    d.s_pk = d.s_sig = d.s_msg = d.s_status = d.s_c = nullptr;
    d.s_bitmap = nullptr;
    d.stage_items = 0;
    JJS_CUDA(ctx, cudaMalloc(&d.s_pk, 64 * items));
    JJS_CUDA(ctx, cudaMalloc(&d.s_sig, 96 * items));
    JJS_CUDA(ctx, cudaMalloc(&d.s_msg, 32 * items));
    JJS_CUDA(ctx, cudaMalloc(&d.s_status, items));
    JJS_CUDA(ctx, cudaMalloc(&d.s_c, 32 * items));
    JJS_CUDA(ctx, cudaMalloc(&d.s_bitmap, 4 * ((items + 31) / 32)));
    d.stage_items = items;
    return JJS_SUCCESS;
}

void variant_fields(int variant, const uint8_t* pk, const uint8_t* sig, const uint8_t* msg, Fields& pts, WireField& fmsg, WireField& fu) {
    fmsg = WireField{msg, 32};
    if (variant == VAR_SINGLE) {
        pts.f[0] = WireField{pk, 32};
        pts.f[1] = WireField{sig + 32, 64};
        pts.f[2] = pts.f[3] = WireField{nullptr, 0};
        fu = WireField{sig, 64};
    } else if (variant == VAR_DOUBLE) {
        pts.f[0] = WireField{pk, 64};
        pts.f[1] = WireField{pk + 32, 64};
        pts.f[2] = WireField{sig + 32, 96};
        pts.f[3] = WireField{sig + 64, 96};
        fu = WireField{sig, 96};
    } else {
        pts.f[0] = WireField{pk, 64};
        pts.f[1] = WireField{pk + 32, 64};
        pts.f[2] = WireField{sig + 32, 64};
        pts.f[3] = WireField{nullptr, 0};
        fu = WireField{sig, 64};
    }
}
inline size_t pk_size(int variant) { return variant == VAR_SINGLE ? 32 : 64; }
inline size_t sig_size(int variant) { return variant == VAR_DOUBLE ? 96 : 64; }

// keys (and the var-gen generator) get their subgroup test in k_decode; signature points get it from the equation
inline uint32_t key_slot_mask(int variant) { return variant == VAR_SINGLE ? 1u : 3u; }

// Hash stage for a chunk of m <= R.cap items: the work list, then the sponge of the items on it.
int enqueue_challenges(jjs_ctx* ctx, DeviceState& d, const Region& R, int variant, size_t m, const WireField& fmsg, const WireField& fu,
                       cudaStream_t stream, bool require_valid_keys) {
    StageTimer t1(ctx, d.device, 1, stream);
    JJS_CUDA(ctx, cudaMemsetAsync(R.rcount + 2, 0, sizeof(uint32_t), stream));
    k_work_list<<<blocks_for(m), BLOCK, 0, stream>>>(variant, R.pflags, m, fmsg, fu, require_valid_keys, R.iflags, R.eqflags, R.cwords, R.eqlist,
                                                    R.rcount + 2);
    k_challenge<<<blocks_for(m), BLOCK, 0, stream>>>(variant, R.pts_u, R.pts_v, m, fmsg, R.eqlist, R.rcount + 2, R.cwords);
    t1.stop(stream);
    ctx->launches += 2;
    return JJS_SUCCESS;
}

// Equation stage for a chunk of m <= R.cap items (after enqueue_challenges, whose work list it shares): the
// equations, then the deferred subgroup tests they asked for.
int enqueue_equations(jjs_ctx* ctx, DeviceState& d, const Region& R, int variant, size_t m, const WireField& fu, cudaStream_t stream) {
    const int neq = variant == VAR_DOUBLE ? 2 : 1;
    Tables T = d.tables();
    StageTimer t3(ctx, d.device, 3, stream);
    JJS_CUDA(ctx, cudaMemsetAsync(R.rcount, 0, sizeof(uint32_t), stream));
    for (size_t first = 0; first < neq * m; first += R.cap) {
        size_t cnt = neq * m - first < R.cap ? neq * m - first : R.cap;
        k_equation<<<blocks_for(cnt), BLOCK, 0, stream>>>(variant, R.pts_u, R.pts_v, R.pflags, m, first, cnt, R.eqlist, R.rcount + 2, fu, R.cwords,
                                                         R.eqflags, R.tab, TAB_THREADS, T, R.rlist, R.rcount);
        ctx->launches++;
    }
    t3.stop(stream);
    StageTimer t5(ctx, d.device, 5, stream);
    k_rtest<<<blocks_for(neq * m), BLOCK, 0, stream>>>(R.pts_u, R.pts_v, R.pflags, R.rlist, R.rcount);
    ctx->launches++;
    t5.stop(stream);
    return JJS_SUCCESS;
}

// The whole pipeline for one chunk of m <= R.cap items (device pointers) on `stream`.
int run_chunk(jjs_ctx* ctx, DeviceState& d, const Region& R, int variant, const uint8_t* pk, const uint8_t* sig, const uint8_t* msg, size_t m,
              uint8_t* status, uint8_t* c_out, cudaStream_t stream, bool challenge_only) {
    const int slots = variant_slots(variant);
    Tables T = d.tables();
    Fields pts;
    WireField fmsg, fu;
    variant_fields(variant, pk, sig, msg, pts, fmsg, fu);
    StageTimer t0(ctx, d.device, 0, stream);
    k_decode<<<blocks_for(slots * m), BLOCK, 0, stream>>>(pts, slots, 0, m, R.pts_u, R.pts_v, R.pflags, T, key_slot_mask(variant));
    t0.stop(stream);
    ctx->launches++;
    int rc = enqueue_challenges(ctx, d, R, variant, m, fmsg, fu, stream, !challenge_only);
    if (rc) return rc;
    if (challenge_only) {
        JJS_CUDA(ctx, cudaMemcpyAsync(c_out, R.cwords, 32 * m, cudaMemcpyDeviceToDevice, stream));
        return JJS_SUCCESS;
    }
    rc = enqueue_equations(ctx, d, R, variant, m, fu, stream);
    if (rc) return rc;
    StageTimer t4(ctx, d.device, 4, stream);
    k_finalize<<<blocks_for(m), BLOCK, 0, stream>>>(variant, R.pflags, R.iflags, R.eqflags, R.cwords, m, status, c_out);
    t4.stop(stream);
    ctx->launches++;
    return JJS_SUCCESS;
}

constexpr size_t SUB_ITEMS = CHUNK_ITEMS / 2;      // items per scratch half
constexpr size_t SUB_CHUNK = size_t(1) << 18;      // items per sub-chunk of an overlapped batch

// Enqueue the whole pipeline for n items (device pointers); the work is ordered after what `stream` holds now and
// `stream` waits for it.  Batches above one sub-chunk are cut into sub-chunks that alternate between the two internal
// streams and scratch halves: kernels of neighbouring sub-chunks overlap, which hides the partially filled last wave
// of every launch (per-thread work is uniform, so a launch ends with SMs idling for up to one block duration).
// With per-stage profiling on, everything stays on `stream` so that the stage timers do not overlap.
int run_device(jjs_ctx* ctx, DeviceState& d, int variant, const uint8_t* pk, const uint8_t* sig, const uint8_t* msg, size_t n,
               uint8_t* status, uint8_t* c_out, cudaStream_t stream, bool challenge_only = false) {
    int rc = ensure_scratch(ctx, d);
    if (rc) return rc;
    JJS_CUDA(ctx, cudaSetDevice(d.device));
    const bool overlap = !ctx->profile && n > SUB_CHUNK;
    if (overlap) {
        JJS_CUDA(ctx, cudaEventRecord(d.fork, stream));
        for (int k = 0; k < 2; k++) JJS_CUDA(ctx, cudaStreamWaitEvent(d.sub[k], d.fork, 0));
    }
    const size_t step = overlap ? SUB_CHUNK : CHUNK_ITEMS;
    size_t i = 0;
    for (size_t off = 0; off < n; off += step, i++) {
        size_t m = n - off < step ? n - off : step;
        Region R = overlap ? region_of(d, (i & 1) * SUB_ITEMS, SUB_ITEMS, (int)(i & 1)) : region_whole(d);
        rc = run_chunk(ctx, d, R, variant, pk + off * pk_size(variant), sig + off * sig_size(variant), msg + off * 32, m,
                       status ? status + off : nullptr, c_out ? c_out + off * 32 : nullptr, overlap ? d.sub[i & 1] : stream, challenge_only);
        if (rc) return rc;
    }
    if (overlap)
        for (int k = 0; k < 2; k++) {
            JJS_CUDA(ctx, cudaEventRecord(d.join[k], d.sub[k]));
            JJS_CUDA(ctx, cudaStreamWaitEvent(stream, d.join[k], 0));
        }
    JJS_CUDA(ctx, cudaGetLastError());
    return JJS_SUCCESS;
}

// Host-buffer path: contiguous shards over the context's devices, one stream each, joined before return.
int run_host(jjs_ctx* ctx, int variant, const uint8_t* pk, const uint8_t* sig, const uint8_t* msg, size_t n, uint8_t* status,
             uint8_t* c_out, bool challenge_only = false, uint32_t* bitmap = nullptr) {
    if (!ctx) return JJS_ERR_ARGUMENT;
    ctx->err[0] = 0;
    if (n == 0) return JJS_SUCCESS;
    if (!pk || !sig || !msg || (!status && !challenge_only && !bitmap) || (challenge_only && !c_out)) return fail(ctx, JJS_ERR_ARGUMENT, "null buffer");
    const size_t g = ctx->dev.size();
    const size_t per = ((n + g - 1) / g + 31) & ~size_t(31);  // shards start on a bitmap word
    const size_t pks = pk_size(variant), sgs = sig_size(variant);
    // Each device's shard is cut into pipeline slices: slice j + 1 is copied in on the copy stream while slice j is
    // being verified (the kernels of one slice run far longer than its 128-192 B/item copy), and consecutive slices
    // alternate between the two internal compute streams and scratch halves, like the sub-chunks of run_device.
    // The first slice is short so that compute starts early.
    for (size_t k = 0; k < g; k++) {
        size_t lo = k * per, hi = lo + per < n ? lo + per : n;
        if (lo >= hi) break;
        DeviceState& d = ctx->dev[k];
        size_t m = hi - lo;
        int rc = ensure_staging(ctx, d, m);
        if (!rc) rc = ensure_scratch(ctx, d);
        if (rc) return rc;
        JJS_CUDA(ctx, cudaSetDevice(d.device));
        uint8_t* dc = (c_out || challenge_only) ? d.s_c : nullptr;
        const bool serial = ctx->profile;  // stage timers must not overlap
        size_t j = 0;
        for (size_t off = 0; off < m; j++) {
            size_t want = j == 0 ? SUB_CHUNK / 4 : SUB_CHUNK;
            size_t cnt = m - off < want ? m - off : want;
            JJS_CUDA(ctx, cudaMemcpyAsync(d.s_pk + off * pks, pk + (lo + off) * pks, cnt * pks, cudaMemcpyHostToDevice, d.copy_stream));
            JJS_CUDA(ctx, cudaMemcpyAsync(d.s_sig + off * sgs, sig + (lo + off) * sgs, cnt * sgs, cudaMemcpyHostToDevice, d.copy_stream));
            JJS_CUDA(ctx, cudaMemcpyAsync(d.s_msg + off * 32, msg + (lo + off) * 32, cnt * 32, cudaMemcpyHostToDevice, d.copy_stream));
            JJS_CUDA(ctx, cudaEventRecord(d.copied, d.copy_stream));
            cudaStream_t cs = serial ? d.stream : d.sub[j & 1];
            JJS_CUDA(ctx, cudaStreamWaitEvent(cs, d.copied, 0));
            Region R = region_of(d, (j & 1) * SUB_ITEMS, SUB_ITEMS, (int)(j & 1));
            rc = run_chunk(ctx, d, R, variant, d.s_pk + off * pks, d.s_sig + off * sgs, d.s_msg + off * 32, cnt, d.s_status + off,
                           dc ? dc + off * 32 : nullptr, cs, challenge_only);
            if (rc) return rc;
            off += cnt;
        }
        if (!serial)
            for (int q = 0; q < 2; q++) {
                JJS_CUDA(ctx, cudaEventRecord(d.join[q], d.sub[q]));
                JJS_CUDA(ctx, cudaStreamWaitEvent(d.stream, d.join[q], 0));
            }
        if (!challenge_only && status) JJS_CUDA(ctx, cudaMemcpyAsync(status + lo, d.s_status, m, cudaMemcpyDeviceToHost, d.stream));
        if (c_out) JJS_CUDA(ctx, cudaMemcpyAsync(c_out + lo * 32, d.s_c, m * 32, cudaMemcpyDeviceToHost, d.stream));
        if (bitmap) {
            k_bitmap<<<blocks_for(m), BLOCK, 0, d.stream>>>(d.s_status, m, d.s_bitmap);
            ctx->launches++;
            JJS_CUDA(ctx, cudaMemcpyAsync(bitmap + lo / 32, d.s_bitmap, 4 * ((m + 31) / 32), cudaMemcpyDeviceToHost, d.stream));
        }
    }
    for (size_t k = 0; k < g; k++) {
        JJS_CUDA(ctx, cudaSetDevice(ctx->dev[k].device));
        JJS_CUDA(ctx, cudaStreamSynchronize(ctx->dev[k].stream));
    }
    return JJS_SUCCESS;
}

// aggregate_pk(..).verify(..) for n items whose wire data already sits on the device.  `h_offsets` is the host copy of
// the n + 1 offsets (needed to plan the chunks); d_offsets the same array on the device.
int run_aggregate_device(jjs_ctx* ctx, DeviceState& d, const uint8_t* d_pks, const uint32_t* d_offsets, const uint32_t* h_offsets,
                         const uint8_t* d_sig, const uint8_t* d_msg, size_t n, uint8_t* d_status, uint8_t* d_c, uint8_t* d_agg,
                         cudaStream_t stream) {
    int rc = ensure_scratch(ctx, d);
    if (rc) return rc;
    JJS_CUDA(ctx, cudaSetDevice(d.device));
    Tables T = d.tables();
    for (size_t off = 0; off < n; off += CHUNK_ITEMS) {
        size_t m = n - off < CHUNK_ITEMS ? n - off : CHUNK_ITEMS;
        uint32_t key_lo = h_offsets[off], key_hi = h_offsets[off + m];
        size_t K = key_hi - key_lo;
        if (K > d.cap_keys) {
            JJS_CUDA(ctx, cudaStreamSynchronize(stream));
            cudaFree(d.keys_u); cudaFree(d.keys_v); cudaFree(d.kflags); cudaFree(d.kcoef);
            d.keys_u = d.keys_v = nullptr; d.kflags = nullptr; d.kcoef = nullptr; d.cap_keys = 0;
            JJS_CUDA(ctx, cudaMalloc(&d.keys_u, sizeof(fq) * K));
            JJS_CUDA(ctx, cudaMalloc(&d.keys_v, sizeof(fq) * K));
            JJS_CUDA(ctx, cudaMalloc(&d.kflags, K));
            JJS_CUDA(ctx, cudaMalloc(&d.kcoef, 32 * K));
            d.cap_keys = K;
        }
        Fields fk, fr;
        fk.f[0] = WireField{d_pks + 32 * (size_t)key_lo, 32};
        fr.f[0] = WireField{d_sig + 64 * off + 32, 64};
        fk.f[1] = fk.f[2] = fk.f[3] = fr.f[1] = fr.f[2] = fr.f[3] = WireField{nullptr, 0};
        WireField fmsg{d_msg + 32 * off, 32}, fu{d_sig + 64 * off, 64};
        StageTimer t0(ctx, d.device, 0, stream);
        if (K) k_decode<<<blocks_for(K), BLOCK, 0, stream>>>(fk, 1, 0, K, d.keys_u, d.keys_v, d.kflags, T, 0u);
        k_decode<<<blocks_for(m), BLOCK, 0, stream>>>(fr, 1, 1, m, d.pts_u, d.pts_v, d.pflags, T, 0u);
        t0.stop(stream);
        // counting sort of the chunk's items by signer count (host side; counts above 63 share the last bucket)
        {
            size_t hist[65] = {0};
            for (size_t i = 0; i < m; i++) {
                uint32_t c = h_offsets[off + i + 1] - h_offsets[off + i];
                hist[(c > 63 ? 63 : c) + 1]++;
            }
            for (int b = 0; b < 64; b++) hist[b + 1] += hist[b];
            d.h_order.resize(m);
            for (size_t i = 0; i < m; i++) {
                uint32_t c = h_offsets[off + i + 1] - h_offsets[off + i];
                d.h_order[hist[c > 63 ? 63 : c]++] = (uint32_t)i;
            }
            if (!d.d_order) JJS_CUDA(ctx, cudaMalloc(&d.d_order, sizeof(uint32_t) * CHUNK_ITEMS));
            JJS_CUDA(ctx, cudaStreamSynchronize(stream));  // h_order is reused by the next chunk / call
            JJS_CUDA(ctx, cudaMemcpyAsync(d.d_order, d.h_order.data(), sizeof(uint32_t) * m, cudaMemcpyHostToDevice, stream));
        }
        StageTimer t2(ctx, d.device, 2, stream);
        k_agg_coeffs<<<blocks_for(m), BLOCK, 0, stream>>>(d.keys_u, d.keys_v, d.kflags, d_offsets + off, d.d_order, key_lo, m, d.kcoef);
        k_aggregate<<<blocks_for(m), BLOCK, 0, stream>>>(d.keys_u, d.keys_v, d.kflags, d_offsets + off, d.d_order, key_lo, m, d.pts_u, d.pts_v,
                                                        d.pflags, d_agg ? d_agg + 32 * off : nullptr, d.tab, TAB_THREADS, d.kcoef);
        t2.stop(stream);
        int rc2 = enqueue_challenges(ctx, d, region_whole(d), VAR_SINGLE, m, fmsg, fu, stream, true);
        if (!rc2) rc2 = enqueue_equations(ctx, d, region_whole(d), VAR_SINGLE, m, fu, stream);
        if (rc2) return rc2;
        StageTimer t4(ctx, d.device, 4, stream);
        k_finalize<<<blocks_for(m), BLOCK, 0, stream>>>(VAR_SINGLE, d.pflags, d.iflags, d.eqflags, d.cwords, m, d_status + off,
                                                       d_c ? d_c + 32 * off : nullptr);
        t4.stop(stream);
        ctx->launches += 5;   // two decodes, the two aggregation kernels, finalize (the helpers count their own)
    }
    JJS_CUDA(ctx, cudaGetLastError());
    return JJS_SUCCESS;
}

// host buffers: contiguous shards over the devices, each shard copied to temporary device buffers
int run_aggregate(jjs_ctx* ctx, const uint8_t* pks, const uint32_t* offsets, const uint8_t* sig, const uint8_t* msg, size_t n, uint8_t* status,
                  uint8_t* c_out, uint8_t* agg_out) {
    if (!ctx) return JJS_ERR_ARGUMENT;
    ctx->err[0] = 0;
    if (n == 0) return JJS_SUCCESS;
    if (!offsets || !sig || !msg || !status) return fail(ctx, JJS_ERR_ARGUMENT, "null buffer");
    for (size_t i = 0; i < n; i++)
        if (offsets[i + 1] < offsets[i]) return fail(ctx, JJS_ERR_ARGUMENT, "offsets must be non-decreasing");
    if (offsets[n] > offsets[0] && !pks) return fail(ctx, JJS_ERR_ARGUMENT, "null key buffer");
    const size_t g = ctx->dev.size();
    const size_t per = (n + g - 1) / g;
    int rc = JJS_SUCCESS;
    for (size_t k = 0; k < g && rc == JJS_SUCCESS; k++) {
        size_t lo = k * per, hi = lo + per < n ? lo + per : n;
        if (lo >= hi) break;
        DeviceState& d = ctx->dev[k];
        size_t m = hi - lo, K = offsets[hi] - offsets[lo];
        cudaSetDevice(d.device);
        // layout: keys | sig | msg | status(+pad) | c | agg | offsets
        size_t o_sig = 32 * K, o_msg = o_sig + 64 * m, o_st = o_msg + 32 * m, o_c = o_st + ((m + 31) / 32) * 32, o_agg = o_c + 32 * m,
               o_off = o_agg + 32 * m, total = o_off + 4 * (m + 1);
        if (total > d.agg_stage_bytes) {
            cudaStreamSynchronize(d.stream);
            cudaFree(d.agg_stage);
            d.agg_stage = nullptr;
            d.agg_stage_bytes = 0;
            cudaError_t e = cudaMalloc(&d.agg_stage, total);
            if (e != cudaSuccess) { rc = fail(ctx, JJS_ERR_NOMEM, "cudaMalloc: %s", cudaGetErrorString(e)); break; }
            d.agg_stage_bytes = total;
        }
        uint8_t* b = d.agg_stage;
        if (K) cudaMemcpyAsync(b, pks + 32 * (size_t)offsets[lo], 32 * K, cudaMemcpyHostToDevice, d.stream);
        cudaMemcpyAsync(b + o_sig, sig + 64 * lo, 64 * m, cudaMemcpyHostToDevice, d.stream);
        cudaMemcpyAsync(b + o_msg, msg + 32 * lo, 32 * m, cudaMemcpyHostToDevice, d.stream);
        cudaMemcpyAsync(b + o_off, offsets + lo, 4 * (m + 1), cudaMemcpyHostToDevice, d.stream);
        // device keys start at offsets[lo]: hand the kernels a key pointer that makes absolute offsets valid
        rc = run_aggregate_device(ctx, d, b - 32 * (size_t)offsets[lo], reinterpret_cast<uint32_t*>(b + o_off), offsets + lo, b + o_sig, b + o_msg, m,
                                  b + o_st, c_out ? b + o_c : nullptr, agg_out ? b + o_agg : nullptr, d.stream);
        if (rc) break;
        cudaMemcpyAsync(status + lo, b + o_st, m, cudaMemcpyDeviceToHost, d.stream);
        if (c_out) cudaMemcpyAsync(c_out + 32 * lo, b + o_c, 32 * m, cudaMemcpyDeviceToHost, d.stream);
        if (agg_out) cudaMemcpyAsync(agg_out + 32 * lo, b + o_agg, 32 * m, cudaMemcpyDeviceToHost, d.stream);
    }
    for (size_t k = 0; k < g; k++) {
        cudaSetDevice(ctx->dev[k].device);
        cudaError_t e = cudaStreamSynchronize(ctx->dev[k].stream);
        if (e != cudaSuccess && rc == JJS_SUCCESS) rc = fail(ctx, JJS_ERR_CUDA, "aggregate verify failed: %s", cudaGetErrorString(e));
    }
    return rc;
}

// typed inputs: points as JubJubExtended coordinates, scalars as canonical bytes; host buffers, device 0
int run_ext(jjs_ctx* ctx, int variant, const uint8_t* pts, const uint8_t* u32, const uint8_t* msg, size_t n, uint8_t* status, uint8_t* c_out) {
    if (!ctx) return JJS_ERR_ARGUMENT;
    ctx->err[0] = 0;
    if (variant < 0 || variant > 2) return fail(ctx, JJS_ERR_ARGUMENT, "bad variant");
    if (n == 0) return JJS_SUCCESS;
    if (!pts || !u32 || !msg || !status) return fail(ctx, JJS_ERR_ARGUMENT, "null buffer");
    const int slots = variant_slots(variant);
    const size_t g = ctx->dev.size(), per = (n + g - 1) / g;
    // Same structure as run_host: every device gets a contiguous shard, staged whole on the device; the shard is cut into
    // slices (a short first one) whose H2D copies run on the copy stream while earlier slices are being verified on the
    // two alternating compute streams / scratch halves.  Nothing is synchronised before every device has its work.
    for (size_t k = 0; k < g; k++) {
        size_t lo = k * per, hi = lo + per < n ? lo + per : n;
        if (lo >= hi) break;
        DeviceState& d = ctx->dev[k];
        int rc = ensure_scratch(ctx, d);
        if (rc) return rc;
        JJS_CUDA(ctx, cudaSetDevice(d.device));
        const size_t m = hi - lo, ptw = 160 * (size_t)slots;
        const size_t need = m * (ptw + 32 + 32 + 32 + 1);
        if (need > d.agg_stage_bytes) {
            JJS_CUDA(ctx, cudaStreamSynchronize(d.stream));
            cudaFree(d.agg_stage);
            d.agg_stage = nullptr;
            d.agg_stage_bytes = 0;
            JJS_CUDA(ctx, cudaMalloc(&d.agg_stage, need));
            d.agg_stage_bytes = need;
        }
        uint8_t *b_pts = d.agg_stage, *b_u = b_pts + ptw * m, *b_msg = b_u + 32 * m, *b_c = b_msg + 32 * m, *b_st = b_c + 32 * m;
        const bool serial = ctx->profile;
        size_t j = 0;
        for (size_t off = 0; off < m; j++) {
            size_t want = j == 0 ? SUB_CHUNK / 4 : SUB_CHUNK;
            size_t cnt = m - off < want ? m - off : want;
            JJS_CUDA(ctx, cudaMemcpyAsync(b_pts + ptw * off, pts + ptw * (lo + off), ptw * cnt, cudaMemcpyHostToDevice, d.copy_stream));
            JJS_CUDA(ctx, cudaMemcpyAsync(b_u + 32 * off, u32 + 32 * (lo + off), 32 * cnt, cudaMemcpyHostToDevice, d.copy_stream));
            JJS_CUDA(ctx, cudaMemcpyAsync(b_msg + 32 * off, msg + 32 * (lo + off), 32 * cnt, cudaMemcpyHostToDevice, d.copy_stream));
            JJS_CUDA(ctx, cudaEventRecord(d.copied, d.copy_stream));
            cudaStream_t cs = serial ? d.stream : d.sub[j & 1];
            JJS_CUDA(ctx, cudaStreamWaitEvent(cs, d.copied, 0));
            Region R = region_of(d, (j & 1) * SUB_ITEMS, SUB_ITEMS, (int)(j & 1));
            WireField fmsg{b_msg + 32 * off, 32}, fu{b_u + 32 * off, 32};
            StageTimer t0(ctx, d.device, 0, cs);
            k_decode_ext<<<blocks_for(slots * cnt), BLOCK, 0, cs>>>(b_pts + ptw * off, slots, cnt, R.pts_u, R.pts_v, R.pflags, key_slot_mask(variant));
            t0.stop(cs);
            rc = enqueue_challenges(ctx, d, R, variant, cnt, fmsg, fu, cs, true);
            if (!rc) rc = enqueue_equations(ctx, d, R, variant, cnt, fu, cs);
            if (rc) return rc;
            k_finalize<<<blocks_for(cnt), BLOCK, 0, cs>>>(variant, R.pflags, R.iflags, R.eqflags, R.cwords, cnt, b_st + off, c_out ? b_c + 32 * off : nullptr);
            ctx->launches += 2;   // decode, finalize
            off += cnt;
        }
        if (!serial)
            for (int q = 0; q < 2; q++) {
                JJS_CUDA(ctx, cudaEventRecord(d.join[q], d.sub[q]));
                JJS_CUDA(ctx, cudaStreamWaitEvent(d.stream, d.join[q], 0));
            }
        JJS_CUDA(ctx, cudaMemcpyAsync(status + lo, b_st, m, cudaMemcpyDeviceToHost, d.stream));
        if (c_out) JJS_CUDA(ctx, cudaMemcpyAsync(c_out + 32 * lo, b_c, 32 * m, cudaMemcpyDeviceToHost, d.stream));
    }
    for (size_t k = 0; k < g; k++) {
        JJS_CUDA(ctx, cudaSetDevice(ctx->dev[k].device));
        cudaError_t e = cudaStreamSynchronize(ctx->dev[k].stream);
        if (e != cudaSuccess) return fail(ctx, JJS_ERR_CUDA, "typed verify failed: %s", cudaGetErrorString(e));
    }
    JJS_CUDA(ctx, cudaGetLastError());
    return JJS_SUCCESS;
}

// multisig::combine for n ragged sessions (host buffers, device 0); chunks hold at most 2^20 participants
int run_msig(jjs_ctx* ctx, const uint8_t* pks, const uint8_t* Rs, const uint8_t* Ss, const uint8_t* zs, const uint32_t* offsets, const uint8_t* msg,
             size_t n, uint8_t* share_ok, uint8_t* status, uint32_t* bad, uint8_t* sig) {
    if (!ctx) return JJS_ERR_ARGUMENT;
    ctx->err[0] = 0;
    if (n == 0) return JJS_SUCCESS;
    if (!offsets || !msg || !status) return fail(ctx, JJS_ERR_ARGUMENT, "null buffer");
    if (offsets[0] != 0) return fail(ctx, JJS_ERR_ARGUMENT, "offsets[0] must be 0");
    for (size_t i = 0; i < n; i++)
        if (offsets[i + 1] < offsets[i]) return fail(ctx, JJS_ERR_ARGUMENT, "offsets must be non-decreasing");
    if (offsets[n] && (!pks || !Rs || !Ss || !zs)) return fail(ctx, JJS_ERR_ARGUMENT, "null buffer");
    DeviceState& d = ctx->dev[0];
    int rc = ensure_scratch(ctx, d);
    if (rc) return rc;
    JJS_CUDA(ctx, cudaSetDevice(d.device));
    Tables T = d.tables();
    std::vector<uint32_t> rel, owner, order;
    for (size_t s0 = 0; s0 < n;) {
        size_t s1 = s0;
        while (s1 < n && s1 - s0 < CHUNK_ITEMS && offsets[s1 + 1] - offsets[s0] <= TAB_THREADS) s1++;
        if (s1 == s0) return fail(ctx, JJS_ERR_ARGUMENT, "a session has more than 2^20 participants");
        const size_t m = s1 - s0, K = offsets[s1] - offsets[s0], k0 = offsets[s0];
        rel.resize(m + 1);
        owner.resize(K ? K : 1);
        order.resize(m);
        size_t hist[34] = {0};
        for (size_t i = 0; i <= m; i++) rel[i] = offsets[s0 + i] - (uint32_t)k0;
        for (size_t i = 0; i < m; i++) {
            uint32_t c = rel[i + 1] - rel[i];
            for (uint32_t j = rel[i]; j < rel[i + 1]; j++) owner[j] = (uint32_t)i;
            hist[(c > 32 ? 32 : c) + 1]++;
        }
        for (int b = 0; b < 33; b++) hist[b + 1] += hist[b];
        for (size_t i = 0; i < m; i++) {
            uint32_t c = rel[i + 1] - rel[i];
            order[hist[c > 32 ? 32 : c]++] = (uint32_t)i;
        }
        // one device buffer: wire inputs | decoded points | scalars | results | index arrays
        size_t o = 0;
        auto take = [&](size_t bytes) { size_t at = o; o += (bytes + 255) / 256 * 256; return at; };
        size_t o_pk = take(32 * K), o_R = take(32 * K), o_S = take(32 * K), o_z = take(32 * K), o_msg = take(32 * m), o_pu = take(sizeof(fq) * 3 * K),
               o_pv = take(sizeof(fq) * 3 * K), o_pf = take(3 * K), o_d = take(32 * K), o_cd = take(32 * K), o_a = take(32 * m), o_ru = take(sizeof(fq) * m),
               o_rv = take(sizeof(fq) * m), o_sf = take(m), o_ok = take(K), o_off = take(4 * (m + 1)), o_own = take(4 * K), o_ord = take(4 * m),
               o_st = take(m), o_bad = take(4 * m), o_sig = take(64 * m);
        if (o > d.agg_stage_bytes) {
            cudaStreamSynchronize(d.stream);
            cudaFree(d.agg_stage);
            d.agg_stage = nullptr;
            d.agg_stage_bytes = 0;
            JJS_CUDA(ctx, cudaMalloc(&d.agg_stage, o));
            d.agg_stage_bytes = o;
        }
        uint8_t* B = d.agg_stage;
        cudaStream_t st = d.stream;
        if (K) {
            cudaMemcpyAsync(B + o_pk, pks + 32 * k0, 32 * K, cudaMemcpyHostToDevice, st);
            cudaMemcpyAsync(B + o_R, Rs + 32 * k0, 32 * K, cudaMemcpyHostToDevice, st);
            cudaMemcpyAsync(B + o_S, Ss + 32 * k0, 32 * K, cudaMemcpyHostToDevice, st);
            cudaMemcpyAsync(B + o_z, zs + 32 * k0, 32 * K, cudaMemcpyHostToDevice, st);
            cudaMemcpyAsync(B + o_own, owner.data(), 4 * K, cudaMemcpyHostToDevice, st);
        }
        cudaMemcpyAsync(B + o_msg, msg + 32 * s0, 32 * m, cudaMemcpyHostToDevice, st);
        cudaMemcpyAsync(B + o_off, rel.data(), 4 * (m + 1), cudaMemcpyHostToDevice, st);
        cudaMemcpyAsync(B + o_ord, order.data(), 4 * m, cudaMemcpyHostToDevice, st);
        MsigBuffers b;
        b.pu = reinterpret_cast<fq*>(B + o_pu); b.pv = reinterpret_cast<fq*>(B + o_pv); b.pf = B + o_pf;
        b.d_words = reinterpret_cast<uint32_t*>(B + o_d); b.cd_words = reinterpret_cast<uint32_t*>(B + o_cd); b.a_words = reinterpret_cast<uint32_t*>(B + o_a);
        b.rsa_u = reinterpret_cast<fq*>(B + o_ru); b.rsa_v = reinterpret_cast<fq*>(B + o_rv); b.sflags = B + o_sf; b.share_ok = B + o_ok;
        b.offsets = reinterpret_cast<uint32_t*>(B + o_off); b.owner = reinterpret_cast<uint32_t*>(B + o_own); b.order = reinterpret_cast<uint32_t*>(B + o_ord);
        Fields f;
        f.f[0] = WireField{B + o_pk, 32}; f.f[1] = WireField{B + o_R, 32}; f.f[2] = WireField{B + o_S, 32}; f.f[3] = WireField{nullptr, 0};
        WireField fmsg{B + o_msg, 32}, fz{B + o_z, 32};
        if (K) k_decode<<<blocks_for(3 * K), BLOCK, 0, st>>>(f, 3, 0, K, b.pu, b.pv, b.pf, T, 0u);
        k_msig_session<<<blocks_for(m), BLOCK, 0, st>>>(b, K, m, fmsg, fz, d.tab, TAB_THREADS);
        if (K) k_msig_share<<<blocks_for(K), BLOCK, 0, st>>>(b, K, fz, d.tab, TAB_THREADS, T);
        k_msig_finalize<<<blocks_for(m), BLOCK, 0, st>>>(b, m, fz, B + o_st, reinterpret_cast<uint32_t*>(B + o_bad), B + o_sig);
        ctx->launches += 4;
        cudaMemcpyAsync(status + s0, B + o_st, m, cudaMemcpyDeviceToHost, st);
        if (bad) cudaMemcpyAsync(bad + s0, B + o_bad, 4 * m, cudaMemcpyDeviceToHost, st);
        if (sig) cudaMemcpyAsync(sig + 64 * s0, B + o_sig, 64 * m, cudaMemcpyDeviceToHost, st);
        if (share_ok && K) cudaMemcpyAsync(share_ok + k0, B + o_ok, K, cudaMemcpyDeviceToHost, st);
        cudaError_t e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) return fail(ctx, JJS_ERR_CUDA, "multisig combine failed: %s", cudaGetErrorString(e));
        s0 = s1;
    }
    return JJS_SUCCESS;
}

int run_sign(jjs_ctx* ctx, int variant, const uint8_t* sk, const uint8_t* rnd, const uint8_t* gsc, const uint8_t* msg, size_t n, uint8_t* pk_out,
             uint8_t* sig_out) {
    if (!ctx) return JJS_ERR_ARGUMENT;
    ctx->err[0] = 0;
    if (variant < 0 || variant > 2) return fail(ctx, JJS_ERR_ARGUMENT, "bad variant");
    if (n == 0) return JJS_SUCCESS;
    if (!sk || !rnd || !msg || !pk_out || !sig_out || (variant == VAR_VARGEN && !gsc)) return fail(ctx, JJS_ERR_ARGUMENT, "null buffer");
    DeviceState& d = ctx->dev[0];
    JJS_CUDA(ctx, cudaSetDevice(d.device));
    const size_t pks = pk_size(variant), sgs = sig_size(variant);
    uint8_t* buf = nullptr;
    const size_t chunk = n < CHUNK_ITEMS ? n : CHUNK_ITEMS;
    JJS_CUDA(ctx, cudaMalloc(&buf, chunk * (4 * 32 + pks + sgs)));
    uint8_t *d_sk = buf, *d_rnd = buf + 32 * chunk, *d_g = buf + 64 * chunk, *d_msg = buf + 96 * chunk, *d_pk = buf + 128 * chunk,
            *d_sig = d_pk + pks * chunk;
    int rc = JJS_SUCCESS;
    for (size_t off = 0; off < n && rc == JJS_SUCCESS; off += chunk) {
        size_t m = n - off < chunk ? n - off : chunk;
        cudaMemcpyAsync(d_sk, sk + 32 * off, 32 * m, cudaMemcpyHostToDevice, d.stream);
        cudaMemcpyAsync(d_rnd, rnd + 32 * off, 32 * m, cudaMemcpyHostToDevice, d.stream);
        if (variant == VAR_VARGEN) cudaMemcpyAsync(d_g, gsc + 32 * off, 32 * m, cudaMemcpyHostToDevice, d.stream);
        cudaMemcpyAsync(d_msg, msg + 32 * off, 32 * m, cudaMemcpyHostToDevice, d.stream);
        if (variant == VAR_SINGLE) k_sign<VAR_SINGLE><<<blocks_for(m), BLOCK, 0, d.stream>>>(d_sk, d_rnd, d_g, d_msg, m, d_pk, d_sig, d.tables());
        else if (variant == VAR_DOUBLE) k_sign<VAR_DOUBLE><<<blocks_for(m), BLOCK, 0, d.stream>>>(d_sk, d_rnd, d_g, d_msg, m, d_pk, d_sig, d.tables());
        else k_sign<VAR_VARGEN><<<blocks_for(m), BLOCK, 0, d.stream>>>(d_sk, d_rnd, d_g, d_msg, m, d_pk, d_sig, d.tables());
        ctx->launches++;
        cudaMemcpyAsync(pk_out + pks * off, d_pk, pks * m, cudaMemcpyDeviceToHost, d.stream);
        cudaMemcpyAsync(sig_out + sgs * off, d_sig, sgs * m, cudaMemcpyDeviceToHost, d.stream);
        cudaError_t e = cudaStreamSynchronize(d.stream);
        if (e != cudaSuccess) rc = fail(ctx, JJS_ERR_CUDA, "sign batch failed: %s", cudaGetErrorString(e));
    }
    cudaFree(buf);
    return rc;
}

int init_device(jjs_ctx* ctx, DeviceState& d) {
    JJS_CUDA(ctx, cudaSetDevice(d.device));
    cudaDeviceProp prop;
    JJS_CUDA(ctx, cudaGetDeviceProperties(&prop, d.device));
    if (prop.major != 10) return fail(ctx, JJS_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", d.device, prop.major, prop.minor);
    JJS_CUDA(ctx, cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking));
    JJS_CUDA(ctx, cudaStreamCreateWithFlags(&d.copy_stream, cudaStreamNonBlocking));
    JJS_CUDA(ctx, cudaEventCreateWithFlags(&d.copied, cudaEventDisableTiming));
    JJS_CUDA(ctx, cudaEventCreateWithFlags(&d.fork, cudaEventDisableTiming));
    for (int k = 0; k < 2; k++) {
        JJS_CUDA(ctx, cudaStreamCreateWithFlags(&d.sub[k], cudaStreamNonBlocking));
        JJS_CUDA(ctx, cudaEventCreateWithFlags(&d.join[k], cudaEventDisableTiming));
    }
    JJS_CUDA(ctx, cudaMalloc(&d.root_tables, sizeof(tables::ROOT_TABLES)));
    JJS_CUDA(ctx, cudaMemcpy(d.root_tables, tables::ROOT_TABLES, sizeof(tables::ROOT_TABLES), cudaMemcpyHostToDevice));
    JJS_CUDA(ctx, cudaMalloc(&d.dlog_hash, sizeof(tables::DLOG_HASH)));
    JJS_CUDA(ctx, cudaMemcpy(d.dlog_hash, tables::DLOG_HASH, sizeof(tables::DLOG_HASH), cudaMemcpyHostToDevice));
    const size_t fb_bytes = sizeof(niels) * FB_WINDOWS * FB_ENTRIES;
    JJS_CUDA(ctx, cudaMalloc(&d.fb_g, fb_bytes));
    JJS_CUDA(ctx, cudaMalloc(&d.fb_gn, fb_bytes));
    k_fb_table<<<blocks_for(FB_WINDOWS * FB_ENTRIES), BLOCK, 0, d.stream>>>(d.fb_g, 0);
    k_fb_table<<<blocks_for(FB_WINDOWS * FB_ENTRIES), BLOCK, 0, d.stream>>>(d.fb_gn, 1);
    ctx->launches += 2;
    JJS_CUDA(ctx, cudaGetLastError());
    JJS_CUDA(ctx, cudaStreamSynchronize(d.stream));
    return JJS_SUCCESS;
}

void free_device(DeviceState& d) {
    if (d.device < 0) return;
    cudaSetDevice(d.device);
    cudaFree(d.root_tables); cudaFree(d.dlog_hash); cudaFree(d.fb_g); cudaFree(d.fb_gn);
    cudaFree(d.pts_u); cudaFree(d.pts_v); cudaFree(d.tab); cudaFree(d.pflags); cudaFree(d.iflags); cudaFree(d.eqflags); cudaFree(d.cwords); cudaFree(d.rlist); cudaFree(d.rcount); cudaFree(d.eqlist);
    cudaFree(d.s_pk); cudaFree(d.s_sig); cudaFree(d.s_msg); cudaFree(d.s_status); cudaFree(d.s_c); cudaFree(d.s_bitmap);
    cudaFree(d.keys_u); cudaFree(d.keys_v); cudaFree(d.kflags); cudaFree(d.kcoef); cudaFree(d.agg_stage); cudaFree(d.d_order);
    if (d.stream) cudaStreamDestroy(d.stream);
    if (d.copy_stream) cudaStreamDestroy(d.copy_stream);
    if (d.copied) cudaEventDestroy(d.copied);
    if (d.fork) cudaEventDestroy(d.fork);
    for (int k = 0; k < 2; k++) {
        if (d.sub[k]) cudaStreamDestroy(d.sub[k]);
        if (d.join[k]) cudaEventDestroy(d.join[k]);
    }
}

int device_entry(jjs_ctx* ctx, int variant, int device_index, const uint8_t* pk, const uint8_t* sig, const uint8_t* msg, size_t n,
                 uint8_t* status, uint8_t* c_out, void* stream) {
    if (!ctx) return JJS_ERR_ARGUMENT;
    ctx->err[0] = 0;
    if (device_index < 0 || (size_t)device_index >= ctx->dev.size()) return fail(ctx, JJS_ERR_ARGUMENT, "device_index out of range");
    if (n == 0) return JJS_SUCCESS;
    if (!pk || !sig || !msg || !status) return fail(ctx, JJS_ERR_ARGUMENT, "null buffer");
    DeviceState& d = ctx->dev[device_index];
    return run_device(ctx, d, variant, pk, sig, msg, n, status, c_out, (cudaStream_t)stream);
}

}  // namespace

extern "C" {

#define JJS_API __attribute__((visibility("default")))

JJS_API int jjs_init(const int* devices, int n_devices, jjs_ctx** out) {
    if (!out || n_devices < 1 || n_devices > 64) return JJS_ERR_ARGUMENT;
    *out = nullptr;
    jjs_ctx* ctx = new (std::nothrow) jjs_ctx();
    if (!ctx) return JJS_ERR_NOMEM;
    ctx->err[0] = 0;
    ctx->launches = 0;
    ctx->profile = false;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count < 1) {
        // no CPU fallback: the context is returned unusable only so the caller can read the reason
        fail(ctx, JJS_ERR_CUDA, "no CUDA device available: %s", e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        *out = ctx;
        return JJS_ERR_CUDA;
    }
    ctx->dev.resize(n_devices);
    for (int i = 0; i < n_devices; i++) {
        ctx->dev[i].device = devices ? devices[i] : i;
        if (ctx->dev[i].device < 0 || ctx->dev[i].device >= count) {
            fail(ctx, JJS_ERR_ARGUMENT, "device ordinal %d not present (%d devices)", ctx->dev[i].device, count);
            ctx->dev[i].device = -1;
            *out = ctx;
            return JJS_ERR_ARGUMENT;
        }
        int rc = init_device(ctx, ctx->dev[i]);
        if (rc) {
            *out = ctx;
            return rc;
        }
    }
    *out = ctx;
    return JJS_SUCCESS;
}

JJS_API void jjs_destroy(jjs_ctx* ctx) {
    if (!ctx) return;
    for (auto& d : ctx->dev) free_device(d);
    delete ctx;
}

JJS_API const char* jjs_last_error(const jjs_ctx* ctx) { return ctx ? ctx->err : "null context"; }
JJS_API int jjs_device_count(const jjs_ctx* ctx) { return ctx ? (int)ctx->dev.size() : 0; }
JJS_API uint64_t jjs_launch_count(const jjs_ctx* ctx) { return ctx ? ctx->launches : 0; }

JJS_API int jjs_verify_single(jjs_ctx* ctx, const uint8_t* pk32, const uint8_t* sig64, const uint8_t* msg32, size_t n, uint8_t* status,
                              uint8_t* c32_or_null) {
    return run_host(ctx, VAR_SINGLE, pk32, sig64, msg32, n, status, c32_or_null);
}
JJS_API int jjs_verify_batch(jjs_ctx* ctx, const uint8_t* pk32, const uint8_t* sig64, const uint8_t* msg32, size_t n, uint32_t* accept_bitmap) {
    if (ctx && n && !accept_bitmap) return fail(ctx, JJS_ERR_ARGUMENT, "null buffer");
    return run_host(ctx, VAR_SINGLE, pk32, sig64, msg32, n, nullptr, nullptr, false, accept_bitmap);
}
JJS_API int jjs_status_bitmap_device(jjs_ctx* ctx, int device_index, const uint8_t* d_status, size_t n, uint32_t* d_accept_bitmap, void* cuda_stream) {
    if (!ctx) return JJS_ERR_ARGUMENT;
    ctx->err[0] = 0;
    if (device_index < 0 || (size_t)device_index >= ctx->dev.size()) return fail(ctx, JJS_ERR_ARGUMENT, "device_index out of range");
    if (n == 0) return JJS_SUCCESS;
    if (!d_status || !d_accept_bitmap) return fail(ctx, JJS_ERR_ARGUMENT, "null buffer");
    JJS_CUDA(ctx, cudaSetDevice(ctx->dev[device_index].device));
    k_bitmap<<<blocks_for(n), BLOCK, 0, (cudaStream_t)cuda_stream>>>(d_status, n, d_accept_bitmap);
    ctx->launches++;
    JJS_CUDA(ctx, cudaGetLastError());
    return JJS_SUCCESS;
}
JJS_API int jjs_verify_double(jjs_ctx* ctx, const uint8_t* pk64, const uint8_t* sig96, const uint8_t* msg32, size_t n, uint8_t* status,
                              uint8_t* c32_or_null) {
    return run_host(ctx, VAR_DOUBLE, pk64, sig96, msg32, n, status, c32_or_null);
}
JJS_API int jjs_verify_vargen(jjs_ctx* ctx, const uint8_t* pk64, const uint8_t* sig64, const uint8_t* msg32, size_t n, uint8_t* status,
                              uint8_t* c32_or_null) {
    return run_host(ctx, VAR_VARGEN, pk64, sig64, msg32, n, status, c32_or_null);
}
JJS_API int jjs_verify_aggregate(jjs_ctx* ctx, const uint8_t* pks32, const uint32_t* offsets, const uint8_t* sig64, const uint8_t* msg32, size_t n,
                                 uint8_t* status, uint8_t* c32_or_null, uint8_t* aggpk32_or_null) {
    return run_aggregate(ctx, pks32, offsets, sig64, msg32, n, status, c32_or_null, aggpk32_or_null);
}
JJS_API int jjs_verify_single_device(jjs_ctx* ctx, int device_index, const uint8_t* d_pk32, const uint8_t* d_sig64, const uint8_t* d_msg32,
                                     size_t n, uint8_t* d_status, uint8_t* d_c32_or_null, void* cuda_stream) {
    return device_entry(ctx, VAR_SINGLE, device_index, d_pk32, d_sig64, d_msg32, n, d_status, d_c32_or_null, cuda_stream);
}
JJS_API int jjs_verify_double_device(jjs_ctx* ctx, int device_index, const uint8_t* d_pk64, const uint8_t* d_sig96, const uint8_t* d_msg32,
                                     size_t n, uint8_t* d_status, uint8_t* d_c32_or_null, void* cuda_stream) {
    return device_entry(ctx, VAR_DOUBLE, device_index, d_pk64, d_sig96, d_msg32, n, d_status, d_c32_or_null, cuda_stream);
}
JJS_API int jjs_verify_vargen_device(jjs_ctx* ctx, int device_index, const uint8_t* d_pk64, const uint8_t* d_sig64, const uint8_t* d_msg32,
                                     size_t n, uint8_t* d_status, uint8_t* d_c32_or_null, void* cuda_stream) {
    return device_entry(ctx, VAR_VARGEN, device_index, d_pk64, d_sig64, d_msg32, n, d_status, d_c32_or_null, cuda_stream);
}
JJS_API int jjs_subgroup_check(jjs_ctx* ctx, const uint8_t* points32, size_t n, int method, uint8_t* out) {
    if (!ctx) return JJS_ERR_ARGUMENT;
    ctx->err[0] = 0;
    if (method < 0 || method > 1) return fail(ctx, JJS_ERR_ARGUMENT, "bad method");
    if (n == 0) return JJS_SUCCESS;
    if (!points32 || !out) return fail(ctx, JJS_ERR_ARGUMENT, "null buffer");
    DeviceState& d = ctx->dev[0];
    int rc = ensure_scratch(ctx, d);
    if (rc) return rc;
    JJS_CUDA(ctx, cudaSetDevice(d.device));
    uint8_t *d_in = nullptr, *d_out = nullptr;
    JJS_CUDA(ctx, cudaMalloc(&d_in, 32 * n));
    JJS_CUDA(ctx, cudaMalloc(&d_out, n));
    cudaMemcpyAsync(d_in, points32, 32 * n, cudaMemcpyHostToDevice, d.stream);
    for (size_t first = 0; first < n; first += TAB_THREADS) {
        size_t cnt = n - first < TAB_THREADS ? n - first : TAB_THREADS;
        k_subgroup_check<<<blocks_for(cnt), BLOCK, 0, d.stream>>>(WireField{d_in, 32}, first, cnt, method, d_out, d.tab, TAB_THREADS, d.tables());
        ctx->launches++;
    }
    cudaMemcpyAsync(out, d_out, n, cudaMemcpyDeviceToHost, d.stream);
    cudaError_t e = cudaStreamSynchronize(d.stream);
    cudaFree(d_in);
    cudaFree(d_out);
    if (e != cudaSuccess) return fail(ctx, JJS_ERR_CUDA, "subgroup check failed: %s", cudaGetErrorString(e));
    return JJS_SUCCESS;
}
JJS_API int jjs_verify_aggregate_device(jjs_ctx* ctx, int device_index, const uint8_t* d_pks32, const uint32_t* d_offsets, const uint32_t* h_offsets,
                                        const uint8_t* d_sig64, const uint8_t* d_msg32, size_t n, uint8_t* d_status, uint8_t* d_c32_or_null,
                                        uint8_t* d_aggpk32_or_null, void* cuda_stream) {
    if (!ctx) return JJS_ERR_ARGUMENT;
    ctx->err[0] = 0;
    if (device_index < 0 || (size_t)device_index >= ctx->dev.size()) return fail(ctx, JJS_ERR_ARGUMENT, "device_index out of range");
    if (n == 0) return JJS_SUCCESS;
    if (!d_offsets || !h_offsets || !d_sig64 || !d_msg32 || !d_status) return fail(ctx, JJS_ERR_ARGUMENT, "null buffer");
    return run_aggregate_device(ctx, ctx->dev[device_index], d_pks32, d_offsets, h_offsets, d_sig64, d_msg32, n, d_status, d_c32_or_null,
                                d_aggpk32_or_null, (cudaStream_t)cuda_stream);
}
JJS_API int jjs_sign_aggregate_batch(jjs_ctx* ctx, const uint8_t* sk32, const uint32_t* offsets, const uint8_t* rnd32, const uint8_t* msg32, size_t n,
                                     uint8_t* pks32_out, uint8_t* sig64_out) {
    if (!ctx) return JJS_ERR_ARGUMENT;
    ctx->err[0] = 0;
    if (n == 0) return JJS_SUCCESS;
    if (!sk32 || !offsets || !rnd32 || !msg32 || !pks32_out || !sig64_out) return fail(ctx, JJS_ERR_ARGUMENT, "null buffer");
    if (offsets[0] != 0) return fail(ctx, JJS_ERR_ARGUMENT, "offsets[0] must be 0");
    DeviceState& d = ctx->dev[0];
    JJS_CUDA(ctx, cudaSetDevice(d.device));
    const size_t K = offsets[n];
    uint8_t* b = nullptr;
    size_t o_off = 32 * K, o_rnd = o_off + ((4 * (n + 1) + 31) / 32) * 32, o_msg = o_rnd + 32 * n, o_pks = o_msg + 32 * n, o_sig = o_pks + 32 * K,
           total = o_sig + 64 * n;
    JJS_CUDA(ctx, cudaMalloc(&b, total));
    cudaMemcpyAsync(b, sk32, 32 * K, cudaMemcpyHostToDevice, d.stream);
    cudaMemcpyAsync(b + o_off, offsets, 4 * (n + 1), cudaMemcpyHostToDevice, d.stream);
    cudaMemcpyAsync(b + o_rnd, rnd32, 32 * n, cudaMemcpyHostToDevice, d.stream);
    cudaMemcpyAsync(b + o_msg, msg32, 32 * n, cudaMemcpyHostToDevice, d.stream);
    k_sign_aggregate<<<blocks_for(n), BLOCK, 0, d.stream>>>(b, reinterpret_cast<uint32_t*>(b + o_off), b + o_rnd, b + o_msg, n, b + o_pks, b + o_sig,
                                                          d.tables());
    ctx->launches++;
    cudaMemcpyAsync(pks32_out, b + o_pks, 32 * K, cudaMemcpyDeviceToHost, d.stream);
    cudaMemcpyAsync(sig64_out, b + o_sig, 64 * n, cudaMemcpyDeviceToHost, d.stream);
    cudaError_t e = cudaStreamSynchronize(d.stream);
    cudaFree(b);
    if (e != cudaSuccess) return fail(ctx, JJS_ERR_CUDA, "sign aggregate batch failed: %s", cudaGetErrorString(e));
    return JJS_SUCCESS;
}
JJS_API int jjs_verify_ext(jjs_ctx* ctx, int variant, const uint8_t* points_ext160, const uint8_t* u32, const uint8_t* msg32, size_t n,
                           uint8_t* status, uint8_t* c32_or_null) {
    return run_ext(ctx, variant, points_ext160, u32, msg32, n, status, c32_or_null);
}
JJS_API int jjs_points_to_ext(jjs_ctx* ctx, const uint8_t* points32, const uint8_t* z_mont32, size_t n, uint8_t* out160) {
    if (!ctx) return JJS_ERR_ARGUMENT;
    ctx->err[0] = 0;
    if (n == 0) return JJS_SUCCESS;
    if (!points32 || !z_mont32 || !out160) return fail(ctx, JJS_ERR_ARGUMENT, "null buffer");
    DeviceState& d = ctx->dev[0];
    JJS_CUDA(ctx, cudaSetDevice(d.device));
    uint8_t* b = nullptr;
    JJS_CUDA(ctx, cudaMalloc(&b, n * (32 + 32 + 160)));
    cudaMemcpyAsync(b, points32, 32 * n, cudaMemcpyHostToDevice, d.stream);
    cudaMemcpyAsync(b + 32 * n, z_mont32, 32 * n, cudaMemcpyHostToDevice, d.stream);
    k_points_to_ext<<<blocks_for(n), BLOCK, 0, d.stream>>>(b, b + 32 * n, n, b + 64 * n, d.tables());
    ctx->launches++;
    cudaMemcpyAsync(out160, b + 64 * n, 160 * n, cudaMemcpyDeviceToHost, d.stream);
    cudaError_t e = cudaStreamSynchronize(d.stream);
    cudaFree(b);
    if (e != cudaSuccess) return fail(ctx, JJS_ERR_CUDA, "points_to_ext failed: %s", cudaGetErrorString(e));
    return JJS_SUCCESS;
}
JJS_API int jjs_multisig_combine(jjs_ctx* ctx, const uint8_t* pks32, const uint8_t* R32, const uint8_t* S32, const uint8_t* z32, const uint32_t* offsets,
                                 const uint8_t* msg32, size_t n, uint8_t* share_ok_or_null, uint8_t* status, uint32_t* bad_index_or_null,
                                 uint8_t* sig64_or_null) {
    return run_msig(ctx, pks32, R32, S32, z32, offsets, msg32, n, share_ok_or_null, status, bad_index_or_null, sig64_or_null);
}
JJS_API void jjs_profile_enable(jjs_ctx* ctx, int on) {
    if (ctx) ctx->profile = on != 0;
}
JJS_API int jjs_profile_collect(jjs_ctx* ctx, double* stage_ms, uint64_t* stage_count) {
    if (!ctx || !stage_ms || !stage_count) return JJS_ERR_ARGUMENT;
    for (int i = 0; i < JJS_N_STAGES; i++) { stage_ms[i] = 0; stage_count[i] = 0; }
    for (auto& r : ctx->records) {
        cudaSetDevice(r.device);
        float ms = 0;
        cudaError_t e = cudaEventSynchronize(r.e1);
        if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, r.e0, r.e1);
        cudaEventDestroy(r.e0);
        cudaEventDestroy(r.e1);
        if (e != cudaSuccess) {
            ctx->records.clear();
            return fail(ctx, JJS_ERR_CUDA, "profile collect: %s", cudaGetErrorString(e));
        }
        stage_ms[r.stage] += ms;
        stage_count[r.stage]++;
    }
    ctx->records.clear();
    return JJS_SUCCESS;
}
JJS_API int jjs_sign_batch(jjs_ctx* ctx, int variant, const uint8_t* sk32, const uint8_t* rnd32, const uint8_t* gen_scalar32_or_null,
                           const uint8_t* msg32, size_t n, uint8_t* pk_out, uint8_t* sig_out) {
    return run_sign(ctx, variant, sk32, rnd32, gen_scalar32_or_null, msg32, n, pk_out, sig_out);
}
JJS_API int jjs_challenge_only(jjs_ctx* ctx, int variant, const uint8_t* pk, const uint8_t* sig, const uint8_t* msg32, size_t n, uint8_t* c32) {
    if (variant < 0 || variant > 2) return fail(ctx, JJS_ERR_ARGUMENT, "bad variant");
    return run_host(ctx, variant, pk, sig, msg32, n, nullptr, c32, true);
}

}  // extern "C"
