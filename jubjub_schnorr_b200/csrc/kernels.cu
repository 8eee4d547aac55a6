// CUDA kernels (sm_100a) and the C ABI of libjjschnorr_b200.so.  See include/jjschnorr_b200.h for the
// contract and verify_core.cuh for the per-item stages.  One thread owns one point (decode, subgroup
// check), one item (challenge sponge, final status) or one verification equation (Straus-style
// u*B + c*PK); intermediates travel between the stages as SoA arrays of 32-byte field elements in HBM,
// so every warp access is a run of 128-bit vector loads.  There is deliberately no CPU path.
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include "../../include/jjschnorr_b200.h"
#include "multisig_core.cuh"
#include "sign_core.cuh"
#include "verify_core.cuh"
#include "safe_tag.h"

// The slot-based evaluation of the verification equation (csrc/fqs.cuh: field elements in shared memory, persistent kernel) is a
// compile-time alternative.  Measured on B200: it executes 6 % fewer multiply-pipe cycles and runs 4-7 %
// slower than the register-operand kernel, so it is off; tests/hostsim keeps its twin under test either way.
#ifndef JJS_EQ_V2
#define JJS_EQ_V2 0
#endif
#if JJS_EQ_V2
#include "fqs.cuh"
#else
#ifndef JJS_EQ_BLOCK
#define JJS_EQ_BLOCK 128
#endif
#endif

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
#define JJS_EQ_MINBLOCKS 4    // fixed-base equation kernel: 4 resident CTAs of 128 threads (128 registers, a few spills) beat 3 by 4 %
#endif
#ifndef JJS_EQ1_MINBLOCKS
#define JJS_EQ1_MINBLOCKS 4   // var-generator equation kernel
#endif
#ifndef JJS_AGG_MINBLOCKS
#define JJS_AGG_MINBLOCKS 4   // key-aggregation kernel (-1.9 % against 3)
#endif
#ifndef JJS_DEC_MINBLOCKS
#define JJS_DEC_MINBLOCKS 8   // decode and deferred-test kernels: 4 -> 12.62 ms per 2^21 points, 6 -> 12.49, 8 -> 12.41
#endif
#ifndef JJS_HASH_MINBLOCKS
#define JJS_HASH_MINBLOCKS 6  // hash kernels: 4 -> 12.78 ms per 2^20 challenges, 6 -> 12.66, 7 -> 12.92, 8 -> 13.06
#endif
#ifndef JJS_EQ_PERSISTENT
#define JJS_EQ_PERSISTENT 0   // measured (gpurun_out/ab2.log): -22 % DRAM reads but +6 % time; off
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

// ---- aggregate-key path: ordering of a chunk by signer count, entirely on the device -----------------------------------
// Items are bucketed by signer count (bucket 64 collects everything above 63) so that the lanes of a warp hash transcripts
// of one length and fold the same number of keys.  Three small kernels replace the host-side counting sort of round 1 (and
// the stream synchronisation it needed): histogram, exclusive scan over the 65 buckets, scatter.  The order inside a bucket
// depends on atomics; results are written by item index, so the outputs do not.
constexpr int AGG_BUCKETS = 65;
struct AggSort {
    uint32_t hist[AGG_BUCKETS], khist[AGG_BUCKETS];    // items / keys per bucket
    uint32_t ibase[AGG_BUCKETS], kbase[AGG_BUCKETS];   // first sorted item / key position of the bucket
    uint32_t cursor[AGG_BUCKETS], kcursor;             // scatter cursors (kcursor: keys of the open-ended bucket)
};
__global__ void __launch_bounds__(BLOCK) k_agg_hist(const uint32_t* offsets, size_t n, AggSort* c) {
    __shared__ uint32_t h[AGG_BUCKETS], kh[AGG_BUCKETS];
    for (int i = threadIdx.x; i < AGG_BUCKETS; i += blockDim.x) { h[i] = 0; kh[i] = 0; }
    __syncthreads();
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        uint32_t cnt = offsets[i + 1] - offsets[i], b = cnt < AGG_BUCKETS - 1 ? cnt : AGG_BUCKETS - 1;
        atomicAdd(&h[b], 1u);
        atomicAdd(&kh[b], cnt);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < AGG_BUCKETS; k += blockDim.x) {
        if (h[k]) atomicAdd(&c->hist[k], h[k]);
        if (kh[k]) atomicAdd(&c->khist[k], kh[k]);
    }
}
__global__ void k_agg_scan(AggSort* c) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    uint32_t ib = 0, kb = 0;
    for (int b = 0; b < AGG_BUCKETS; b++) {
        c->ibase[b] = ib;
        c->kbase[b] = kb;
        ib += c->hist[b];
        kb += c->khist[b];
    }
}
// order[p]: item at sorted position p;  kmap[q] / kitem[q]: key (relative to the chunk's first key) and item at sorted key position q
__global__ void __launch_bounds__(BLOCK) k_agg_scatter(const uint32_t* offsets, uint32_t key_base, size_t n, AggSort* c, uint32_t* order, uint32_t* kmap,
                                                       uint32_t* kitem) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = i < n;
    uint32_t lo = 0, cnt = 0, b = 0xffffffffu;
    if (valid) {
        lo = offsets[i];
        cnt = offsets[i + 1] - lo;
        b = cnt < AGG_BUCKETS - 1 ? cnt : AGG_BUCKETS - 1;
    }
    // one atomic per distinct bucket of the warp
    const int lane = threadIdx.x & 31;
    uint32_t pos = 0;
    uint32_t remaining = __ballot_sync(0xffffffffu, valid);
    while (remaining) {
        int leader = __ffs(remaining) - 1;
        uint32_t bb = __shfl_sync(0xffffffffu, b, leader);
        uint32_t grp = __ballot_sync(0xffffffffu, valid && b == bb);
        uint32_t base = 0;
        if (lane == leader) base = atomicAdd(&c->cursor[bb], (uint32_t)__popc(grp));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (valid && b == bb) pos = base + __popc(grp & ((1u << lane) - 1u));
        remaining &= ~grp;
    }
    if (!valid) return;
    order[c->ibase[b] + pos] = (uint32_t)i;
    uint32_t kpos = b < AGG_BUCKETS - 1 ? c->kbase[b] + pos * cnt : c->kbase[b] + atomicAdd(&c->kcursor, cnt);
    for (uint32_t j = 0; j < cnt; j++) {
        kmap[kpos + j] = lo - key_base + j;
        kitem[kpos + j] = (uint32_t)i;
    }
}

// one thread hashes the delinearisation coefficient of ONE signer key (sorted key order: the threads of a warp hash transcripts of
// one length).  Per key rather than per item: an item with n signers costs n hashes of 2 + 2 n elements, and spread over n
// threads that work no longer depends on the signer count of the neighbouring items.
__global__ void __launch_bounds__(BLOCK, JJS_HASH_MINBLOCKS) k_agg_coeffs(const fq* keys_u, const fq* keys_v, const uint8_t* kflags,
                                                                          const uint32_t* offsets, const uint32_t* kmap, const uint32_t* kitem,
                                                                          uint32_t key_base, size_t n_keys, uint32_t* d_words, Tables T) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_keys) return;
    size_t item = kitem[t];
    stage_aggregate_coeff_key(keys_u, keys_v, kflags, offsets[item] - key_base, offsets[item + 1] - key_base, kmap[t], d_words, T.safe_tags);
}

// one thread folds the signer keys of one item into its aggregate key (slot 0 of the single-variant point arrays)
// `order` lists the items sorted by signer count, so the lanes of a warp loop over the same number of signers
__global__ void __launch_bounds__(BLOCK, JJS_AGG_MINBLOCKS) k_aggregate(const fq* keys_u, const fq* keys_v, const uint8_t* kflags, const uint32_t* offsets,
                                                     const uint32_t* order, uint32_t key_base, size_t n, fq* pts_u, fq* pts_v, uint8_t* pflags,
                                                     uint8_t* agg_out, fq* tab, size_t stride, uint32_t* d_words, Tables T) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    size_t item = order[t];
    uint32_t w[8];
    stage_aggregate(keys_u, keys_v, kflags, offsets[item] - key_base, offsets[item + 1] - key_base, pts_u, pts_v, pflags, item, w, tab + t, stride,
                    T.safe_tags, d_words);
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
__global__ void __launch_bounds__(BLOCK, JJS_HASH_MINBLOCKS) k_challenge(int variant, const fq* pts_u, const fq* pts_v, size_t n, WireField msg,
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
// (instantiated per kind of equation: MODE 0 fixed base -- single, double -- and MODE 1 variable base -- var-gen)
template <int MODE>
__global__ void __launch_bounds__(BLOCK, MODE ? JJS_EQ1_MINBLOCKS : JJS_EQ_MINBLOCKS) k_equation(int variant, const fq* pts_u, const fq* pts_v, uint8_t* pflags, size_t n,
                                                    size_t first, size_t count, const uint32_t* list, const uint32_t* lcount, WireField usc,
                                                    const uint32_t* cwords, uint8_t* eqflags, fq* tab, size_t stride, Tables T,
                                                    uint32_t* rlist, uint32_t* rcount) {
    // Persistent form (JJS_EQ_PERSISTENT): the grid is one wave of resident CTAs and every thread walks the work list with the
    // grid's stride, so the two per-thread tables live in a scratch indexed by RESIDENT thread (stride = gridDim.x * blockDim.x,
    // ~130 MB in all, close to the L2's size) instead of by equation (600 MB per 2^18-item sub-chunk).
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int neq = variant == VAR_DOUBLE ? 2 : 1;
    const size_t total = (size_t)neq * *lcount, end = first + count < total ? first + count : total;
#if JJS_EQ_PERSISTENT
    const size_t step = (size_t)gridDim.x * blockDim.x;
#pragma unroll 1
    for (size_t g = first + t; g < end; g += step) {
#else
    {
        const size_t g = first + t;
        if (t >= count || g >= end) return;
#endif
        int eq = (int)(g % neq);
        size_t item = list[g / neq];
        bool need_r_test;
        bool ok = stage_equation_item<MODE>(variant, eq, pts_u, pts_v, pflags, n, item, (variant == VAR_DOUBLE && eq == 1) ? T.fb_gn : T.fb_g, usc, cwords,
                                            tab + t, tab + 36 * stride + t, stride, &need_r_test);
        if (need_r_test) {
            int pk_slot, r_slot, base_slot;
            equation_slots(variant, eq, pk_slot, r_slot, base_slot);
            rlist[atomicAdd(rcount, 1u)] = (uint32_t)((size_t)r_slot * n + item);
        }
        eqflags[(size_t)eq * n + item] = ok ? 1 : 0;
    }
}

// The same stage on the slot-based evaluation (fqs.cuh): field elements in shared memory, addressed by handle, so that no
// product needs its operands marshalled through registers.  Persistent: the grid is one wave of resident CTAs and every
// thread walks the work list with the grid's stride, which lets the two per-thread tables live in a scratch indexed by
// RESIDENT thread (gridDim.x * blockDim.x threads x 2 KiB, L2 sized) instead of by item.
#ifndef JJS_EQ2_MINBLOCKS
#define JJS_EQ2_MINBLOCKS 4
#endif
#if JJS_EQ_V2
__global__ void __launch_bounds__(JJS_EQ_BLOCK, JJS_EQ2_MINBLOCKS) k_equation2(int variant, const fq* pts_u, const fq* pts_v, uint8_t* pflags, size_t n,
                                                                                const uint32_t* list, const uint32_t* lcount, WireField usc,
                                                                                const uint32_t* cwords, uint8_t* eqflags, fq* tab, Tables T,
                                                                                uint32_t* rlist, uint32_t* rcount) {
    const Eq2Slots s = eq2_slots(threadIdx.x);
    const size_t nthreads = (size_t)gridDim.x * blockDim.x, tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int neq = variant == VAR_DOUBLE ? 2 : 1;
    const size_t total = (size_t)neq * *lcount;
#pragma unroll 1
    for (size_t g = tid; g < total; g += nthreads) {
        int eq = (int)(g % neq);
        size_t item = list[g / neq];
        bool need_r_test;
        bool ok = eq2_equation_item(s, variant, eq, pts_u, pts_v, pflags, n, item, (variant == VAR_DOUBLE && eq == 1) ? T.fb_gn : T.fb_g, usc, cwords,
                                    tab + tid, tab + EQ2_TAB_FQ * nthreads + tid, nthreads, &need_r_test);
        if (need_r_test) {
            int pk_slot, r_slot, base_slot;
            equation_slots(variant, eq, pk_slot, r_slot, base_slot);
            rlist[atomicAdd(rcount, 1u)] = (uint32_t)((size_t)r_slot * n + item);
        }
        eqflags[(size_t)eq * n + item] = ok ? 1 : 0;
    }
}

#endif  // JJS_EQ_V2

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

// Fixed-base tables in two passes (curve.cuh, fb_combine_entries).  Pass 1: the small tables of every window, affine points,
// small[w][0 .. 2^FB_LO) = jl 2^(FB_W w) B  and  small[w][2^FB_LO .. 2^FB_LO + 2^FB_HI) = jh 2^(FB_W w + FB_LO) B.
constexpr int FB_SMALL = (1 << FB_LO) + (1 << FB_HI);
__global__ void __launch_bounds__(BLOCK) k_fb_small(fq* small_u, fq* small_v, int which) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= FB_WINDOWS * FB_SMALL) return;
    fq bu, bv, u, v;
    fq_load_const(bu, which ? JJS_C(GEN_NUMS_UV)[0] : JJS_C(GEN_UV)[0]);
    fq_load_const(bv, which ? JJS_C(GEN_NUMS_UV)[1] : JJS_C(GEN_UV)[1]);
    int w = t / FB_SMALL, e = t % FB_SMALL;
    if (e < (1 << FB_LO)) fb_affine_multiple(u, v, bu, bv, w * FB_W, e, FB_LO);
    else fb_affine_multiple(u, v, bu, bv, w * FB_W + FB_LO, e - (1 << FB_LO), FB_HI);
    small_u[t] = u;
    small_v[t] = v;
}
// Pass 2: thread t combines entries [FB_BATCH t, FB_BATCH (t + 1)) of the table (FB_BATCH divides 2^FB_LO, so they share jh).
__global__ void __launch_bounds__(BLOCK) k_fb_combine(niels* out, const fq* small_u, const fq* small_v) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)FB_WINDOWS * FB_ENTRIES / FB_BATCH) return;
    size_t first = t * FB_BATCH;
    int w = (int)(first / FB_ENTRIES), j = (int)(first % FB_ENTRIES), jh = j >> FB_LO, jl = j & ((1 << FB_LO) - 1);
    const fq* su = small_u + (size_t)w * FB_SMALL;
    const fq* sv = small_v + (size_t)w * FB_SMALL;
    fb_combine_entries(out + first, su + jl, sv + jl, su[(1 << FB_LO) + jh], sv[(1 << FB_LO) + jh]);
}
// the definition, entry by entry (jjs_fb_table_check compares a sample of the built tables with it)
__global__ void __launch_bounds__(BLOCK) k_fb_check(const niels* table, int which, const uint32_t* entries, int n, uint32_t* mismatches) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    fq u, v;
    fq_load_const(u, which ? JJS_C(GEN_NUMS_UV)[0] : JJS_C(GEN_UV)[0]);
    fq_load_const(v, which ? JJS_C(GEN_NUMS_UV)[1] : JJS_C(GEN_UV)[1]);
    niels e;
    uint32_t idx = entries[t] % (uint32_t)(FB_WINDOWS * FB_ENTRIES);
    fb_table_entry(e, u, v, (int)(idx / FB_ENTRIES), (int)(idx % FB_ENTRIES));
    const niels& g = table[idx];
    if (!(fq_eq(e.ypx, g.ypx) && fq_eq(e.ymx, g.ymx) && fq_eq(e.t2d, g.t2d))) atomicAdd(mismatches, 1u);
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

__global__ void __launch_bounds__(BLOCK) k_msig_session(MsigBuffers b, size_t K, size_t n, WireField msg, WireField zf, fq* tab, size_t stride, Tables T) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    size_t s = b.order[t];
    stage_msig_session(b.pu, b.pv, b.pf, K, b.offsets[s], b.offsets[s + 1], msg, zf, s, b.d_words, b.cd_words, b.a_words, b.rsa_u, b.rsa_v, b.sflags,
                       tab + t, stride, T.safe_tags);
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

struct AggScratch {   // aggregate-key path, one per scratch half: decoded signer keys, their coefficients, the sorted orders
    fq *keys_u = nullptr, *keys_v = nullptr;
    uint8_t* kflags = nullptr;
    uint32_t *kcoef = nullptr, *kmap = nullptr, *kitem = nullptr, *order = nullptr;
    AggSort* sort = nullptr;
    size_t cap_keys = 0;
};

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
    fq* safe_tags = nullptr;     // SAFE tags by number of absorbed elements (multisig transcripts of run-time length)
    size_t n_tags = 0;
    // pipeline scratch, sized for CHUNK_ITEMS items of the widest variant (4 points, 2 equations)
    fq *pts_u = nullptr, *pts_v = nullptr, *tab = nullptr;
    uint8_t *pflags = nullptr, *iflags = nullptr, *eqflags = nullptr;
    uint32_t* cwords = nullptr;
    uint32_t *rlist = nullptr, *rcount = nullptr;  // signature points awaiting the deferred subgroup test
    uint32_t* eqlist = nullptr;                    // items that go on to the hash and equation kernels (k_work_list)
    // persistent equation kernel: one wave of resident CTAs, two per-thread tables per resident thread, one such scratch
    // per compute stream (the two sub-chunk streams may run their equation kernels back to back or side by side)
    fq* eqtab = nullptr;
    size_t eqtab_half = 0;         // elements per stream's scratch
    int eq_grid = 0;
    AggScratch agg[2];
    // staging of the host-buffer entry points: one grow-only device buffer, carved up per call
    uint8_t* stage = nullptr;
    size_t stage_bytes = 0;
    Tables tables() const { return Tables{root_tables, dlog_hash, fb_g, fb_gn, safe_tags}; }
};

// A slice of the pipeline scratch able to hold `cap` items of the widest variant; `tab` serves `cap` threads (the table
// scratch keeps its global stride TAB_THREADS).
struct Region {
    fq *pts_u, *pts_v, *tab, *eqtab;
    uint8_t *pflags, *iflags, *eqflags;
    uint32_t *cwords, *rlist, *rcount, *eqlist;   // rcount[0]: deferred subgroup tests, rcount[2]: length of the work list
    size_t cap;
    int half;
};
inline Region region_of(const DeviceState& d, size_t first_item, size_t cap, int half) {
    return Region{d.pts_u + 4 * first_item, d.pts_v + 4 * first_item, d.tab + first_item,
                  d.eqtab + (size_t)(half & 1) * d.eqtab_half, d.pflags + 4 * first_item, d.iflags + first_item,
                  d.eqflags + 2 * first_item, d.cwords + 8 * first_item, d.rlist + 2 * first_item, d.rcount + 4 * (half & 1), d.eqlist + 2 * first_item, cap,
                  half & 1};
}
constexpr size_t SUB_ITEMS = CHUNK_ITEMS / 2;      // items per scratch half
#ifndef JJS_SUB_CHUNK_LOG2
#define JJS_SUB_CHUNK_LOG2 19   // 2^18: 20.86 M/s device-resident on 2^20 singles, 2^19: 21.19 (fewer partially filled waves)
#endif
#ifndef JJS_FIRST_SLICE_LOG2
#define JJS_FIRST_SLICE_LOG2 17  // first slice of a host-buffer shard: short, so that compute starts while the rest is still being copied
#endif
constexpr size_t SUB_CHUNK = size_t(1) << JJS_SUB_CHUNK_LOG2;      // items per sub-chunk of an overlapped batch (<= SUB_ITEMS)
inline Region region_whole(const DeviceState& d) { return region_of(d, 0, CHUNK_ITEMS, 0); }
inline Region region_half(const DeviceState& d, size_t j) { return region_of(d, (j & 1) * SUB_ITEMS, SUB_ITEMS, (int)(j & 1)); }

}  // namespace

struct StageRecord {
    int stage;
    int device;
    cudaEvent_t e0, e1;
};

struct jjs_ctx {
    std::vector<DeviceState> dev;
    bool ready = false;              // jjs_init succeeded on every device; nothing else may touch the devices
    std::mutex mu;                   // guards err and records (the host entry points run one worker thread per device)
    char err[512];
    std::atomic<uint64_t> launches{0};
    bool profile = false;
    std::vector<StageRecord> records;
    std::vector<fq> tags;            // host copy of the SAFE tag table
};

namespace {

int fail(jjs_ctx* ctx, int code, const char* fmt, ...) {
    if (ctx) {
        std::lock_guard<std::mutex> lock(ctx->mu);
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

// Every entry point starts here: a context whose jjs_init failed exists only so that the caller can read the reason.
#define JJS_ENTER(ctx)                                                                                         \
    do {                                                                                                      \
        if (!(ctx)) return JJS_ERR_ARGUMENT;                                                                  \
        if (!(ctx)->ready || (ctx)->dev.empty()) return JJS_ERR_CUDA; /* err keeps the message of jjs_init */ \
        { std::lock_guard<std::mutex> lock_((ctx)->mu); (ctx)->err[0] = 0; }                                  \
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
        std::lock_guard<std::mutex> lock(ctx->mu);
        ctx->records.push_back(rec);
    }
};

inline unsigned blocks_for(size_t threads) { return (unsigned)((threads + BLOCK - 1) / BLOCK); }
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

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
    JJS_CUDA(ctx, cudaMalloc(&d.rcount, 8 * sizeof(uint32_t)));
    JJS_CUDA(ctx, cudaMalloc(&d.eqlist, sizeof(uint32_t) * 2 * CHUNK_ITEMS));
    JJS_CUDA(ctx, cudaMalloc(&d.tab, sizeof(fq) * AGG_GROUP * R32_TAB_FQ * TAB_THREADS));   // the key-aggregation kernel's four tables per thread
    {
        cudaDeviceProp prop;
        JJS_CUDA(ctx, cudaGetDeviceProperties(&prop, d.device));
        int per_sm = 0;
#if JJS_EQ_V2
        const size_t smem = sizeof(uint4) * 2 * EQ2_SLOTS * JJS_EQ_BLOCK;
        JJS_CUDA(ctx, cudaFuncSetAttribute(k_equation2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        JJS_CUDA(ctx, cudaFuncSetAttribute(k_equation2, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        JJS_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_equation2, JJS_EQ_BLOCK, smem));
        const size_t per_thread = 2 * EQ2_TAB_FQ, block = JJS_EQ_BLOCK;
#else
        JJS_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_equation<1>, BLOCK, 0));
        const size_t per_thread = 3 * 36, block = BLOCK;   // the var-gen equation builds three tables
#endif
        if (per_sm < 1) return fail(ctx, JJS_ERR_CUDA, "the equation kernel does not fit on an SM of device %d", d.device);
        d.eq_grid = per_sm * prop.multiProcessorCount;
        d.eqtab_half = per_thread * (size_t)d.eq_grid * block;
        JJS_CUDA(ctx, cudaMalloc(&d.eqtab, sizeof(fq) * 2 * d.eqtab_half));
    }
    return JJS_SUCCESS;
}

// grow-only staging buffer of the host-buffer entry points; growing waits for the device (nothing of this context is in
// flight at that point: the host entry points are synchronous)
int ensure_stage(jjs_ctx* ctx, DeviceState& d, size_t bytes) {
    if (bytes <= d.stage_bytes) return JJS_SUCCESS;
    JJS_CUDA(ctx, cudaSetDevice(d.device));
    JJS_CUDA(ctx, cudaDeviceSynchronize());
    cudaFree(d.stage);
    d.stage = nullptr;
    d.stage_bytes = 0;
    JJS_CUDA(ctx, cudaMalloc(&d.stage, bytes));
    d.stage_bytes = bytes;
    return JJS_SUCCESS;
}

// SAFE tags up to `n_absorb` absorbed elements on device d (multisig transcripts).  The table is sized at jjs_init for
// 1 024 elements (511 signers per aggregate key, 255 participants per session); a larger transcript grows it, which
// waits for the device once.
int ensure_tags(jjs_ctx* ctx, DeviceState& d, size_t n_absorb) {
    if (n_absorb < d.n_tags) return JJS_SUCCESS;
    size_t want = d.n_tags ? d.n_tags : 1024;
    while (want <= n_absorb) want *= 2;
    {
        std::lock_guard<std::mutex> lock(ctx->mu);
        while (ctx->tags.size() < want) {
            fq t;
            safe_tag_mont(t.l, (uint32_t)ctx->tags.size());
            ctx->tags.push_back(t);
        }
    }
    JJS_CUDA(ctx, cudaSetDevice(d.device));
    JJS_CUDA(ctx, cudaDeviceSynchronize());
    cudaFree(d.safe_tags);
    d.safe_tags = nullptr;
    d.n_tags = 0;
    JJS_CUDA(ctx, cudaMalloc(&d.safe_tags, sizeof(fq) * want));
    JJS_CUDA(ctx, cudaMemcpy(d.safe_tags, ctx->tags.data(), sizeof(fq) * want, cudaMemcpyHostToDevice));
    d.n_tags = want;
    return JJS_SUCCESS;
}

// key scratch of the aggregate path, both halves; growing waits for the device
constexpr size_t AGG_KEYS_INITIAL = size_t(1) << 21;
int ensure_agg_scratch(jjs_ctx* ctx, DeviceState& d, size_t keys) {
    if (keys <= d.agg[0].cap_keys) return JJS_SUCCESS;
    size_t want = d.agg[0].cap_keys ? d.agg[0].cap_keys : AGG_KEYS_INITIAL;
    while (want < keys) want *= 2;
    JJS_CUDA(ctx, cudaSetDevice(d.device));
    JJS_CUDA(ctx, cudaDeviceSynchronize());
    for (int h = 0; h < 2; h++) {
        AggScratch& a = d.agg[h];
        cudaFree(a.keys_u); cudaFree(a.keys_v); cudaFree(a.kflags); cudaFree(a.kcoef); cudaFree(a.kmap); cudaFree(a.kitem);
        a.keys_u = a.keys_v = nullptr; a.kflags = nullptr; a.kcoef = a.kmap = a.kitem = nullptr; a.cap_keys = 0;
        JJS_CUDA(ctx, cudaMalloc(&a.keys_u, sizeof(fq) * want));
        JJS_CUDA(ctx, cudaMalloc(&a.keys_v, sizeof(fq) * want));
        JJS_CUDA(ctx, cudaMalloc(&a.kflags, want));
        JJS_CUDA(ctx, cudaMalloc(&a.kcoef, 32 * want));
        JJS_CUDA(ctx, cudaMalloc(&a.kmap, sizeof(uint32_t) * want));
        JJS_CUDA(ctx, cudaMalloc(&a.kitem, sizeof(uint32_t) * want));
        if (!a.order) JJS_CUDA(ctx, cudaMalloc(&a.order, sizeof(uint32_t) * CHUNK_ITEMS));
        if (!a.sort) JJS_CUDA(ctx, cudaMalloc(&a.sort, sizeof(AggSort)));
        a.cap_keys = want;
    }
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
#if JJS_EQ_V2
    {
        const size_t smem = sizeof(uint4) * 2 * EQ2_SLOTS * JJS_EQ_BLOCK;
        const size_t want = (neq * m + JJS_EQ_BLOCK - 1) / JJS_EQ_BLOCK;
        const unsigned grid = (unsigned)(want < (size_t)d.eq_grid ? want : (size_t)d.eq_grid);
        k_equation2<<<grid, JJS_EQ_BLOCK, smem, stream>>>(variant, R.pts_u, R.pts_v, R.pflags, m, R.eqlist, R.rcount + 2, fu, R.cwords, R.eqflags,
                                                         R.eqtab, T, R.rlist, R.rcount);
        ctx->launches++;
    }
#elif JJS_EQ_PERSISTENT
    {
        const size_t want = blocks_for(neq * m);
        const unsigned grid = (unsigned)(want < (size_t)d.eq_grid ? want : (size_t)d.eq_grid);
        (variant == VAR_VARGEN ? k_equation<1> : k_equation<0>)<<<grid, BLOCK, 0, stream>>>(variant, R.pts_u, R.pts_v, R.pflags, m, 0, neq * m, R.eqlist, R.rcount + 2, fu,
                                                                                            R.cwords, R.eqflags, R.eqtab, (size_t)d.eq_grid * BLOCK, T, R.rlist, R.rcount);
        ctx->launches++;
    }
#else
    for (size_t first = 0; first < neq * m; first += R.cap) {
        size_t cnt = neq * m - first < R.cap ? neq * m - first : R.cap;
        (variant == VAR_VARGEN ? k_equation<1> : k_equation<0>)<<<blocks_for(cnt), BLOCK, 0, stream>>>(variant, R.pts_u, R.pts_v, R.pflags, m, first, cnt, R.eqlist,
                                                                                                       R.rcount + 2, fu, R.cwords, R.eqflags, R.tab, TAB_THREADS, T, R.rlist,
                                                                                                       R.rcount);
        ctx->launches++;
    }
#endif
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

// Typed inputs: the same pipeline behind k_decode_ext (points as JubJubExtended coordinates, 160 bytes each, item-major).
int run_chunk_ext(jjs_ctx* ctx, DeviceState& d, const Region& R, int variant, const uint8_t* pts, const uint8_t* u32, const uint8_t* msg, size_t m,
                  uint8_t* status, uint8_t* c_out, cudaStream_t stream) {
    const int slots = variant_slots(variant);
    WireField fmsg{msg, 32}, fu{u32, 32};
    StageTimer t0(ctx, d.device, 0, stream);
    k_decode_ext<<<blocks_for(slots * m), BLOCK, 0, stream>>>(pts, slots, m, R.pts_u, R.pts_v, R.pflags, key_slot_mask(variant));
    t0.stop(stream);
    ctx->launches++;
    int rc = enqueue_challenges(ctx, d, R, variant, m, fmsg, fu, stream, true);
    if (!rc) rc = enqueue_equations(ctx, d, R, variant, m, fu, stream);
    if (rc) return rc;
    k_finalize<<<blocks_for(m), BLOCK, 0, stream>>>(variant, R.pflags, R.iflags, R.eqflags, R.cwords, m, status, c_out);
    ctx->launches++;
    return JJS_SUCCESS;
}

// aggregate_pk(..).verify(..) for one chunk of m <= R.cap items with K <= cap_keys signer keys; everything on `stream`,
// nothing on the host.  d_keys: the chunk's first key; d_offsets: the chunk's m + 1 offsets (absolute, first = key_lo).
int run_chunk_aggregate(jjs_ctx* ctx, DeviceState& d, const Region& R, const uint8_t* d_keys, const uint32_t* d_offsets, uint32_t key_lo, size_t K,
                        const uint8_t* d_sig, const uint8_t* d_msg, size_t m, uint8_t* d_status, uint8_t* d_c, uint8_t* d_agg, cudaStream_t stream) {
    AggScratch& A = d.agg[R.half];
    Tables T = d.tables();
    Fields fk, fr;
    fk.f[0] = WireField{d_keys, 32};
    fr.f[0] = WireField{d_sig + 32, 64};
    fk.f[1] = fk.f[2] = fk.f[3] = fr.f[1] = fr.f[2] = fr.f[3] = WireField{nullptr, 0};
    WireField fmsg{d_msg, 32}, fu{d_sig, 64};
    StageTimer t0(ctx, d.device, 0, stream);
    if (K) k_decode<<<blocks_for(K), BLOCK, 0, stream>>>(fk, 1, 0, K, A.keys_u, A.keys_v, A.kflags, T, 0u);
    k_decode<<<blocks_for(m), BLOCK, 0, stream>>>(fr, 1, 1, m, R.pts_u, R.pts_v, R.pflags, T, 0u);
    t0.stop(stream);
    StageTimer t2(ctx, d.device, 2, stream);
    JJS_CUDA(ctx, cudaMemsetAsync(A.sort, 0, sizeof(AggSort), stream));
    k_agg_hist<<<blocks_for(m), BLOCK, 0, stream>>>(d_offsets, m, A.sort);
    k_agg_scan<<<1, 32, 0, stream>>>(A.sort);
    k_agg_scatter<<<blocks_for(m), BLOCK, 0, stream>>>(d_offsets, key_lo, m, A.sort, A.order, A.kmap, A.kitem);
    if (K) k_agg_coeffs<<<blocks_for(K), BLOCK, 0, stream>>>(A.keys_u, A.keys_v, A.kflags, d_offsets, A.kmap, A.kitem, key_lo, K, A.kcoef, T);
    k_aggregate<<<blocks_for(m), BLOCK, 0, stream>>>(A.keys_u, A.keys_v, A.kflags, d_offsets, A.order, key_lo, m, R.pts_u, R.pts_v, R.pflags, d_agg, R.tab,
                                                    TAB_THREADS, A.kcoef, T);
    t2.stop(stream);
    ctx->launches += 7;
    int rc = enqueue_challenges(ctx, d, R, VAR_SINGLE, m, fmsg, fu, stream, true);
    if (!rc) rc = enqueue_equations(ctx, d, R, VAR_SINGLE, m, fu, stream);
    if (rc) return rc;
    StageTimer t4(ctx, d.device, 4, stream);
    k_finalize<<<blocks_for(m), BLOCK, 0, stream>>>(VAR_SINGLE, R.pflags, R.iflags, R.eqflags, R.cwords, m, d_status, d_c);
    t4.stop(stream);
    ctx->launches++;
    return JJS_SUCCESS;
}

// Sub-chunks of an overlapped batch alternate between the two internal streams and scratch halves: kernels of neighbouring
// sub-chunks overlap, which hides the partially filled last wave of every launch (per-thread work is uniform, so a launch
// ends with SMs idling for up to one block duration).  With per-stage profiling on, everything stays on `stream` so that
// the stage timers do not overlap.
struct Overlap {
    jjs_ctx* ctx;
    DeviceState& d;
    cudaStream_t stream;
    bool on;
    size_t i = 0;
    Overlap(jjs_ctx* c, DeviceState& dev, cudaStream_t s, bool enable) : ctx(c), d(dev), stream(s), on(enable && !c->profile) {}
    int begin() {
        if (!on) return JJS_SUCCESS;
        JJS_CUDA(ctx, cudaEventRecord(d.fork, stream));
        for (int k = 0; k < 2; k++) JJS_CUDA(ctx, cudaStreamWaitEvent(d.sub[k], d.fork, 0));
        return JJS_SUCCESS;
    }
    cudaStream_t next_stream() const { return on ? d.sub[i & 1] : stream; }
    Region next_region() const { return on ? region_half(d, i) : region_whole(d); }
    void advance() { i++; }
    int end() {
        if (!on) return JJS_SUCCESS;
        for (int k = 0; k < 2; k++) {
            JJS_CUDA(ctx, cudaEventRecord(d.join[k], d.sub[k]));
            JJS_CUDA(ctx, cudaStreamWaitEvent(stream, d.join[k], 0));
        }
        return JJS_SUCCESS;
    }
};

// Enqueue the whole pipeline for n items (device pointers); the work is ordered after what `stream` holds now and
// `stream` waits for it.
int run_device(jjs_ctx* ctx, DeviceState& d, int variant, const uint8_t* pk, const uint8_t* sig, const uint8_t* msg, size_t n,
               uint8_t* status, uint8_t* c_out, cudaStream_t stream, bool challenge_only = false) {
    int rc = ensure_scratch(ctx, d);
    if (rc) return rc;
    JJS_CUDA(ctx, cudaSetDevice(d.device));
    Overlap ov(ctx, d, stream, n > SUB_CHUNK);
    rc = ov.begin();
    if (rc) return rc;
    const size_t step = ov.on ? SUB_CHUNK : CHUNK_ITEMS;
    for (size_t off = 0; off < n; off += step, ov.advance()) {
        size_t m = n - off < step ? n - off : step;
        rc = run_chunk(ctx, d, ov.next_region(), variant, pk + off * pk_size(variant), sig + off * sig_size(variant), msg + off * 32, m,
                       status ? status + off : nullptr, c_out ? c_out + off * 32 : nullptr, ov.next_stream(), challenge_only);
        if (rc) return rc;
    }
    rc = ov.end();
    if (rc) return rc;
    JJS_CUDA(ctx, cudaGetLastError());
    return JJS_SUCCESS;
}

// Chunks of an aggregate-key batch: at most `max_items` items and `max_keys` signer keys each (an item with more keys than
// that forms a chunk of its own, for which the caller grows the key scratch first).
inline size_t agg_chunk_end(const uint32_t* offsets, size_t first, size_t n, size_t max_items, size_t max_keys) {
    size_t end = first;
    while (end < n && end - first < max_items && (size_t)(offsets[end + 1] - offsets[first]) <= max_keys) end++;
    return end == first ? first + 1 : end;
}
inline size_t agg_max_item_keys(const uint32_t* offsets, size_t n) {
    size_t mx = 0;
    for (size_t i = 0; i < n; i++) {
        size_t c = offsets[i + 1] - offsets[i];
        if (c > mx) mx = c;
    }
    return mx;
}

// aggregate_pk(..).verify(..) for n items whose wire data already sits on the device.  `h_offsets` is the host copy of
// the n + 1 offsets (it plans the chunks and sizes the tag table; it is read during the call only); d_offsets the same
// array on the device.  Enqueue-only unless a scratch has to grow (first call, or more keys / signers than ever before).
int run_aggregate_device(jjs_ctx* ctx, DeviceState& d, const uint8_t* d_pks, const uint32_t* d_offsets, const uint32_t* h_offsets,
                         const uint8_t* d_sig, const uint8_t* d_msg, size_t n, uint8_t* d_status, uint8_t* d_c, uint8_t* d_agg,
                         cudaStream_t stream) {
    int rc = ensure_scratch(ctx, d);
    const size_t max_cnt = agg_max_item_keys(h_offsets, n);
    if (!rc) rc = ensure_tags(ctx, d, 2 + 2 * max_cnt);
    if (!rc) rc = ensure_agg_scratch(ctx, d, max_cnt > AGG_KEYS_INITIAL ? max_cnt : AGG_KEYS_INITIAL);
    if (rc) return rc;
    JJS_CUDA(ctx, cudaSetDevice(d.device));
    Overlap ov(ctx, d, stream, n > SUB_CHUNK);
    rc = ov.begin();
    if (rc) return rc;
    const size_t step = ov.on ? SUB_CHUNK : SUB_ITEMS;
    for (size_t off = 0; off < n; ov.advance()) {
        size_t end = agg_chunk_end(h_offsets, off, n, step, d.agg[0].cap_keys);
        size_t m = end - off;
        uint32_t key_lo = h_offsets[off];
        size_t K = h_offsets[end] - key_lo;
        Region R = ov.on ? ov.next_region() : region_half(d, 0);
        rc = run_chunk_aggregate(ctx, d, R, d_pks + 32 * (size_t)key_lo, d_offsets + off, key_lo, K, d_sig + 64 * off, d_msg + 32 * off, m, d_status + off,
                                 d_c ? d_c + 32 * off : nullptr, d_agg ? d_agg + 32 * off : nullptr, ov.next_stream());
        if (rc) return rc;
        off = end;
    }
    rc = ov.end();
    if (rc) return rc;
    JJS_CUDA(ctx, cudaGetLastError());
    return JJS_SUCCESS;
}

// ---- host-buffer entry points ---------------------------------------------------------------------------------------------
// A call verifies one or more PARTS (each a homogeneous batch: one of the three wire variants, aggregate-key items, or typed
// items).  Every part is cut into contiguous shards, one per device of the context (small parts go whole to the least
// loaded device), and every device is driven by its own host thread: copies in on the copy stream, pipeline slices alternating
// between the two compute streams, results copied out at the end, one synchronisation.  A single host thread would serialise
// the devices whenever the caller's memory is pageable (cudaMemcpyAsync from pageable memory returns only after staging).
enum PartMode : int { PART_WIRE = 0, PART_AGGREGATE = 1, PART_TYPED = 2 };
struct HostPart {
    int mode, variant;
    const uint8_t *pk, *sig, *msg;    // typed: pk = points_ext160, sig = u32
    const uint32_t* offsets;          // aggregate only
    size_t n;
    uint8_t *status, *c_out, *agg_out;
    uint32_t* bitmap;
    bool challenge_only;
};
struct Shard {
    size_t lo, hi;
};

inline double part_weight(const HostPart& p) {   // relative cost of one item (executed multiplies, single = 1)
    if (p.mode == PART_AGGREGATE) {
        double avg = p.n ? (double)(p.offsets[p.n] - p.offsets[0]) / (double)p.n : 0.0;
        return 1.0 + 0.75 * avg;
    }
    return p.variant == VAR_DOUBLE ? 1.9 : (p.variant == VAR_VARGEN ? 1.4 : 1.0);
}

// shards[part][device].  Large parts are split evenly (every device then carries the same share of every kind, so the
// devices finish together whatever the mix); parts too small to be worth splitting go whole to the least loaded device.
constexpr size_t MIN_SPLIT_ITEMS = 8192;
void plan_shards(const std::vector<HostPart>& parts, size_t g, std::vector<std::vector<Shard>>& shards) {
    shards.assign(parts.size(), std::vector<Shard>(g, Shard{0, 0}));
    std::vector<double> load(g, 0.0);
    for (size_t p = 0; p < parts.size(); p++) {
        const size_t n = parts[p].n;
        if (n >= g * MIN_SPLIT_ITEMS) {
            const size_t per = ((n + g - 1) / g + 31) & ~size_t(31);   // shards start on a bitmap word
            for (size_t k = 0; k < g; k++) {
                size_t lo = k * per < n ? k * per : n, hi = lo + per < n ? lo + per : n;
                shards[p][k] = Shard{lo, hi};
                load[k] += part_weight(parts[p]) * (double)(hi - lo);
            }
        }
    }
    for (size_t p = 0; p < parts.size(); p++) {
        const size_t n = parts[p].n;
        if (n == 0 || n >= g * MIN_SPLIT_ITEMS) continue;
        size_t best = 0;
        for (size_t k = 1; k < g; k++)
            if (load[k] < load[best]) best = k;
        shards[p][best] = Shard{0, n};
        load[best] += part_weight(parts[p]) * (double)n;
    }
}

struct StagedPart {   // device staging of one part's shard
    uint8_t *pk = nullptr, *sig = nullptr, *msg = nullptr, *status = nullptr, *c = nullptr, *agg = nullptr;
    uint32_t *offsets = nullptr, *bitmap = nullptr;
};

inline size_t part_in_bytes(const HostPart& p, size_t lo, size_t hi, size_t* pk_bytes, size_t* sig_bytes) {
    const size_t m = hi - lo;
    if (p.mode == PART_AGGREGATE) {
        *pk_bytes = 32 * (size_t)(p.offsets[hi] - p.offsets[lo]);
        *sig_bytes = 64 * m;
    } else if (p.mode == PART_TYPED) {
        *pk_bytes = 160 * (size_t)variant_slots(p.variant) * m;
        *sig_bytes = 32 * m;
    } else {
        *pk_bytes = pk_size(p.variant) * m;
        *sig_bytes = sig_size(p.variant) * m;
    }
    return *pk_bytes + *sig_bytes + 32 * m;
}

// Everything device k does for one call.  Runs on its own host thread when the context has several devices.
int device_job(jjs_ctx* ctx, size_t k, const std::vector<HostPart>& parts, const std::vector<std::vector<Shard>>& shards) {
    DeviceState& d = ctx->dev[k];
    bool any = false;
    for (size_t p = 0; p < parts.size(); p++) any = any || shards[p][k].hi > shards[p][k].lo;
    if (!any) return JJS_SUCCESS;
    JJS_CUDA(ctx, cudaSetDevice(d.device));
    int rc = ensure_scratch(ctx, d);
    if (rc) return rc;
    // staging layout
    std::vector<StagedPart> st(parts.size());
    std::vector<size_t> o_pk(parts.size()), o_sig(parts.size()), o_msg(parts.size()), o_st(parts.size()), o_c(parts.size()), o_agg(parts.size()),
        o_off(parts.size()), o_bm(parts.size());
    size_t total = 0;
    size_t max_cnt = 0;
    for (size_t p = 0; p < parts.size(); p++) {
        const HostPart& P = parts[p];
        const size_t lo = shards[p][k].lo, hi = shards[p][k].hi, m = hi - lo;
        if (!m) continue;
        size_t pkb, sgb;
        part_in_bytes(P, lo, hi, &pkb, &sgb);
        o_pk[p] = total; total += align_up(pkb, 256);
        o_sig[p] = total; total += align_up(sgb, 256);
        o_msg[p] = total; total += align_up(32 * m, 256);
        o_st[p] = total; total += align_up(m, 256);
        o_c[p] = total; total += align_up(32 * m, 256);
        o_bm[p] = total; total += align_up(4 * ((m + 31) / 32), 256);
        if (P.mode == PART_AGGREGATE) {
            o_agg[p] = total; total += align_up(32 * m, 256);
            o_off[p] = total; total += align_up(4 * (m + 1), 256);
            size_t mc = agg_max_item_keys(P.offsets + lo, m);
            if (mc > max_cnt) max_cnt = mc;
            rc = ensure_agg_scratch(ctx, d, mc > AGG_KEYS_INITIAL ? mc : AGG_KEYS_INITIAL);
            if (rc) return rc;
        }
    }
    if (max_cnt) {
        rc = ensure_tags(ctx, d, 2 + 2 * max_cnt);
        if (rc) return rc;
    }
    rc = ensure_stage(ctx, d, total);
    if (rc) return rc;
    const bool serial = ctx->profile;   // stage timers must not overlap
    size_t j = 0;                       // slice counter of this device: slices alternate between the compute streams / scratch halves
    for (size_t p = 0; p < parts.size(); p++) {
        const HostPart& P = parts[p];
        const size_t lo = shards[p][k].lo, hi = shards[p][k].hi, m = hi - lo;
        if (!m) continue;
        StagedPart& S = st[p];
        S.pk = d.stage + o_pk[p]; S.sig = d.stage + o_sig[p]; S.msg = d.stage + o_msg[p]; S.status = d.stage + o_st[p]; S.c = d.stage + o_c[p];
        S.bitmap = reinterpret_cast<uint32_t*>(d.stage + o_bm[p]);
        if (P.mode == PART_AGGREGATE) {
            S.agg = d.stage + o_agg[p];
            S.offsets = reinterpret_cast<uint32_t*>(d.stage + o_off[p]);
            JJS_CUDA(ctx, cudaMemcpyAsync(S.offsets, P.offsets + lo, 4 * (m + 1), cudaMemcpyHostToDevice, d.copy_stream));
        }
        const bool want_c = P.c_out != nullptr || P.challenge_only;
        const size_t pks = P.mode == PART_TYPED ? 160 * (size_t)variant_slots(P.variant) : pk_size(P.variant);
        const size_t sgs = P.mode == PART_TYPED ? 32 : (P.mode == PART_AGGREGATE ? 64 : sig_size(P.variant));
        // Pipeline slices: slice j + 1 is copied in on the copy stream while slice j is being verified (the kernels of one
        // slice run far longer than its copy); the first slice of a device is short so that compute starts early.
        for (size_t off = 0; off < m; j++) {
            const size_t want = j == 0 ? (size_t(1) << JJS_FIRST_SLICE_LOG2) : SUB_CHUNK;
            size_t cnt = m - off < want ? m - off : want;
            cudaStream_t cs = serial ? d.stream : d.sub[j & 1];
            Region R = region_half(d, j);
            if (P.mode == PART_AGGREGATE) {
                const uint32_t* ho = P.offsets + lo;   // host offsets of this shard
                size_t end = agg_chunk_end(ho, off, m, want, d.agg[0].cap_keys);
                cnt = end - off;
                const uint32_t key0 = ho[0], key_lo = ho[off];
                const size_t K = ho[end] - key_lo;
                if (K) JJS_CUDA(ctx, cudaMemcpyAsync(S.pk + 32 * (size_t)(key_lo - key0), P.pk + 32 * (size_t)key_lo, 32 * K, cudaMemcpyHostToDevice, d.copy_stream));
                JJS_CUDA(ctx, cudaMemcpyAsync(S.sig + 64 * off, P.sig + 64 * (lo + off), 64 * cnt, cudaMemcpyHostToDevice, d.copy_stream));
                JJS_CUDA(ctx, cudaMemcpyAsync(S.msg + 32 * off, P.msg + 32 * (lo + off), 32 * cnt, cudaMemcpyHostToDevice, d.copy_stream));
                JJS_CUDA(ctx, cudaEventRecord(d.copied, d.copy_stream));
                JJS_CUDA(ctx, cudaStreamWaitEvent(cs, d.copied, 0));
                rc = run_chunk_aggregate(ctx, d, R, S.pk + 32 * (size_t)(key_lo - key0), S.offsets + off, key_lo, K, S.sig + 64 * off, S.msg + 32 * off, cnt,
                                         S.status + off, want_c ? S.c + 32 * off : nullptr, P.agg_out ? S.agg + 32 * off : nullptr, cs);
            } else {
                JJS_CUDA(ctx, cudaMemcpyAsync(S.pk + off * pks, P.pk + (lo + off) * pks, cnt * pks, cudaMemcpyHostToDevice, d.copy_stream));
                JJS_CUDA(ctx, cudaMemcpyAsync(S.sig + off * sgs, P.sig + (lo + off) * sgs, cnt * sgs, cudaMemcpyHostToDevice, d.copy_stream));
                JJS_CUDA(ctx, cudaMemcpyAsync(S.msg + off * 32, P.msg + (lo + off) * 32, cnt * 32, cudaMemcpyHostToDevice, d.copy_stream));
                JJS_CUDA(ctx, cudaEventRecord(d.copied, d.copy_stream));
                JJS_CUDA(ctx, cudaStreamWaitEvent(cs, d.copied, 0));
                if (P.mode == PART_TYPED)
                    rc = run_chunk_ext(ctx, d, R, P.variant, S.pk + off * pks, S.sig + off * 32, S.msg + off * 32, cnt, S.status + off,
                                       want_c ? S.c + off * 32 : nullptr, cs);
                else
                    rc = run_chunk(ctx, d, R, P.variant, S.pk + off * pks, S.sig + off * sgs, S.msg + off * 32, cnt, S.status + off,
                                   want_c ? S.c + off * 32 : nullptr, cs, P.challenge_only);
            }
            if (rc) return rc;
            off += cnt;
        }
    }
    if (!serial)
        for (int q = 0; q < 2; q++) {
            JJS_CUDA(ctx, cudaEventRecord(d.join[q], d.sub[q]));
            JJS_CUDA(ctx, cudaStreamWaitEvent(d.stream, d.join[q], 0));
        }
    for (size_t p = 0; p < parts.size(); p++) {
        const HostPart& P = parts[p];
        const size_t lo = shards[p][k].lo, hi = shards[p][k].hi, m = hi - lo;
        if (!m) continue;
        const StagedPart& S = st[p];
        if (!P.challenge_only && P.status) JJS_CUDA(ctx, cudaMemcpyAsync(P.status + lo, S.status, m, cudaMemcpyDeviceToHost, d.stream));
        if (P.c_out) JJS_CUDA(ctx, cudaMemcpyAsync(P.c_out + lo * 32, S.c, m * 32, cudaMemcpyDeviceToHost, d.stream));
        if (P.agg_out) JJS_CUDA(ctx, cudaMemcpyAsync(P.agg_out + lo * 32, S.agg, m * 32, cudaMemcpyDeviceToHost, d.stream));
        if (P.bitmap) {
            k_bitmap<<<blocks_for(m), BLOCK, 0, d.stream>>>(S.status, m, S.bitmap);
            ctx->launches++;
            JJS_CUDA(ctx, cudaMemcpyAsync(P.bitmap + lo / 32, S.bitmap, 4 * ((m + 31) / 32), cudaMemcpyDeviceToHost, d.stream));
        }
    }
    JJS_CUDA(ctx, cudaGetLastError());
    JJS_CUDA(ctx, cudaStreamSynchronize(d.stream));
    return JJS_SUCCESS;
}

// device_job, and on failure a drain of the device: nothing may still be reading or writing the caller's buffers when the
// call returns
int device_job_drained(jjs_ctx* ctx, size_t k, const std::vector<HostPart>& parts, const std::vector<std::vector<Shard>>& shards) {
    int rc = device_job(ctx, k, parts, shards);
    if (rc) {
        cudaSetDevice(ctx->dev[k].device);
        cudaDeviceSynchronize();
    }
    return rc;
}

int validate_part(jjs_ctx* ctx, const HostPart& P) {
    if (P.mode == PART_WIRE || P.mode == PART_TYPED) {
        if (P.variant < 0 || P.variant > 2) return fail(ctx, JJS_ERR_ARGUMENT, "bad variant");
    } else if (P.mode != PART_AGGREGATE) {
        return fail(ctx, JJS_ERR_ARGUMENT, "bad part kind");
    }
    if (P.n == 0) return JJS_SUCCESS;
    if (P.challenge_only) {
        if (!P.pk || !P.sig || !P.msg || !P.c_out) return fail(ctx, JJS_ERR_ARGUMENT, "null buffer");
        return JJS_SUCCESS;
    }
    if (!P.sig || !P.msg || (!P.status && !P.bitmap)) return fail(ctx, JJS_ERR_ARGUMENT, "null buffer");
    if (P.mode == PART_AGGREGATE) {
        if (!P.offsets) return fail(ctx, JJS_ERR_ARGUMENT, "null buffer");
        for (size_t i = 0; i < P.n; i++)
            if (P.offsets[i + 1] < P.offsets[i]) return fail(ctx, JJS_ERR_ARGUMENT, "offsets must be non-decreasing");
        if (P.offsets[P.n] > P.offsets[0] && !P.pk) return fail(ctx, JJS_ERR_ARGUMENT, "null key buffer");
    } else if (!P.pk) {
        return fail(ctx, JJS_ERR_ARGUMENT, "null buffer");
    }
    return JJS_SUCCESS;
}

int run_parts(jjs_ctx* ctx, const std::vector<HostPart>& parts) {
    JJS_ENTER(ctx);
    bool any = false;
    for (const HostPart& P : parts) {
        int rc = validate_part(ctx, P);
        if (rc) return rc;
        any = any || P.n > 0;
    }
    if (!any) return JJS_SUCCESS;
    const size_t g = ctx->dev.size();
    std::vector<std::vector<Shard>> shards;
    plan_shards(parts, g, shards);
    std::vector<int> rcs(g, JJS_SUCCESS);
    if (g == 1) {
        rcs[0] = device_job_drained(ctx, 0, parts, shards);
    } else {
        std::vector<std::thread> workers;
        workers.reserve(g);
        for (size_t k = 0; k < g; k++) workers.emplace_back([&, k] { rcs[k] = device_job_drained(ctx, k, parts, shards); });
        for (auto& w : workers) w.join();
    }
    for (size_t k = 0; k < g; k++)
        if (rcs[k]) return rcs[k];
    return JJS_SUCCESS;
}

HostPart wire_part(int variant, const uint8_t* pk, const uint8_t* sig, const uint8_t* msg, size_t n, uint8_t* status, uint8_t* c_out, uint32_t* bitmap,
                   bool challenge_only = false) {
    return HostPart{PART_WIRE, variant, pk, sig, msg, nullptr, n, status, c_out, nullptr, bitmap, challenge_only};
}
HostPart aggregate_part(const uint8_t* pks, const uint32_t* offsets, const uint8_t* sig, const uint8_t* msg, size_t n, uint8_t* status, uint8_t* c_out,
                        uint8_t* agg_out, uint32_t* bitmap) {
    return HostPart{PART_AGGREGATE, VAR_SINGLE, pks, sig, msg, offsets, n, status, c_out, agg_out, bitmap, false};
}

// multisig::combine for n ragged sessions (host buffers, device 0); chunks hold at most 2^20 participants
int run_msig(jjs_ctx* ctx, const uint8_t* pks, const uint8_t* Rs, const uint8_t* Ss, const uint8_t* zs, const uint32_t* offsets, const uint8_t* msg,
             size_t n, uint8_t* share_ok, uint8_t* status, uint32_t* bad, uint8_t* sig) {
    JJS_ENTER(ctx);
    if (n == 0) return JJS_SUCCESS;
    if (!offsets || !msg || !status) return fail(ctx, JJS_ERR_ARGUMENT, "null buffer");
    if (offsets[0] != 0) return fail(ctx, JJS_ERR_ARGUMENT, "offsets[0] must be 0");
    for (size_t i = 0; i < n; i++)
        if (offsets[i + 1] < offsets[i]) return fail(ctx, JJS_ERR_ARGUMENT, "offsets must be non-decreasing");
    if (offsets[n] && (!pks || !Rs || !Ss || !zs)) return fail(ctx, JJS_ERR_ARGUMENT, "null buffer");
    DeviceState& d = ctx->dev[0];
    int rc = ensure_scratch(ctx, d);
    if (!rc) rc = ensure_tags(ctx, d, 3 + 4 * agg_max_item_keys(offsets, n));
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
        rc = ensure_stage(ctx, d, o);
        if (rc) return rc;
        uint8_t* B = d.stage;
        cudaStream_t st = d.stream;
        cudaError_t ce = cudaSuccess;
        auto up = [&](size_t at, const void* src, size_t bytes) { if (ce == cudaSuccess && bytes) ce = cudaMemcpyAsync(B + at, src, bytes, cudaMemcpyHostToDevice, st); };
        up(o_pk, pks + 32 * k0, 32 * K); up(o_R, Rs + 32 * k0, 32 * K); up(o_S, Ss + 32 * k0, 32 * K); up(o_z, zs + 32 * k0, 32 * K);
        up(o_own, owner.data(), 4 * K); up(o_msg, msg + 32 * s0, 32 * m); up(o_off, rel.data(), 4 * (m + 1)); up(o_ord, order.data(), 4 * m);
        MsigBuffers b;
        b.pu = reinterpret_cast<fq*>(B + o_pu); b.pv = reinterpret_cast<fq*>(B + o_pv); b.pf = B + o_pf;
        b.d_words = reinterpret_cast<uint32_t*>(B + o_d); b.cd_words = reinterpret_cast<uint32_t*>(B + o_cd); b.a_words = reinterpret_cast<uint32_t*>(B + o_a);
        b.rsa_u = reinterpret_cast<fq*>(B + o_ru); b.rsa_v = reinterpret_cast<fq*>(B + o_rv); b.sflags = B + o_sf; b.share_ok = B + o_ok;
        b.offsets = reinterpret_cast<uint32_t*>(B + o_off); b.owner = reinterpret_cast<uint32_t*>(B + o_own); b.order = reinterpret_cast<uint32_t*>(B + o_ord);
        Fields f;
        f.f[0] = WireField{B + o_pk, 32}; f.f[1] = WireField{B + o_R, 32}; f.f[2] = WireField{B + o_S, 32}; f.f[3] = WireField{nullptr, 0};
        WireField fmsg{B + o_msg, 32}, fz{B + o_z, 32};
        if (K) k_decode<<<blocks_for(3 * K), BLOCK, 0, st>>>(f, 3, 0, K, b.pu, b.pv, b.pf, T, 0u);
        k_msig_session<<<blocks_for(m), BLOCK, 0, st>>>(b, K, m, fmsg, fz, d.tab, TAB_THREADS, T);
        if (K) k_msig_share<<<blocks_for(K), BLOCK, 0, st>>>(b, K, fz, d.tab, TAB_THREADS, T);
        k_msig_finalize<<<blocks_for(m), BLOCK, 0, st>>>(b, m, fz, B + o_st, reinterpret_cast<uint32_t*>(B + o_bad), B + o_sig);
        ctx->launches += 4;
        auto down = [&](void* dst, size_t at, size_t bytes) { if (ce == cudaSuccess && bytes) ce = cudaMemcpyAsync(dst, B + at, bytes, cudaMemcpyDeviceToHost, st); };
        down(status + s0, o_st, m);
        if (bad) down(bad + s0, o_bad, 4 * m);
        if (sig) down(sig + 64 * s0, o_sig, 64 * m);
        if (share_ok) down(share_ok + k0, o_ok, K);
        cudaError_t e = cudaStreamSynchronize(st);   // also the drain of the error path: the host vectors above are reused by the next chunk
        if (ce != cudaSuccess) e = ce;
        if (e != cudaSuccess) return fail(ctx, JJS_ERR_CUDA, "multisig combine failed: %s", cudaGetErrorString(e));
        s0 = s1;
    }
    return JJS_SUCCESS;
}

int run_sign(jjs_ctx* ctx, int variant, const uint8_t* sk, const uint8_t* rnd, const uint8_t* gsc, const uint8_t* msg, size_t n, uint8_t* pk_out,
             uint8_t* sig_out) {
    JJS_ENTER(ctx);
    if (variant < 0 || variant > 2) return fail(ctx, JJS_ERR_ARGUMENT, "bad variant");
    if (n == 0) return JJS_SUCCESS;
    if (!sk || !rnd || !msg || !pk_out || !sig_out || (variant == VAR_VARGEN && !gsc)) return fail(ctx, JJS_ERR_ARGUMENT, "null buffer");
    DeviceState& d = ctx->dev[0];
    JJS_CUDA(ctx, cudaSetDevice(d.device));
    const size_t pks = pk_size(variant), sgs = sig_size(variant);
    const size_t chunk = n < CHUNK_ITEMS ? n : CHUNK_ITEMS;
    int rc = ensure_stage(ctx, d, chunk * (4 * 32 + pks + sgs));
    if (rc) return rc;
    uint8_t* buf = d.stage;
    uint8_t *d_sk = buf, *d_rnd = buf + 32 * chunk, *d_g = buf + 64 * chunk, *d_msg = buf + 96 * chunk, *d_pk = buf + 128 * chunk,
            *d_sig = d_pk + pks * chunk;
    for (size_t off = 0; off < n; off += chunk) {
        size_t m = n - off < chunk ? n - off : chunk;
        cudaError_t ce = cudaMemcpyAsync(d_sk, sk + 32 * off, 32 * m, cudaMemcpyHostToDevice, d.stream);
        if (ce == cudaSuccess) ce = cudaMemcpyAsync(d_rnd, rnd + 32 * off, 32 * m, cudaMemcpyHostToDevice, d.stream);
        if (ce == cudaSuccess && variant == VAR_VARGEN) ce = cudaMemcpyAsync(d_g, gsc + 32 * off, 32 * m, cudaMemcpyHostToDevice, d.stream);
        if (ce == cudaSuccess) ce = cudaMemcpyAsync(d_msg, msg + 32 * off, 32 * m, cudaMemcpyHostToDevice, d.stream);
        if (ce == cudaSuccess) {
            if (variant == VAR_SINGLE) k_sign<VAR_SINGLE><<<blocks_for(m), BLOCK, 0, d.stream>>>(d_sk, d_rnd, d_g, d_msg, m, d_pk, d_sig, d.tables());
            else if (variant == VAR_DOUBLE) k_sign<VAR_DOUBLE><<<blocks_for(m), BLOCK, 0, d.stream>>>(d_sk, d_rnd, d_g, d_msg, m, d_pk, d_sig, d.tables());
            else k_sign<VAR_VARGEN><<<blocks_for(m), BLOCK, 0, d.stream>>>(d_sk, d_rnd, d_g, d_msg, m, d_pk, d_sig, d.tables());
            ctx->launches++;
            ce = cudaMemcpyAsync(pk_out + pks * off, d_pk, pks * m, cudaMemcpyDeviceToHost, d.stream);
        }
        if (ce == cudaSuccess) ce = cudaMemcpyAsync(sig_out + sgs * off, d_sig, sgs * m, cudaMemcpyDeviceToHost, d.stream);
        cudaError_t e = cudaStreamSynchronize(d.stream);
        if (ce != cudaSuccess) e = ce;
        if (e != cudaSuccess) return fail(ctx, JJS_ERR_CUDA, "sign batch failed: %s", cudaGetErrorString(e));
    }
    return JJS_SUCCESS;
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
    {
        static_assert((1 << FB_LO) % FB_BATCH == 0, "a batch of table entries must share its high half");
        fq *small_u = nullptr, *small_v = nullptr;
        JJS_CUDA(ctx, cudaMalloc(&small_u, sizeof(fq) * FB_WINDOWS * FB_SMALL));
        {
            cudaError_t em = cudaMalloc(&small_v, sizeof(fq) * FB_WINDOWS * FB_SMALL);
            if (em != cudaSuccess) cudaFree(small_u);
            JJS_CUDA(ctx, em);
        }
        for (int which = 0; which < 2; which++) {
            k_fb_small<<<blocks_for(FB_WINDOWS * FB_SMALL), BLOCK, 0, d.stream>>>(small_u, small_v, which);
            k_fb_combine<<<blocks_for((size_t)FB_WINDOWS * FB_ENTRIES / FB_BATCH), BLOCK, 0, d.stream>>>(which ? d.fb_gn : d.fb_g, small_u, small_v);
        }
        ctx->launches += 4;
        cudaError_t e1 = cudaGetLastError(), e2 = cudaStreamSynchronize(d.stream);
        cudaFree(small_u);
        cudaFree(small_v);
        JJS_CUDA(ctx, e1);
        JJS_CUDA(ctx, e2);
    }
    return ensure_tags(ctx, d, 1023);
}

void free_device(DeviceState& d) {
    if (d.device < 0) return;
    cudaSetDevice(d.device);
    cudaDeviceSynchronize();
    cudaFree(d.root_tables); cudaFree(d.dlog_hash); cudaFree(d.fb_g); cudaFree(d.fb_gn); cudaFree(d.safe_tags);
    cudaFree(d.pts_u); cudaFree(d.pts_v); cudaFree(d.tab); cudaFree(d.eqtab); cudaFree(d.pflags); cudaFree(d.iflags); cudaFree(d.eqflags); cudaFree(d.cwords);
    cudaFree(d.rlist); cudaFree(d.rcount); cudaFree(d.eqlist); cudaFree(d.stage);
    for (int h = 0; h < 2; h++) {
        AggScratch& a = d.agg[h];
        cudaFree(a.keys_u); cudaFree(a.keys_v); cudaFree(a.kflags); cudaFree(a.kcoef); cudaFree(a.kmap); cudaFree(a.kitem); cudaFree(a.order); cudaFree(a.sort);
    }
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
    JJS_ENTER(ctx);
    if (device_index < 0 || (size_t)device_index >= ctx->dev.size()) return fail(ctx, JJS_ERR_ARGUMENT, "device_index out of range");
    if (n == 0) return JJS_SUCCESS;
    if (!pk || !sig || !msg || !status) return fail(ctx, JJS_ERR_ARGUMENT, "null buffer");
    DeviceState& d = ctx->dev[device_index];
    return run_device(ctx, d, variant, pk, sig, msg, n, status, c_out, (cudaStream_t)stream);
}

// small synchronous utilities on device 0: copy in, one kernel, copy out
template <typename Launch>
int run_small(jjs_ctx* ctx, const char* what, size_t in_bytes, size_t out_bytes, const void* const* srcs, const size_t* src_bytes, int n_src, void* dst,
              Launch launch) {
    DeviceState& d = ctx->dev[0];
    JJS_CUDA(ctx, cudaSetDevice(d.device));
    int rc = ensure_stage(ctx, d, align_up(in_bytes, 256) + out_bytes);
    if (rc) return rc;
    cudaError_t ce = cudaSuccess;
    size_t at = 0;
    for (int i = 0; i < n_src && ce == cudaSuccess; i++) {
        if (src_bytes[i]) ce = cudaMemcpyAsync(d.stage + at, srcs[i], src_bytes[i], cudaMemcpyHostToDevice, d.stream);
        at += src_bytes[i];
    }
    uint8_t* d_out = d.stage + align_up(in_bytes, 256);
    if (ce == cudaSuccess) {
        launch(d, d.stage, d_out);
        ce = cudaGetLastError();
    }
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(dst, d_out, out_bytes, cudaMemcpyDeviceToHost, d.stream);
    cudaError_t e = cudaStreamSynchronize(d.stream);
    if (ce != cudaSuccess) e = ce;
    if (e != cudaSuccess) return fail(ctx, JJS_ERR_CUDA, "%s failed: %s", what, cudaGetErrorString(e));
    return JJS_SUCCESS;
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
    *out = ctx;   // on failure the context is returned unusable, only so that the caller can read the reason: no CPU fallback
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count < 1)
        return fail(ctx, JJS_ERR_CUDA, "no CUDA device available: %s", e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    ctx->dev.resize(n_devices);
    for (int i = 0; i < n_devices; i++) {
        int ord = devices ? devices[i] : i;
        if (ord < 0 || ord >= count) return fail(ctx, JJS_ERR_ARGUMENT, "device ordinal %d not present (%d devices)", ord, count);
        for (int k = 0; k < i; k++)
            if (ctx->dev[k].device == ord) return fail(ctx, JJS_ERR_ARGUMENT, "device ordinal %d listed twice", ord);
        ctx->dev[i].device = ord;
        int rc = init_device(ctx, ctx->dev[i]);
        if (rc) return rc;
    }
    ctx->ready = true;
    return JJS_SUCCESS;
}

JJS_API void jjs_destroy(jjs_ctx* ctx) {
    if (!ctx) return;
    for (auto& d : ctx->dev) free_device(d);
    delete ctx;
}

JJS_API const char* jjs_last_error(const jjs_ctx* ctx) { return ctx ? ctx->err : "null context"; }
JJS_API int jjs_device_count(const jjs_ctx* ctx) { return ctx && ctx->ready ? (int)ctx->dev.size() : 0; }
JJS_API uint64_t jjs_launch_count(const jjs_ctx* ctx) { return ctx ? ctx->launches.load() : 0; }

JJS_API int jjs_verify_single(jjs_ctx* ctx, const uint8_t* pk32, const uint8_t* sig64, const uint8_t* msg32, size_t n, uint8_t* status,
                              uint8_t* c32_or_null) {
    return run_parts(ctx, {wire_part(VAR_SINGLE, pk32, sig64, msg32, n, status, c32_or_null, nullptr)});
}
JJS_API int jjs_verify_double(jjs_ctx* ctx, const uint8_t* pk64, const uint8_t* sig96, const uint8_t* msg32, size_t n, uint8_t* status,
                              uint8_t* c32_or_null) {
    return run_parts(ctx, {wire_part(VAR_DOUBLE, pk64, sig96, msg32, n, status, c32_or_null, nullptr)});
}
JJS_API int jjs_verify_vargen(jjs_ctx* ctx, const uint8_t* pk64, const uint8_t* sig64, const uint8_t* msg32, size_t n, uint8_t* status,
                              uint8_t* c32_or_null) {
    return run_parts(ctx, {wire_part(VAR_VARGEN, pk64, sig64, msg32, n, status, c32_or_null, nullptr)});
}
JJS_API int jjs_verify_aggregate(jjs_ctx* ctx, const uint8_t* pks32, const uint32_t* offsets, const uint8_t* sig64, const uint8_t* msg32, size_t n,
                                 uint8_t* status, uint8_t* c32_or_null, uint8_t* aggpk32_or_null) {
    if (ctx && ctx->ready && n && !status) return fail(ctx, JJS_ERR_ARGUMENT, "null buffer");
    return run_parts(ctx, {aggregate_part(pks32, offsets, sig64, msg32, n, status, c32_or_null, aggpk32_or_null, nullptr)});
}
JJS_API int jjs_verify_batch(jjs_ctx* ctx, const uint8_t* pk32, const uint8_t* sig64, const uint8_t* msg32, size_t n, uint32_t* accept_bitmap) {
    if (ctx && ctx->ready && n && !accept_bitmap) return fail(ctx, JJS_ERR_ARGUMENT, "null buffer");
    return run_parts(ctx, {wire_part(VAR_SINGLE, pk32, sig64, msg32, n, nullptr, nullptr, accept_bitmap)});
}
JJS_API int jjs_verify_batch_double(jjs_ctx* ctx, const uint8_t* pk64, const uint8_t* sig96, const uint8_t* msg32, size_t n, uint32_t* accept_bitmap) {
    if (ctx && ctx->ready && n && !accept_bitmap) return fail(ctx, JJS_ERR_ARGUMENT, "null buffer");
    return run_parts(ctx, {wire_part(VAR_DOUBLE, pk64, sig96, msg32, n, nullptr, nullptr, accept_bitmap)});
}
JJS_API int jjs_verify_batch_vargen(jjs_ctx* ctx, const uint8_t* pk64, const uint8_t* sig64, const uint8_t* msg32, size_t n, uint32_t* accept_bitmap) {
    if (ctx && ctx->ready && n && !accept_bitmap) return fail(ctx, JJS_ERR_ARGUMENT, "null buffer");
    return run_parts(ctx, {wire_part(VAR_VARGEN, pk64, sig64, msg32, n, nullptr, nullptr, accept_bitmap)});
}
JJS_API int jjs_verify_batch_aggregate(jjs_ctx* ctx, const uint8_t* pks32, const uint32_t* offsets, const uint8_t* sig64, const uint8_t* msg32, size_t n,
                                       uint32_t* accept_bitmap) {
    if (ctx && ctx->ready && n && !accept_bitmap) return fail(ctx, JJS_ERR_ARGUMENT, "null buffer");
    return run_parts(ctx, {aggregate_part(pks32, offsets, sig64, msg32, n, nullptr, nullptr, nullptr, accept_bitmap)});
}
JJS_API int jjs_verify_mixed(jjs_ctx* ctx, const jjs_part* parts, size_t n_parts) {
    JJS_ENTER(ctx);
    if (n_parts && !parts) return fail(ctx, JJS_ERR_ARGUMENT, "null buffer");
    std::vector<HostPart> hp;
    hp.reserve(n_parts);
    for (size_t i = 0; i < n_parts; i++) {
        const jjs_part& p = parts[i];
        if (p.kind < JJS_KIND_SINGLE || p.kind > JJS_KIND_AGGREGATE) return fail(ctx, JJS_ERR_ARGUMENT, "part %zu: bad kind", i);
        if (p.kind == JJS_KIND_AGGREGATE) hp.push_back(aggregate_part(p.pk, p.offsets, p.sig, p.msg32, p.n, p.status, p.c32, p.aggpk32, p.accept_bitmap));
        else hp.push_back(wire_part(p.kind, p.pk, p.sig, p.msg32, p.n, p.status, p.c32, p.accept_bitmap));
    }
    return run_parts(ctx, hp);
}
JJS_API int jjs_status_bitmap_device(jjs_ctx* ctx, int device_index, const uint8_t* d_status, size_t n, uint32_t* d_accept_bitmap, void* cuda_stream) {
    JJS_ENTER(ctx);
    if (device_index < 0 || (size_t)device_index >= ctx->dev.size()) return fail(ctx, JJS_ERR_ARGUMENT, "device_index out of range");
    if (n == 0) return JJS_SUCCESS;
    if (!d_status || !d_accept_bitmap) return fail(ctx, JJS_ERR_ARGUMENT, "null buffer");
    JJS_CUDA(ctx, cudaSetDevice(ctx->dev[device_index].device));
    k_bitmap<<<blocks_for(n), BLOCK, 0, (cudaStream_t)cuda_stream>>>(d_status, n, d_accept_bitmap);
    ctx->launches++;
    JJS_CUDA(ctx, cudaGetLastError());
    return JJS_SUCCESS;
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
JJS_API int jjs_verify_aggregate_device(jjs_ctx* ctx, int device_index, const uint8_t* d_pks32, const uint32_t* d_offsets, const uint32_t* h_offsets,
                                        const uint8_t* d_sig64, const uint8_t* d_msg32, size_t n, uint8_t* d_status, uint8_t* d_c32_or_null,
                                        uint8_t* d_aggpk32_or_null, void* cuda_stream) {
    JJS_ENTER(ctx);
    if (device_index < 0 || (size_t)device_index >= ctx->dev.size()) return fail(ctx, JJS_ERR_ARGUMENT, "device_index out of range");
    if (n == 0) return JJS_SUCCESS;
    if (!d_offsets || !h_offsets || !d_sig64 || !d_msg32 || !d_status) return fail(ctx, JJS_ERR_ARGUMENT, "null buffer");
    for (size_t i = 0; i < n; i++)
        if (h_offsets[i + 1] < h_offsets[i]) return fail(ctx, JJS_ERR_ARGUMENT, "offsets must be non-decreasing");
    if (h_offsets[n] > h_offsets[0] && !d_pks32) return fail(ctx, JJS_ERR_ARGUMENT, "null key buffer");
    return run_aggregate_device(ctx, ctx->dev[device_index], d_pks32, d_offsets, h_offsets, d_sig64, d_msg32, n, d_status, d_c32_or_null,
                                d_aggpk32_or_null, (cudaStream_t)cuda_stream);
}
JJS_API int jjs_verify_ext(jjs_ctx* ctx, int variant, const uint8_t* points_ext160, const uint8_t* u32, const uint8_t* msg32, size_t n,
                           uint8_t* status, uint8_t* c32_or_null) {
    if (ctx && ctx->ready && n && !status) return fail(ctx, JJS_ERR_ARGUMENT, "null buffer");
    return run_parts(ctx, {HostPart{PART_TYPED, variant, points_ext160, u32, msg32, nullptr, n, status, c32_or_null, nullptr, nullptr, false}});
}
JJS_API int jjs_challenge_only(jjs_ctx* ctx, int variant, const uint8_t* pk, const uint8_t* sig, const uint8_t* msg32, size_t n, uint8_t* c32) {
    return run_parts(ctx, {wire_part(variant, pk, sig, msg32, n, nullptr, c32, nullptr, true)});
}
JJS_API int jjs_subgroup_check(jjs_ctx* ctx, const uint8_t* points32, size_t n, int method, uint8_t* out) {
    JJS_ENTER(ctx);
    if (method < 0 || method > 1) return fail(ctx, JJS_ERR_ARGUMENT, "bad method");
    if (n == 0) return JJS_SUCCESS;
    if (!points32 || !out) return fail(ctx, JJS_ERR_ARGUMENT, "null buffer");
    int rc = ensure_scratch(ctx, ctx->dev[0]);
    if (rc) return rc;
    const void* srcs[1] = {points32};
    const size_t bytes[1] = {32 * n};
    return run_small(ctx, "subgroup check", 32 * n, n, srcs, bytes, 1, out, [&](DeviceState& d, uint8_t* in, uint8_t* o) {
        for (size_t first = 0; first < n; first += TAB_THREADS) {
            size_t cnt = n - first < TAB_THREADS ? n - first : TAB_THREADS;
            k_subgroup_check<<<blocks_for(cnt), BLOCK, 0, d.stream>>>(WireField{in, 32}, first, cnt, method, o, d.tab, TAB_THREADS, d.tables());
            ctx->launches++;
        }
    });
}
JJS_API int jjs_fb_table_check(jjs_ctx* ctx, int which, const uint32_t* entries, size_t n, uint32_t* mismatches) {
    JJS_ENTER(ctx);
    if (which < 0 || which > 1) return fail(ctx, JJS_ERR_ARGUMENT, "bad table");
    if (!mismatches) return fail(ctx, JJS_ERR_ARGUMENT, "null buffer");
    *mismatches = 0;
    if (n == 0) return JJS_SUCCESS;
    if (!entries) return fail(ctx, JJS_ERR_ARGUMENT, "null buffer");
    const void* srcs[1] = {entries};
    const size_t bytes[1] = {4 * n};
    uint32_t count = 0;
    int rc = run_small(ctx, "fixed-base table check", 4 * n, 4, srcs, bytes, 1, &count, [&](DeviceState& d, uint8_t* in, uint8_t* o) {
        cudaMemsetAsync(o, 0, 4, d.stream);
        k_fb_check<<<blocks_for(n), BLOCK, 0, d.stream>>>(which ? d.fb_gn : d.fb_g, which, reinterpret_cast<const uint32_t*>(in), (int)n,
                                                         reinterpret_cast<uint32_t*>(o));
        ctx->launches++;
    });
    *mismatches = count;
    return rc;
}
JJS_API int jjs_sign_aggregate_batch(jjs_ctx* ctx, const uint8_t* sk32, const uint32_t* offsets, const uint8_t* rnd32, const uint8_t* msg32, size_t n,
                                     uint8_t* pks32_out, uint8_t* sig64_out) {
    JJS_ENTER(ctx);
    if (n == 0) return JJS_SUCCESS;
    if (!sk32 || !offsets || !rnd32 || !msg32 || !pks32_out || !sig64_out) return fail(ctx, JJS_ERR_ARGUMENT, "null buffer");
    if (offsets[0] != 0) return fail(ctx, JJS_ERR_ARGUMENT, "offsets[0] must be 0");
    for (size_t i = 0; i < n; i++)
        if (offsets[i + 1] < offsets[i]) return fail(ctx, JJS_ERR_ARGUMENT, "offsets must be non-decreasing");
    const size_t K = offsets[n];
    const size_t off_bytes = align_up(4 * (n + 1), 32);
    const void* srcs[4] = {sk32, offsets, rnd32, msg32};
    const size_t bytes[4] = {32 * K, off_bytes, 32 * n, 32 * n};   // the offsets slot is padded; only 4 (n + 1) bytes are meaningful
    // the padded tail of the offsets copy must stay inside the caller's array: copy exactly, pad on the device side
    const size_t exact[4] = {32 * K, 4 * (n + 1), 32 * n, 32 * n};
    DeviceState& d = ctx->dev[0];
    JJS_CUDA(ctx, cudaSetDevice(d.device));
    const size_t in_bytes = bytes[0] + bytes[1] + bytes[2] + bytes[3];
    int rc = ensure_stage(ctx, d, align_up(in_bytes, 256) + 32 * K + 64 * n);
    if (rc) return rc;
    cudaError_t ce = cudaSuccess;
    size_t at = 0;
    for (int i = 0; i < 4 && ce == cudaSuccess; i++) {
        if (exact[i]) ce = cudaMemcpyAsync(d.stage + at, srcs[i], exact[i], cudaMemcpyHostToDevice, d.stream);
        at += bytes[i];
    }
    uint8_t *b = d.stage, *o_pks = d.stage + align_up(in_bytes, 256), *o_sig = o_pks + 32 * K;
    if (ce == cudaSuccess) {
        k_sign_aggregate<<<blocks_for(n), BLOCK, 0, d.stream>>>(b, reinterpret_cast<uint32_t*>(b + bytes[0]), b + bytes[0] + bytes[1],
                                                              b + bytes[0] + bytes[1] + bytes[2], n, o_pks, o_sig, d.tables());
        ctx->launches++;
        ce = cudaGetLastError();
    }
    if (ce == cudaSuccess && K) ce = cudaMemcpyAsync(pks32_out, o_pks, 32 * K, cudaMemcpyDeviceToHost, d.stream);
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(sig64_out, o_sig, 64 * n, cudaMemcpyDeviceToHost, d.stream);
    cudaError_t e = cudaStreamSynchronize(d.stream);
    if (ce != cudaSuccess) e = ce;
    if (e != cudaSuccess) return fail(ctx, JJS_ERR_CUDA, "sign aggregate batch failed: %s", cudaGetErrorString(e));
    return JJS_SUCCESS;
}
JJS_API int jjs_points_to_ext(jjs_ctx* ctx, const uint8_t* points32, const uint8_t* z_mont32, size_t n, uint8_t* out160) {
    JJS_ENTER(ctx);
    if (n == 0) return JJS_SUCCESS;
    if (!points32 || !z_mont32 || !out160) return fail(ctx, JJS_ERR_ARGUMENT, "null buffer");
    const void* srcs[2] = {points32, z_mont32};
    const size_t bytes[2] = {32 * n, 32 * n};
    return run_small(ctx, "points_to_ext", 64 * n, 160 * n, srcs, bytes, 2, out160, [&](DeviceState& d, uint8_t* in, uint8_t* o) {
        k_points_to_ext<<<blocks_for(n), BLOCK, 0, d.stream>>>(in, in + 32 * n, n, o, d.tables());
        ctx->launches++;
    });
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
    std::vector<StageRecord> records;
    {
        std::lock_guard<std::mutex> lock(ctx->mu);
        records.swap(ctx->records);
    }
    int rc = JJS_SUCCESS;
    for (auto& r : records) {
        cudaSetDevice(r.device);
        float ms = 0;
        cudaError_t e = cudaEventSynchronize(r.e1);
        if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, r.e0, r.e1);
        cudaEventDestroy(r.e0);
        cudaEventDestroy(r.e1);
        if (e != cudaSuccess) {
            if (rc == JJS_SUCCESS) rc = fail(ctx, JJS_ERR_CUDA, "profile collect: %s", cudaGetErrorString(e));
            continue;
        }
        stage_ms[r.stage] += ms;
        stage_count[r.stage]++;
    }
    return rc;
}
JJS_API int jjs_sign_batch(jjs_ctx* ctx, int variant, const uint8_t* sk32, const uint8_t* rnd32, const uint8_t* gen_scalar32_or_null,
                           const uint8_t* msg32, size_t n, uint8_t* pk_out, uint8_t* sig_out) {
    return run_sign(ctx, variant, sk32, rnd32, gen_scalar32_or_null, msg32, n, pk_out, sig_out);
}

}  // extern "C"
