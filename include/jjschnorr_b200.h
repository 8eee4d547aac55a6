/*
 * jjschnorr_b200 -- C ABI of the B200-native batch verifier for dusk-network/jubjub-schnorr signatures.
 *
 * This is the drop-in boundary for the reference's verify path: every entry point consumes the
 * reference's own wire encodings and reproduces the result of the reference call it replaces, bit for
 * bit.  A Rust caller binds these through `extern "C"`; nothing here exposes C++
 * or torch types, nothing throws, and there is no CPU fallback: if no usable CUDA device is present the
 * calls return JJS_ERR_CUDA.
 *
 * Wire encodings (all little-endian, exactly `to_bytes()` of the reference types):
 *   PublicKey        32 B  compressed point                       reference src/keys/public.rs:80-94
 *   Signature        64 B  u (32) || R (32)                       reference src/signatures.rs:101-118
 *   PublicKeyDouble  64 B  pk || pk'                              reference src/keys/public/double.rs:169-186
 *   SignatureDouble  96 B  u || R || R'                           reference src/signatures/double.rs:122-147
 *   PublicKeyVarGen  64 B  pk || generator                        reference src/keys/public/var_gen.rs:54-79
 *   SignatureVarGen  64 B  u || R                                 reference src/signatures/var_gen.rs:95-112
 *   message          32 B  BlsScalar::to_bytes()
 * Arrays are contiguous, item i at base + i * size.
 *
 * Per-item result (status byte), mirroring Result<(), Error> of the reference (src/error.rs:13-26):
 *   JJS_OK                 Ok(())
 *   JJS_INVALID_SIGNATURE  Err(Error::InvalidSignature)   the verification equation does not hold
 *   JJS_INVALID_POINT      Err(Error::InvalidPoint)       a key / signature point fails is_valid()
 *   JJS_BYTES_ERROR        Err(Error::BytesError(_))      some field fails from_bytes()
 * with the reference's precedence: decoding, then point validity, then the equation.
 * The optional challenge output receives c = challenge_hash(..) as JubJubScalar::to_bytes() for items
 * whose status is JJS_OK or JJS_INVALID_SIGNATURE and 32 zero bytes otherwise (the reference does not
 * compute a challenge for those).
 */
#ifndef JJSCHNORR_B200_H
#define JJSCHNORR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JJS_OK 0
#define JJS_INVALID_SIGNATURE 1
#define JJS_INVALID_POINT 2
#define JJS_BYTES_ERROR 3
/* multisig::combine only */
#define JJS_INVALID_MULTISIG_TRANSCRIPT 4
#define JJS_INVALID_MULTISIG_SHARE 5

/* call return codes */
#define JJS_SUCCESS 0
#define JJS_ERR_ARGUMENT (-1)
#define JJS_ERR_CUDA (-2)
#define JJS_ERR_NOMEM (-3)

typedef struct jjs_ctx jjs_ctx;

/* Create a context over the given CUDA device ordinals (n_devices >= 1; devices == NULL means device 0
 * .. n_devices-1).  Builds the per-device constant tables -- among them the fixed-base window tables for the two generators,
 * 2.4 GB each (12 windows of 2^21 points; ~70 ms per device) -- and allocates no batch memory yet; the first verification adds
 * ~9.3 GiB of scratch per device (sized for 2^20 items in flight, whatever the batch size).
 * On failure *out still receives a context, for jjs_last_error only: every other entry point returns JJS_ERR_CUDA on
 * it (there is no CPU fallback) and jjs_destroy frees it. */
int jjs_init(const int* devices, int n_devices, jjs_ctx** out);
void jjs_destroy(jjs_ctx* ctx);
/* Human-readable description of the last failure on this context ("" if none). */
const char* jjs_last_error(const jjs_ctx* ctx);
int jjs_device_count(const jjs_ctx* ctx);

/* ---- host-buffer entry points: shard the batch contiguously over the context's devices, copy in,
 *      verify, copy the status bytes (and challenges) back.  All pointers are host memory (pinned or pageable: every
 *      device of the context is driven by its own host thread for the duration of the call, so staged copies from
 *      pageable memory on one device do not hold up the others).  The calls return when everything is done; if a call
 *      fails, the devices are drained first, so nothing still reads or writes the caller's buffers. -------------- */

/* PublicKey::verify for n (key, signature, message) triples.  reference src/keys/public.rs:114-135 */
int jjs_verify_single(jjs_ctx* ctx, const uint8_t* pk32, const uint8_t* sig64, const uint8_t* msg32, size_t n,
                      uint8_t* status, uint8_t* c32_or_null);
/* NEW in this library (BASELINE.json north_star): verify_batch(&[(PublicKey, Signature, BlsScalar)]) -> Vec<bool>.
 * Same work as jjs_verify_single; the result comes back as a packed accept bitmap, bit (i % 32) of word i / 32 set iff
 * PublicKey::verify(pk_i, sig_i, msg_i) is Ok (src/keys/public.rs:114-135).  accept_bitmap has (n + 31) / 32 words;
 * unused high bits of the last word are 0.  The bits are packed on the GPU with one warp ballot per word. */
int jjs_verify_batch(jjs_ctx* ctx, const uint8_t* pk32, const uint8_t* sig64, const uint8_t* msg32, size_t n,
                     uint32_t* accept_bitmap);
/* The same packed accept bitmap for the other three kinds. */
int jjs_verify_batch_double(jjs_ctx* ctx, const uint8_t* pk64, const uint8_t* sig96, const uint8_t* msg32, size_t n,
                            uint32_t* accept_bitmap);
int jjs_verify_batch_vargen(jjs_ctx* ctx, const uint8_t* pk64, const uint8_t* sig64, const uint8_t* msg32, size_t n,
                            uint32_t* accept_bitmap);
int jjs_verify_batch_aggregate(jjs_ctx* ctx, const uint8_t* pks32, const uint32_t* offsets, const uint8_t* sig64,
                               const uint8_t* msg32, size_t n, uint32_t* accept_bitmap);
/* PublicKeyDouble::verify.  reference src/keys/public/double.rs:86-117 */
int jjs_verify_double(jjs_ctx* ctx, const uint8_t* pk64, const uint8_t* sig96, const uint8_t* msg32, size_t n,
                      uint8_t* status, uint8_t* c32_or_null);
/* PublicKeyVarGen::verify.  reference src/keys/public/var_gen.rs:107-133 */
int jjs_verify_vargen(jjs_ctx* ctx, const uint8_t* pk64, const uint8_t* sig64, const uint8_t* msg32, size_t n,
                      uint8_t* status, uint8_t* c32_or_null);
/* multisig::aggregate_pk(&pks[offsets[i]..offsets[i+1]]).verify(sig_i, msg_i).
 * reference src/multisig.rs:154-156, 393-429 then src/keys/public.rs:114-135.
 * offsets has n + 1 entries; aggpk32_or_null receives PublicKey::to_bytes() of each aggregate key.  Any number of signers
 * per item, as in the reference (the sponge tag of a 2 + 2 n element transcript is computed at call time); an item without
 * signers aggregates to the identity and is reported JJS_INVALID_POINT, which is what verify() returns for that key. */
int jjs_verify_aggregate(jjs_ctx* ctx, const uint8_t* pks32, const uint32_t* offsets, const uint8_t* sig64,
                         const uint8_t* msg32, size_t n, uint8_t* status, uint8_t* c32_or_null,
                         uint8_t* aggpk32_or_null);

/* Several batches of different kinds in ONE call (BASELINE.json configs[3] and [4]: var-generator + aggregate-key items,
 * single + double items).  Every part is split evenly over the context's devices (parts below 8 192 items per device go
 * whole to the least loaded device, weighted by the cost of their kind), so each device carries the same share of every
 * kind and all devices finish together; nothing is synchronised between the parts.  Per part, any of status / c32 /
 * aggpk32 / accept_bitmap may be NULL, but status and accept_bitmap not both. */
#define JJS_KIND_SINGLE 0
#define JJS_KIND_DOUBLE 1
#define JJS_KIND_VARGEN 2
#define JJS_KIND_AGGREGATE 3
typedef struct jjs_part {
    int kind;                 /* JJS_KIND_* */
    const uint8_t* pk;        /* kind 0: pk32;  1, 2: pk64;  3: pks32 (ragged, see offsets) */
    const uint32_t* offsets;  /* kind 3: n + 1 entries;  NULL otherwise */
    const uint8_t* sig;       /* kind 1: sig96;  otherwise sig64 */
    const uint8_t* msg32;
    size_t n;
    uint8_t* status;          /* n status bytes */
    uint8_t* c32;             /* n challenges (32 bytes each) */
    uint8_t* aggpk32;         /* kind 3: n aggregate keys */
    uint32_t* accept_bitmap;  /* (n + 31) / 32 words, as jjs_verify_batch */
} jjs_part;
int jjs_verify_mixed(jjs_ctx* ctx, const jjs_part* parts, size_t n_parts);

/* ---- device-buffer entry points: inputs and outputs already live on device `device_index` of the context
 *      (an index into the list given to jjs_init); the work is ordered after everything `cuda_stream` holds at
 *      the call (a cudaStream_t used as is: NULL is the legacy default stream) and is complete when that stream
 *      is: batches above 2^18 items run as sub-chunks on two streams of the context, forked from and joined back
 *      into `cuda_stream` with events.  Calls on one context share its scratch memory, so keep them on one stream
 *      or order them yourself.  No host copies. --- */
int jjs_verify_single_device(jjs_ctx* ctx, int device_index, const uint8_t* d_pk32, const uint8_t* d_sig64,
                             const uint8_t* d_msg32, size_t n, uint8_t* d_status, uint8_t* d_c32_or_null,
                             void* cuda_stream);
int jjs_verify_double_device(jjs_ctx* ctx, int device_index, const uint8_t* d_pk64, const uint8_t* d_sig96,
                             const uint8_t* d_msg32, size_t n, uint8_t* d_status, uint8_t* d_c32_or_null,
                             void* cuda_stream);
int jjs_verify_vargen_device(jjs_ctx* ctx, int device_index, const uint8_t* d_pk64, const uint8_t* d_sig64,
                             const uint8_t* d_msg32, size_t n, uint8_t* d_status, uint8_t* d_c32_or_null,
                             void* cuda_stream);

/* Pack n status bytes (any *_device entry point's output) into the accept bitmap described at jjs_verify_batch;
 * enqueued on `cuda_stream`. */
int jjs_status_bitmap_device(jjs_ctx* ctx, int device_index, const uint8_t* d_status, size_t n, uint32_t* d_accept_bitmap,
                             void* cuda_stream);

/* Aggregate-key verification on device buffers.  d_offsets / h_offsets: the same n + 1 offsets on the device and on
 * the host (the host copy plans the chunks and is only read during the call).  Enqueue-only like the other *_device
 * calls: the ordering of a chunk by signer count happens on the device.  The one exception is growth of the library's
 * key scratch or tag table (the first aggregate call on a device, more than 2^21 signer keys in one 2^18-item chunk, or
 * more than 511 signers in one item), which waits for the device once. */
int jjs_verify_aggregate_device(jjs_ctx* ctx, int device_index, const uint8_t* d_pks32, const uint32_t* d_offsets,
                                const uint32_t* h_offsets, const uint8_t* d_sig64, const uint8_t* d_msg32, size_t n,
                                uint8_t* d_status, uint8_t* d_c32_or_null, uint8_t* d_aggpk32_or_null, void* cuda_stream);

/* Typed inputs (what a Rust caller holds before any to_bytes()): every point is given by its JubJubExtended
 * coordinates (u, v, z, t1, t2), each the in-memory BlsScalar of the reference (4 x u64 little-endian Montgomery limbs,
 * R = 2^256): 160 bytes per point.  points_ext160 is item-major, the points of one item in the order
 *   variant 0: PK, R        variant 1: PK, PK', R, R'        variant 2: PK, generator, R
 * u32 = JubJubScalar::to_bytes() of the signature scalar, msg32 = BlsScalar::to_bytes().  The GPU normalises the points
 * (no per-point inversion on the CPU), applies is_valid() exactly as verify() does for typed values -- z != 0, curve
 * equation, t1 t2 consistency, identity, torsion -- and then runs the same pipeline.  Status 3 is returned when u32 / msg32
 * is not canonical or a coordinate is not a reduced field element.  Host buffers. */
int jjs_verify_ext(jjs_ctx* ctx, int variant, const uint8_t* points_ext160, const uint8_t* u32, const uint8_t* msg32,
                   size_t n, uint8_t* status, uint8_t* c32_or_null);

/* Test-data utility for the typed path: wire point i -> the 160-byte coordinates (u z, v z, z, u z, v) with the caller's
 * Montgomery value z (z_mont32, any non-zero reduced field element); zeros if the point does not decode.  Host buffers. */
int jjs_points_to_ext(jjs_ctx* ctx, const uint8_t* points32, const uint8_t* z_mont32, size_t n, uint8_t* out160);

/* multisig::combine for n sessions (reference src/multisig.rs:311-347; the share check is verify_share,
 * src/multisig.rs:255-281, 366-387).  Session i owns participants offsets[i] .. offsets[i+1]-1 of the ragged arrays
 * pks32 (PublicKey::to_bytes), R32 / S32 (compressed commitment points) and z32 (JubJubScalar::to_bytes shares);
 * msg32 has one BlsScalar per session; any number of participants per session.  Per session:
 *   status 0: Ok(Signature), sig64 = (sum z_i) || RSa       5: Err(InvalidMultisigShare(bad_index))
 *          3: some field fails from_bytes                     4: Err(InvalidMultisigTranscript) (no participants)
 * share_ok (one byte per participant) is what verify_share returns for every participant, not only the first bad
 * one.  Like the reference, no subgroup validation is applied to any point.  Host buffers, device 0. */
int jjs_multisig_combine(jjs_ctx* ctx, const uint8_t* pks32, const uint8_t* R32, const uint8_t* S32, const uint8_t* z32,
                         const uint32_t* offsets, const uint8_t* msg32, size_t n, uint8_t* share_ok_or_null, uint8_t* status,
                         uint32_t* bad_index_or_null, uint8_t* sig64_or_null);

/* Challenge hash only (hash parity hook): c = challenge_hash(..) for already-valid encodings, no curve
 * check.  variant: 0 single, 1 double, 2 var-generator.  Host buffers. */
int jjs_challenge_only(jjs_ctx* ctx, int variant, const uint8_t* pk, const uint8_t* sig, const uint8_t* msg32,
                       size_t n, uint8_t* c32);

/* is_torsion_free() of n wire-encoded points (reference src/keys/public.rs:160), by the production method
 * (method 0: order-8 Tate pairing) or by the definition (method 1: [r]P == identity by scalar multiplication).
 * out[i] = 1 / 0, or 0xff if the encoding does not decode.  Host buffers; a cross-check hook for the tests. */
int jjs_subgroup_check(jjs_ctx* ctx, const uint8_t* points32, size_t n, int method, uint8_t* out);

/* Fixed-base table self-check (a cross-check hook for the tests, like jjs_subgroup_check): the library builds its window tables
 * for G (which = 0) and G' (which = 1) incrementally; this recomputes the n table entries named by `entries` (flat indices,
 * window * 2^width + digit, reduced modulo the table size) from their definition digit * 2^(width * window) * B by plain
 * double-and-add on device 0 and counts the entries that differ.  Host buffers. */
int jjs_fb_table_check(jjs_ctx* ctx, int which, const uint32_t* entries, size_t n, uint32_t* mismatches);

/* Batch key derivation + signing on device 0 of the context (used to make large
 * synthetic batches and by the tests).  For each item: pk = PublicKey::from(&sk) (reference
 * src/keys/public.rs:54-60; PublicKeyDouble / PublicKeyVarGen likewise) and sig = sk.sign(rng, msg) with the
 * reference's hedged nonce (src/keys/secret.rs:174-194, src/nonce.rs:32-86; sign_double: src/keys/secret/double.rs:57-85;
 * var-gen: src/keys/secret/var_gen.rs), where rnd32 is the JubJubScalar the reference draws from its RNG
 * and, for variant 2, the generator is gen_scalar * GENERATOR_EXTENDED as in SecretKeyVarGen::random.
 * All scalars are canonical 32-byte little-endian (sk, rnd, gen_scalar < r; msg < q); an item with an
 * out-of-range input yields all-zero outputs.  Host buffers.  NOT constant-time: test data only. */
int jjs_sign_batch(jjs_ctx* ctx, int variant, const uint8_t* sk32, const uint8_t* rnd32,
                   const uint8_t* gen_scalar32_or_null, const uint8_t* msg32, size_t n, uint8_t* pk_out,
                   uint8_t* sig_out);

/* Synthetic aggregate-key items: signer keys pk_j = sk_j * G (written to pks32_out, ragged like sk32 / offsets, at most 8
 * signers per item), and a hedged Schnorr signature under the aggregate secret sum_j d_j sk_j with the reference's
 * delinearisation coefficients (src/multisig.rs:393-409) -- i.e. what jjs_verify_aggregate accepts.  Host buffers. */
int jjs_sign_aggregate_batch(jjs_ctx* ctx, const uint8_t* sk32, const uint32_t* offsets, const uint8_t* rnd32,
                             const uint8_t* msg32, size_t n, uint8_t* pks32_out, uint8_t* sig64_out);

/* Per-stage device timing (CUDA events on the launching stream around each pipeline stage):
 * stage 0 point decode + subgroup test of the keys, 1 challenge hash, 2 key aggregation (aggregate path only),
 * 3 verification equations, 4 status, 5 deferred subgroup tests of signature points (only those the equation did not
 * already settle).
 * jjs_profile_collect waits for the recorded events, adds their durations (ms) and occurrence counts per
 * stage into the two JJS_N_STAGES-long arrays, and clears the records. */
#define JJS_N_STAGES 6
void jjs_profile_enable(jjs_ctx* ctx, int on);
int jjs_profile_collect(jjs_ctx* ctx, double* stage_ms, uint64_t* stage_count);

/* Kernels launched by this context since creation (for the bench's gpu_launches accounting). */
uint64_t jjs_launch_count(const jjs_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* JJSCHNORR_B200_H */
