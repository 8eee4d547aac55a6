// Extracted from INTEGRATION.md by tools/extract_rust_shim.py -- edit the markdown, not this file.
// NOT compiled in this repository's environment (no cargo/rustc in the image).

use crate::{Error, PublicKey, PublicKeyDouble, PublicKeyVarGen, Signature, SignatureDouble, SignatureVarGen};
use dusk_bls12_381::BlsScalar;
use dusk_bytes::Serializable;

pub struct GpuVerifier { ctx: *mut ffi::jjs_ctx }
unsafe impl Send for GpuVerifier {}       // one host thread at a time per context (header: threading)

impl GpuVerifier {
    /// `devices`: CUDA ordinals; the batch is split into contiguous shards, one stream per GPU, no collective.
    pub fn new(devices: &[i32]) -> Result<Self, String> {
        let mut ctx = core::ptr::null_mut();
        let rc = unsafe { ffi::jjs_init(devices.as_ptr(), devices.len() as i32, &mut ctx) };
        if rc != 0 {
            let msg = unsafe { std::ffi::CStr::from_ptr(ffi::jjs_last_error(ctx)) }.to_string_lossy().into_owned();
            unsafe { ffi::jjs_destroy(ctx) };
            return Err(msg);
        }
        Ok(Self { ctx })
    }

    /// NEW: verify_batch(&[(PublicKey, Signature, BlsScalar)]) -> Vec<bool>, bit-exact with
    /// `items.iter().map(|(pk, sig, m)| pk.verify(sig, *m).is_ok())`.
    /// Goes through `jjs_verify_batch`, which returns the accept bits packed on the GPU (one warp ballot per 32 items).
    pub fn verify_batch(&mut self, items: &[(PublicKey, Signature, BlsScalar)]) -> Vec<bool> {
        let n = items.len();
        let (pk, sig, msg) = Self::wire(items);                     // the three to_bytes() loops of verify_batch_status
        let mut words = vec![0u32; (n + 31) / 32];
        let rc = unsafe { ffi::jjs_verify_batch(self.ctx, pk.as_ptr(), sig.as_ptr(), msg.as_ptr(), n, words.as_mut_ptr()) };
        assert_eq!(rc, 0, "jjs_verify_batch failed");
        (0..n).map(|i| words[i / 32] >> (i % 32) & 1 == 1).collect()
    }

    /// Same, keeping the reference's error variant per item.
    pub fn verify_batch_status(&mut self, items: &[(PublicKey, Signature, BlsScalar)]) -> Vec<Result<(), Error>> {
        let n = items.len();
        let (mut pk, mut sig, mut msg) = (vec![0u8; 32 * n], vec![0u8; 64 * n], vec![0u8; 32 * n]);
        for (i, (p, s, m)) in items.iter().enumerate() {
            pk[32 * i..32 * i + 32].copy_from_slice(&p.to_bytes());    // src/keys/public.rs:83-85
            sig[64 * i..64 * i + 64].copy_from_slice(&s.to_bytes());   // src/signatures.rs:104-109
            msg[32 * i..32 * i + 32].copy_from_slice(&m.to_bytes());
        }
        let mut status = vec![0u8; n];
        let rc = unsafe { ffi::jjs_verify_single(self.ctx, pk.as_ptr(), sig.as_ptr(), msg.as_ptr(), n,
                                                 status.as_mut_ptr(), core::ptr::null_mut()) };
        assert_eq!(rc, 0, "jjs_verify_single failed");
        status.into_iter().map(|s| match s {
            ffi::JJS_OK => Ok(()),
            ffi::JJS_INVALID_SIGNATURE => Err(Error::InvalidSignature),
            ffi::JJS_INVALID_POINT => Err(Error::InvalidPoint),
            _ => Err(Error::BytesError(dusk_bytes::Error::InvalidData)),
        }).collect()
    }
    // verify_batch_double(&[(PublicKeyDouble, SignatureDouble, BlsScalar)]) -> jjs_verify_double   (64 / 96 / 32 bytes)
    // verify_batch_var_gen(&[(PublicKeyVarGen, SignatureVarGen, BlsScalar)]) -> jjs_verify_vargen  (64 / 64 / 32 bytes)
    // verify_batch_aggregate(&[(&[PublicKey], Signature, BlsScalar)])       -> jjs_verify_aggregate (ragged keys + offsets)
}

impl Drop for GpuVerifier { fn drop(&mut self) { unsafe { ffi::jjs_destroy(self.ctx) } } }

/// Scalar drop-ins: same signature and result as the reference methods they replace.
impl PublicKey {
    pub fn verify_gpu(&self, gpu: &mut GpuVerifier, sig: &Signature, message: BlsScalar) -> Result<(), Error> {
        gpu.verify_batch_status(&[(*self, *sig, message)]).pop().unwrap()      // replaces src/keys/public.rs:114-135
    }
}

// ---- typed inputs (jjs_verify_ext) ----
fn push_point(buf: &mut Vec<u8>, p: &JubJubExtended) {
    // BlsScalar is `pub struct Scalar(pub [u64; 4])`: Montgomery limbs, little-endian; no arithmetic needed here.
    for c in [p.get_u(), p.get_v(), p.get_z(), p.get_t1(), p.get_t2()] {
        for limb in c.0 { buf.extend_from_slice(&limb.to_le_bytes()); }
    }
}

pub fn verify_batch_typed(&mut self, items: &[(PublicKey, Signature, BlsScalar)]) -> Vec<Result<(), Error>> {
    let n = items.len();
    let (mut pts, mut u, mut msg) = (Vec::with_capacity(320 * n), Vec::with_capacity(32 * n), Vec::with_capacity(32 * n));
    for (pk, sig, m) in items {
        push_point(&mut pts, pk.as_ref());        // src/keys/public.rs:73-77
        push_point(&mut pts, sig.R());            // src/signatures.rs:74-76
        u.extend_from_slice(&sig.u().to_bytes()); // one Montgomery reduction, no inversion
        msg.extend_from_slice(&m.to_bytes());
    }
    let mut status = vec![0u8; n];
    let rc = unsafe { ffi::jjs_verify_ext(self.ctx, 0, pts.as_ptr(), u.as_ptr(), msg.as_ptr(), n, status.as_mut_ptr(), core::ptr::null_mut()) };
    assert_eq!(rc, 0);
    status.into_iter().map(status_to_result).collect()
}
