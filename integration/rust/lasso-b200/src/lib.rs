//! Safe Rust surface over `include/lasso_b200.h`.  Method names and argument meaning follow the reference
//! (a16z/Lasso): `DensifiedRepresentation::from_lookup_indices` (src/lasso/densified.rs:22), `.commit`
//! (densified.rs:78), `SparsePolyCommitmentGens::new` (src/lasso/surge.rs:32),
//! `SparsePolynomialEvaluationProof::prove` (surge.rs:119).  Outputs are the ark-serialize (compressed) bytes of
//! the reference's structs: `SparsePolynomialCommitment::deserialize_compressed(&bytes[..])` /
//! `SparsePolynomialEvaluationProof::deserialize_compressed(..)` give back the reference's own types, so
//! `proof.verify(&commitment, &r, &gens, &mut transcript)` of the reference runs unchanged.
//!
//! `Fr`, `EdwardsAffine`, `EdwardsProjective` cross the boundary as raw pointers to their in-memory layout
//! (4 / 8 / 16 little-endian u64 limbs in Montgomery form) — no conversion, no copy on the Rust side.
//! NOT BUILT in the repository's own image (no cargo there); see integration/rust/README.md.
#![allow(non_snake_case)]
pub mod msm;
pub mod sys;

use std::ffi::{CStr, CString};
use std::ptr;

use ark_curve25519::{EdwardsAffine, Fr};

/// `SubtableStrategy` impls of the reference as runtime values (src/subtables/{and,or,xor,lt,range_check}.rs)
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub enum Strategy {
    And,
    Or,
    Xor,
    Lt,
    RangeCheck { log_r: i32 },
}
impl Strategy {
    fn kind(self) -> i32 {
        match self {
            Strategy::And => 0,
            Strategy::Or => 1,
            Strategy::Xor => 2,
            Strategy::Lt => 3,
            Strategy::RangeCheck { .. } => 4,
        }
    }
    fn log_r(self) -> i32 {
        if let Strategy::RangeCheck { log_r } = self {
            log_r
        } else {
            0
        }
    }
    /// `SubtableStrategy::NUM_MEMORIES` (src/subtables/mod.rs:31-93)
    pub fn num_memories(self, c: usize) -> usize {
        if self == Strategy::Lt {
            2 * c
        } else {
            c
        }
    }
}

/// Error = the reference's panic / `Err` condition (code > 0, include/lasso_b200.h) or a CUDA / internal error (< 0).
#[derive(Debug)]
pub struct Error {
    pub code: i32,
    pub message: String,
}
fn check(rc: i32) -> Result<(), Error> {
    if rc == 0 {
        return Ok(());
    }
    let message = unsafe { CStr::from_ptr(sys::lasso_last_error()) }.to_string_lossy().into_owned();
    Err(Error { code: rc, message })
}

/// One per GPU: device, stream, memory pool, scratch.  There is no CPU fallback: creation fails without a device.
pub struct Context {
    pub(crate) raw: *mut sys::lasso_ctx,
}
impl Context {
    pub fn new(device_id: i32) -> Result<Self, Error> {
        let mut raw = ptr::null_mut();
        check(unsafe { sys::lasso_ctx_create(&mut raw, device_id) })?;
        Ok(Context { raw })
    }
    /// One proof sharded over `world` ranks of one node (one process per GPU); `id` = the 128 bytes rank 0 got from
    /// `Context::unique_id`, delivered out of band.  Afterwards densify / commit / prove are collective.
    pub fn init_comm(&mut self, id: &[u8; 128], rank: i32, world: i32) -> Result<(), Error> {
        check(unsafe { sys::lasso_ctx_init_comm(self.raw, id.as_ptr(), rank, world) })
    }
    pub fn unique_id() -> Result<[u8; 128], Error> {
        let mut id = [0u8; 128];
        check(unsafe { sys::lasso_comm_unique_id(id.as_mut_ptr()) })?;
        Ok(id)
    }
}
impl Drop for Context {
    fn drop(&mut self) {
        unsafe { sys::lasso_ctx_destroy(self.raw) }
    }
}

/// src/lasso/surge.rs:25-58.  `stream` = `gens_n.G ‖ gens_1.G[0] ‖ h` of the widest `PolyCommitmentGens`,
/// normalised to affine (the three of them are prefixes of one generator stream, commitments.rs:22-44).
pub struct SparsePolyCommitmentGens<'c> {
    ctx: &'c Context,
    raw: *mut sys::lasso_gens,
}
impl<'c> SparsePolyCommitmentGens<'c> {
    pub fn points_needed(c: usize, s: usize, num_memories: usize, log_m: usize) -> usize {
        unsafe { sys::lasso_gens_points_needed(c, s, num_memories, log_m) }
    }
    /// `MultiCommitGens::new(n + 1, label)`'s sampling (Shake256 -> ChaCha20Rng -> G::rand), `count` affine points
    pub fn sample(label: &[u8], count: usize) -> Result<Vec<EdwardsAffine>, Error> {
        let label = CString::new(label).expect("label without NUL");
        let mut out = vec![EdwardsAffine::default(); count];
        check(unsafe { sys::lasso_sample_generators(label.as_ptr(), count, out.as_mut_ptr() as *mut u64) })?;
        Ok(out)
    }
    pub fn new(ctx: &'c Context, stream: &[EdwardsAffine], c: usize, s: usize, num_memories: usize, log_m: usize) -> Result<Self, Error> {
        let mut raw = ptr::null_mut();
        check(unsafe {
            sys::lasso_gens_create(ctx.raw, stream.as_ptr() as *const u64, stream.len(), c, s, num_memories, log_m, &mut raw)
        })?;
        Ok(SparsePolyCommitmentGens { ctx, raw })
    }
}
impl Drop for SparsePolyCommitmentGens<'_> {
    fn drop(&mut self) {
        let _ = self.ctx;
        unsafe { sys::lasso_gens_destroy(self.raw) }
    }
}

/// src/lasso/densified.rs:8-96, device resident
pub struct DensifiedRepresentation<'c, const C: usize> {
    ctx: &'c Context,
    raw: *mut sys::lasso_dense,
    pub log_m: usize,
}
impl<'c, const C: usize> DensifiedRepresentation<'c, C> {
    /// densified.rs:22 — `indices` is the reference's `&Vec<[usize; C]>` (usize = u64 on the targets CUDA supports)
    pub fn from_lookup_indices(ctx: &'c Context, indices: &Vec<[usize; C]>, log_m: usize) -> Result<Self, Error> {
        const _: () = assert!(std::mem::size_of::<usize>() == 8);
        let mut raw = ptr::null_mut();
        check(unsafe { sys::lasso_densify(ctx.raw, indices.as_ptr() as *const u64, indices.len(), C, log_m, &mut raw) })?;
        Ok(DensifiedRepresentation { ctx, raw, log_m })
    }
    pub fn s(&self) -> usize {
        unsafe { sys::lasso_dense_s(self.raw) }
    }
    /// densified.rs:78 -> the bytes of `SparsePolynomialCommitment` (surge.rs:61-68), ark-serialize compressed
    pub fn commit(&self, gens: &SparsePolyCommitmentGens) -> Result<Vec<u8>, Error> {
        let mut out = vec![0u8; 1 << 22];
        let mut len = 0usize;
        check(unsafe { sys::lasso_commit(self.ctx.raw, self.raw, gens.raw, out.as_mut_ptr(), out.len(), &mut len) })?;
        out.truncate(len);
        Ok(out)
    }
}
impl<const C: usize> Drop for DensifiedRepresentation<'_, C> {
    fn drop(&mut self) {
        unsafe { sys::lasso_dense_destroy(self.raw) }
    }
}

/// src/lasso/surge.rs:92-211: the proof as ark-serialize (compressed) bytes + every Fiat-Shamir challenge in order
pub struct SparsePolynomialEvaluationProof {
    pub bytes: Vec<u8>,
    pub challenges: Vec<Fr>,
}
impl SparsePolynomialEvaluationProof {
    /// surge.rs:119.  `transcript_label` = the label of `Transcript::new` (b"example" in bench.rs:59), `tape_label`
    /// that of `RandomTape::new` (b"proof"), `tape_seed` the scalar `RandomTape::new` draws (utils/random.rs:15-30).
    pub fn prove<const C: usize>(
        ctx: &Context,
        strategy: Strategy,
        dense: &mut DensifiedRepresentation<C>,
        r: &Vec<Fr>,
        gens: &SparsePolyCommitmentGens,
        transcript_label: &[u8],
        tape_label: &[u8],
        tape_seed: &Fr,
    ) -> Result<Self, Error> {
        let tl = CString::new(transcript_label).expect("label without NUL");
        let pl = CString::new(tape_label).expect("label without NUL");
        let mut bytes = vec![0u8; 1 << 22];
        let mut challenges = vec![Fr::from(0u64); 1 << 14];
        let (mut len, mut nch) = (0usize, 0usize);
        check(unsafe {
            sys::lasso_prove(
                ctx.raw,
                strategy.kind(),
                strategy.log_r(),
                dense.raw,
                r.as_ptr() as *const u64,
                r.len(),
                gens.raw,
                tl.as_ptr(),
                pl.as_ptr(),
                tape_seed as *const Fr as *const u64,
                bytes.as_mut_ptr(),
                bytes.len(),
                &mut len,
                challenges.as_mut_ptr() as *mut u64,
                challenges.len(),
                &mut nch,
            )
        })?;
        bytes.truncate(len);
        challenges.truncate(nch);
        Ok(SparsePolynomialEvaluationProof { bytes, challenges })
    }
}
