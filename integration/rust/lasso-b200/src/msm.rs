//! Drop-in for `crate::msm::VariableBaseMSM` of the reference (src/msm/mod.rs:14-52) on curve25519, at the seam the
//! crate already has: every MSM call site picks its implementation with
//! `#[cfg(feature = "ark-msm")] use ark_ec::VariableBaseMSM; #[cfg(not(..))] use crate::msm::VariableBaseMSM;`
//! (src/poly/commitments.rs:8-12, src/poly/dense_mlpoly.rs:18-22, src/subprotocols/sumcheck.rs:16-20,
//! src/subprotocols/bullet.rs:17-21).  A `b200` feature adds a third arm: `use lasso_b200::msm::VariableBaseMSM`.
use ark_curve25519::{EdwardsAffine, EdwardsProjective, Fr};

use crate::{sys, Context, Error};

pub trait VariableBaseMSM: Sized {
    /// `msm/mod.rs:36-40`: `Err(min_len)` when the lengths differ, like the reference
    fn msm(ctx: &Context, bases: &[EdwardsAffine], scalars: &[Fr]) -> Result<Self, usize>;
}
impl VariableBaseMSM for EdwardsProjective {
    fn msm(ctx: &Context, bases: &[EdwardsAffine], scalars: &[Fr]) -> Result<Self, usize> {
        if bases.len() != scalars.len() {
            return Err(bases.len().min(scalars.len()));
        }
        let mut out = EdwardsProjective::default();
        let rc = unsafe {
            sys::lasso_msm(ctx.raw, bases.as_ptr() as *const u64, scalars.as_ptr() as *const u64, scalars.len(),
                           &mut out as *mut EdwardsProjective as *mut u64)
        };
        assert_eq!(rc, 0, "lasso_msm failed");
        Ok(out)
    }
}

/// `DensePolynomial::commit_inner` (src/poly/dense_mlpoly.rs:109-128): Z viewed as L x R, one point per row
pub fn commit_rows(ctx: &Context, gens: &[EdwardsAffine], z: &[Fr], l_size: usize, r_size: usize) -> Result<Vec<EdwardsProjective>, Error> {
    assert!(gens.len() >= r_size && z.len() == l_size * r_size);
    let mut out = vec![EdwardsProjective::default(); l_size];
    let rc = unsafe {
        sys::lasso_commit_rows(ctx.raw, gens.as_ptr() as *const u64, z.as_ptr() as *const u64, l_size, r_size,
                               out.as_mut_ptr() as *mut u64)
    };
    if rc != 0 {
        return Err(Error { code: rc, message: "lasso_commit_rows failed".into() });
    }
    Ok(out)
}
