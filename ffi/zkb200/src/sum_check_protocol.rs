//! `sum_check::sum_check_protocol::{prove, verify, gkr_prove}` with the reference's signatures
//! (sum_check_protocol.rs:8-115).  One upload, all rounds on the device, (d+1) elements back per round.
use crate::field::{limbs, limbs_mut, zeroed, Zkb200Field};
use crate::multilinear_polynomial::{MultilinearPoly, SumPoly};
use crate::{check, ctx};
use zkb200_sys as sys;

#[derive(Debug, Clone)]
pub struct Proof<F: Zkb200Field> {
    pub proof_polynomials: Vec<Vec<F>>,
    pub claimed_sum: F,
}
pub struct GkrProof<F: Zkb200Field> {
    pub proof_polynomials: Vec<Vec<F>>, // coefficient vectors of the UnivariatePoly round messages, trimmed
    pub claimed_sum: F,
    pub random_challenges: Vec<F>,
}

pub fn prove<F: Zkb200Field>(polynomial: &MultilinearPoly<F>) -> Proof<F> {
    let c = ctx(F::FIELD_ID, sys::ZKB_MODE_COMPAT);
    let h = polynomial.upload(&c);
    let n = polynomial.num_of_vars as usize;
    let mut claimed = [F::zero()];
    let mut msgs = zeroed::<F>(2 * n.max(1));
    // flags = 1: absorb the whole table into the transcript first, as the reference does (:27)
    check(c.0, unsafe { sys::zkb_sumcheck_prove(c.0, h, 1, limbs_mut(&mut claimed), limbs_mut(&mut msgs), core::ptr::null_mut()) });
    unsafe { sys::zkb_mle_free(c.0, h) };
    Proof { proof_polynomials: msgs[..2 * n].chunks(2).map(|m| m.to_vec()).collect(), claimed_sum: claimed[0] }
}

pub fn verify<F: Zkb200Field>(polynomial: &MultilinearPoly<F>, proof: Proof<F>) -> bool {
    let c = ctx(F::FIELD_ID, sys::ZKB_MODE_COMPAT);
    let h = polynomial.upload(&c);
    let flat: Vec<F> = proof.proof_polynomials.iter().flat_map(|m| MultilinearPoly::new(m.clone()).evaluation).collect();
    let mut ok = 0i32;
    check(c.0, unsafe {
        sys::zkb_sumcheck_verify(c.0, h, 1, limbs(&[proof.claimed_sum]), limbs(&flat), proof.proof_polynomials.len() as u32, &mut ok)
    });
    unsafe { sys::zkb_mle_free(c.0, h) };
    ok != 0
}

/// `transcript` is the library's host transcript handle (same Keccak-256 construction as
/// fiat_shamir::Transcript; `zkb_transcript_*`).
pub fn gkr_prove<F: Zkb200Field>(claimed_sum: F, composed_polynomial: &SumPoly<F>, transcript: *mut sys::zkb_transcript) -> GkrProof<F> {
    let c = ctx(F::FIELD_ID, sys::ZKB_MODE_COMPAT);
    let (sp, tabs) = composed_polynomial.upload(&c);
    let n = composed_polynomial.polys[0].evaluation[0].num_of_vars as usize;
    let d = composed_polynomial.get_degree();
    let mut coeffs = zeroed::<F>(n.max(1) * (d + 1));
    let mut lens = vec![0i32; n.max(1)];
    let mut chals = zeroed::<F>(n.max(1));
    check(c.0, unsafe {
        sys::zkb_gkr_sumcheck_prove(c.0, transcript, limbs(&[claimed_sum]), sp, limbs_mut(&mut coeffs), lens.as_mut_ptr(), limbs_mut(&mut chals), core::ptr::null_mut())
    });
    unsafe { sys::zkb_sumpoly_free(c.0, sp) };
    for t in tabs {
        unsafe { sys::zkb_mle_free(c.0, t) };
    }
    GkrProof {
        proof_polynomials: (0..n).map(|k| coeffs[k * (d + 1)..k * (d + 1) + lens[k] as usize].to_vec()).collect(),
        claimed_sum,
        random_challenges: chals[..n].to_vec(),
    }
}
