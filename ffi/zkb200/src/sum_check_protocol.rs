//! `sum_check::sum_check_protocol::{prove, verify, gkr_prove, gkr_verify}` with the reference's signatures
//! (sum_check_protocol.rs:8-150).  One upload, all rounds on the device, (d+1) elements back per round.
use crate::field::{limbs, limbs_mut, zeroed, Zkb200Field};
use crate::fiat_shamir::Transcript;
use crate::multilinear_polynomial::{MultilinearPoly, SumPoly};
use crate::univariate_polynomial::UnivariatePoly;
use crate::{check, ctx};
use zkb200_sys as sys;

#[derive(Debug, Clone)]
pub struct Proof<F: Zkb200Field> {
    pub proof_polynomials: Vec<Vec<F>>,
    pub claimed_sum: F,
}
pub struct GkrProof<F: Zkb200Field> {
    pub proof_polynomials: Vec<UnivariatePoly<F>>, // trimmed coefficient vectors (univariate_polynomial_dense.rs:71)
    pub claimed_sum: F,
    pub random_challenges: Vec<F>,
}
pub struct GkrVerify<F: Zkb200Field> {
    pub verified: bool,
    pub final_claimed_sum: F,
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
    // The C side reads exactly two elements per round message.  The reference builds MultilinearPoly::new(poly) (panics
    // unless the length is a power of two, :63) and reads evaluation[0], evaluation[1] (:66,73-74: a length-1 message
    // panics on the index), so: same panics here, and `flat` is built from exactly m[0], m[1].
    let mut flat: Vec<F> = Vec::with_capacity(2 * proof.proof_polynomials.len());
    for m in &proof.proof_polynomials {
        let e = MultilinearPoly::new(m.clone()).evaluation; // "Invalid evaluations" on a bad length
        assert!(e.len() >= 2, "index out of bounds: the len is {} but the index is 1", e.len());
        flat.push(e[0]);
        flat.push(e[1]);
    }
    let mut ok = 0i32;
    check(c.0, unsafe {
        sys::zkb_sumcheck_verify(c.0, h, 1, limbs(&[proof.claimed_sum]), limbs(&flat), proof.proof_polynomials.len() as u32, &mut ok)
    });
    unsafe { sys::zkb_mle_free(c.0, h) };
    ok != 0
}

/// sum_check_protocol.rs:86-115.  `claimed_sum` is echoed, not used (:87,112).  One upload, all rounds on the device;
/// the caller's transcript absorbs every round message and ends in the same state as the reference's.
pub fn gkr_prove<F: Zkb200Field>(claimed_sum: F, composed_polynomial: &SumPoly<F>, transcript: &mut Transcript<F>) -> GkrProof<F> {
    let c = ctx(F::FIELD_ID, sys::ZKB_MODE_COMPAT);
    let (sp, tabs) = composed_polynomial.upload(&c);
    let n = composed_polynomial.polys[0].evaluation[0].num_of_vars as usize;
    let d = composed_polynomial.get_degree();
    let mut coeffs = zeroed::<F>(n.max(1) * (d + 1));
    let mut lens = vec![0i32; n.max(1)];
    let mut chals = zeroed::<F>(n.max(1));
    check(c.0, unsafe {
        sys::zkb_gkr_sumcheck_prove(c.0, transcript.raw, limbs(&[claimed_sum]), sp, limbs_mut(&mut coeffs), lens.as_mut_ptr(), limbs_mut(&mut chals), core::ptr::null_mut())
    });
    unsafe { sys::zkb_sumpoly_free(c.0, sp) };
    for t in tabs {
        unsafe { sys::zkb_mle_free(c.0, t) };
    }
    GkrProof {
        proof_polynomials: (0..n).map(|k| UnivariatePoly::new(coeffs[k * (d + 1)..k * (d + 1) + lens[k] as usize].to_vec())).collect(),
        claimed_sum,
        random_challenges: chals[..n].to_vec(),
    }
}

/// sum_check_protocol.rs:117-150 (host arithmetic of the library; no device needed).  On a failed round check the
/// reference returns `verified: false, final_claimed_sum: 0, random_challenges: [0]` (:129-133): so does this.
pub fn gkr_verify<F: Zkb200Field>(round_polys: Vec<UnivariatePoly<F>>, claimed_sum: F, transcript: &mut Transcript<F>) -> GkrVerify<F> {
    let n = round_polys.len();
    let slots = round_polys.iter().map(|q| q.coefficient.len()).max().unwrap_or(0).max(1);
    let mut coeffs = zeroed::<F>(n.max(1) * slots);
    let mut lens = vec![0i32; n.max(1)];
    for (k, q) in round_polys.iter().enumerate() {
        lens[k] = q.coefficient.len() as i32;
        coeffs[k * slots..k * slots + q.coefficient.len()].copy_from_slice(&q.coefficient);
    }
    let (mut ok, mut fin, mut chals) = (0i32, [F::zero()], zeroed::<F>(n.max(1)));
    check(core::ptr::null_mut(), unsafe {
        sys::zkb_gkr_sumcheck_verify(transcript.raw, n as u32, slots as u32, limbs(&coeffs), lens.as_ptr(), limbs(&[claimed_sum]), &mut ok,
                                     limbs_mut(&mut fin), limbs_mut(&mut chals))
    });
    if ok == 0 {
        return GkrVerify { verified: false, final_claimed_sum: F::zero(), random_challenges: vec![F::zero()] };
    }
    GkrVerify { verified: true, final_claimed_sum: fin[0], random_challenges: chals[..n].to_vec() }
}
