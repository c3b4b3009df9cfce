//! `fiat_shamir::fiat_shamir_transcript::{Transcript, fq_vec_to_bytes}` (fiat_shamir_transcript.rs:5-37) on the
//! library's host transcript (`zkb_transcript_*`: the same Keccak-256 running hash, re-seeded with its digest after
//! every challenge).  The handle is what `sum_check_protocol::{gkr_prove, gkr_verify}` pass to the C ABI, so the
//! caller's `&mut Transcript<F>` continues across a proof exactly as in the reference.
//!
//! NOT COMPILED in the build container (no Rust toolchain there); see lib.rs.
use crate::field::{limbs_mut, Zkb200Field};
use std::marker::PhantomData;
use zkb200_sys as sys;

pub struct Transcript<F: Zkb200Field> {
    _field: PhantomData<F>,
    pub(crate) raw: *mut sys::zkb_transcript,
}

impl<F: Zkb200Field> Transcript<F> {
    pub fn new() -> Self {
        let mut raw = std::ptr::null_mut();
        crate::check(std::ptr::null_mut(), unsafe { sys::zkb_transcript_new(F::FIELD_ID, &mut raw) });
        Self { _field: PhantomData, raw }
    }
    /// fiat_shamir_transcript.rs:19-21
    pub fn append(&mut self, preimage: &[u8]) {
        crate::check(std::ptr::null_mut(), unsafe { sys::zkb_transcript_append(self.raw, preimage.as_ptr(), preimage.len()) });
    }
    /// fiat_shamir_transcript.rs:23-29
    pub fn get_random_challenge(&mut self) -> F {
        let mut out = [F::zero()];
        crate::check(std::ptr::null_mut(), unsafe { sys::zkb_transcript_challenge(self.raw, limbs_mut(&mut out)) });
        out[0]
    }
}
impl<F: Zkb200Field> Default for Transcript<F> {
    fn default() -> Self {
        Self::new()
    }
}
impl<F: Zkb200Field> Drop for Transcript<F> {
    fn drop(&mut self) {
        unsafe { sys::zkb_transcript_free(self.raw) };
    }
}

/// fiat_shamir_transcript.rs:32-37: 32-byte little-endian canonical integers.
pub fn fq_vec_to_bytes<F: Zkb200Field>(values: &[F]) -> Vec<u8> {
    use ark_ff::{BigInteger, PrimeField};
    values.iter().flat_map(|x| PrimeField::into_bigint(*x).to_bytes_le()).collect()
}
