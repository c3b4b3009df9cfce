//! Which library field an ark-ff type maps to, and the zero-cost limb view.
use ark_ff::PrimeField;
use zkb200_sys as sys;

pub trait Zkb200Field: PrimeField {
    const FIELD_ID: i32;
}
impl Zkb200Field for ark_bn254::Fr {
    const FIELD_ID: i32 = sys::ZKB_FIELD_BN254_FR;
}
impl Zkb200Field for ark_bn254::Fq {
    const FIELD_ID: i32 = sys::ZKB_FIELD_BN254_FQ;
}
impl Zkb200Field for ark_bls12_381::Fr {
    const FIELD_ID: i32 = sys::ZKB_FIELD_BLS12_381_FR;
}

/// `&[F]` as the `*const u64` the C ABI expects (Montgomery limbs, 4 per element).
pub fn limbs<F: Zkb200Field>(v: &[F]) -> *const u64 {
    debug_assert_eq!(core::mem::size_of::<F>(), 32);
    v.as_ptr() as *const u64
}
pub fn limbs_mut<F: Zkb200Field>(v: &mut [F]) -> *mut u64 {
    v.as_mut_ptr() as *mut u64
}
pub fn zeroed<F: Zkb200Field>(n: usize) -> Vec<F> {
    vec![F::zero(); n]
}
