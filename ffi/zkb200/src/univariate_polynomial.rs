//! `univariate_polynomial::univariate_polynomial_dense::UnivariatePoly` (univariate_polynomial_dense.rs:4-74) for the
//! round messages of the composed sumcheck: ascending coefficients, Lagrange interpolation that ends with `trim()`
//! (trailing zeros removed, possibly down to an empty vector).  Host arithmetic of the library (`zkb_uni_*`).
//!
//! NOT COMPILED in the build container (no Rust toolchain there); see lib.rs.
use crate::field::{limbs, limbs_mut, zeroed, Zkb200Field};
use zkb200_sys as sys;

#[derive(Debug, Clone)]
pub struct UnivariatePoly<F: Zkb200Field> {
    pub coefficient: Vec<F>,
}

impl<F: Zkb200Field> UnivariatePoly<F> {
    pub fn new(coeff: Vec<F>) -> Self {
        UnivariatePoly { coefficient: coeff }
    }
    /// univariate_polynomial_dense.rs:20-26
    pub fn evaluate(&self, x: F) -> F {
        let mut out = [F::zero()];
        crate::check(std::ptr::null_mut(), unsafe {
            sys::zkb_uni_evaluate(F::FIELD_ID, limbs(&self.coefficient), self.coefficient.len() as u32, limbs(&[x]), limbs_mut(&mut out))
        });
        out[0]
    }
    /// univariate_polynomial_dense.rs:28-32
    pub fn degree(&mut self) -> usize {
        while self.coefficient.last() == Some(&F::zero()) {
            self.coefficient.pop();
        }
        if self.coefficient.is_empty() { 0 } else { self.coefficient.len() - 1 }
    }
    /// univariate_polynomial_dense.rs:48-74
    pub fn interpolate(points: Vec<(F, F)>) -> UnivariatePoly<F> {
        let xs: Vec<F> = points.iter().map(|p| p.0).collect();
        let ys: Vec<F> = points.iter().map(|p| p.1).collect();
        let mut coeffs = zeroed::<F>(points.len().max(1));
        let mut len = 0u32;
        crate::check(std::ptr::null_mut(), unsafe {
            sys::zkb_uni_interpolate(F::FIELD_ID, limbs(&xs), limbs(&ys), points.len() as u32, limbs_mut(&mut coeffs), &mut len)
        });
        coeffs.truncate(len as usize);
        UnivariatePoly { coefficient: coeffs }
    }
}
