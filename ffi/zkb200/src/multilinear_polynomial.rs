//! `multilinear_polynomial::multilinear_polynomial_evaluation::MultilinearPoly` and
//! `multilinear_polynomial::composed_polynomial::{ProductPoly, SumPoly}` with the reference's public
//! fields and methods.  `evaluation: Vec<F>` stays a host vector (callers read it directly,
//! sum_check_protocol.rs:27, gkr_protocol.rs:259); the device copy is made inside each call, so the
//! coarse entry points (`sum_check_protocol::{prove, gkr_prove}`) are where residency pays off.
use crate::field::{limbs, limbs_mut, zeroed, Zkb200Field};
use crate::{check, ctx};
use zkb200_sys as sys;

#[derive(Debug, Clone, PartialEq)]
pub struct MultilinearPoly<F: Zkb200Field> {
    pub evaluation: Vec<F>,
    pub num_of_vars: u32,
}

impl<F: Zkb200Field> MultilinearPoly<F> {
    pub fn new(evaluations: Vec<F>) -> Self {
        if !evaluations.len().is_power_of_two() {
            panic!("Invalid evaluations");
        }
        let num_of_vars = evaluations.len().ilog2();
        Self { evaluation: evaluations, num_of_vars }
    }
    pub(crate) fn upload(&self, c: &crate::Ctx) -> sys::zkb_mle {
        let mut h = 0;
        check(c.0, unsafe { sys::zkb_mle_upload(c.0, limbs(&self.evaluation), self.evaluation.len() as u64, &mut h) });
        h
    }
    fn download(c: &crate::Ctx, h: sys::zkb_mle) -> Self {
        let mut nv = 0u32;
        check(c.0, unsafe { sys::zkb_mle_num_vars(c.0, h, &mut nv) });
        let mut out = zeroed::<F>(1usize << nv);
        check(c.0, unsafe { sys::zkb_mle_download(c.0, h, limbs_mut(&mut out)) });
        unsafe { sys::zkb_mle_free(c.0, h) };
        Self::new(out)
    }
    pub fn partial_evaluate(&self, bit: u32, value: &F) -> Self {
        let c = ctx(F::FIELD_ID, sys::ZKB_MODE_COMPAT);
        let (h, mut o) = (self.upload(&c), 0);
        check(c.0, unsafe { sys::zkb_mle_partial_evaluate(c.0, h, bit, limbs(core::slice::from_ref(value)), &mut o) });
        unsafe { sys::zkb_mle_free(c.0, h) };
        Self::download(&c, o)
    }
    pub fn multi_partial_evaluate(&self, values: &[F]) -> Self {
        let c = ctx(F::FIELD_ID, sys::ZKB_MODE_COMPAT);
        let (h, mut o) = (self.upload(&c), 0);
        check(c.0, unsafe { sys::zkb_mle_multi_partial_evaluate(c.0, h, limbs(values), values.len() as u32, &mut o) });
        unsafe { sys::zkb_mle_free(c.0, h) };
        Self::download(&c, o)
    }
    pub fn evaluate(&self, values: Vec<F>) -> F {
        let c = ctx(F::FIELD_ID, sys::ZKB_MODE_COMPAT);
        let h = self.upload(&c);
        let mut out = [F::zero()];
        check(c.0, unsafe { sys::zkb_mle_evaluate(c.0, h, limbs(&values), values.len() as u32, limbs_mut(&mut out)) });
        unsafe { sys::zkb_mle_free(c.0, h) };
        out[0]
    }
    pub fn scale(&self, value: F) -> Self {
        let c = ctx(F::FIELD_ID, sys::ZKB_MODE_COMPAT);
        let (h, mut o) = (self.upload(&c), 0);
        check(c.0, unsafe { sys::zkb_mle_scale(c.0, h, limbs(&[value]), &mut o) });
        unsafe { sys::zkb_mle_free(c.0, h) };
        Self::download(&c, o)
    }
}

#[derive(Clone, Debug, PartialEq)]
pub struct ProductPoly<F: Zkb200Field> {
    pub evaluation: Vec<MultilinearPoly<F>>,
}
impl<F: Zkb200Field> ProductPoly<F> {
    pub fn new(evaluations: Vec<Vec<F>>) -> Self {
        let length_1 = evaluations[0].len();
        if evaluations.iter().any(|e| e.len() != length_1) {
            panic!("all evaluations must have same length");
        }
        Self { evaluation: evaluations.into_iter().map(MultilinearPoly::new).collect() }
    }
    pub fn get_degree(&self) -> usize {
        self.evaluation.len()
    }
}

#[derive(Clone, Debug, PartialEq)]
pub struct SumPoly<F: Zkb200Field> {
    pub polys: Vec<ProductPoly<F>>,
}
impl<F: Zkb200Field> SumPoly<F> {
    pub fn new(polys: Vec<ProductPoly<F>>) -> Self {
        let degree_1 = polys[0].get_degree();
        if polys.iter().any(|p| p.get_degree() != degree_1) {
            panic!("all product polys must have same degree");
        }
        Self { polys }
    }
    pub fn get_degree(&self) -> usize {
        self.polys[0].get_degree()
    }
    /// Upload every table and build the device-side composed polynomial.
    pub(crate) fn upload(&self, c: &crate::Ctx) -> (sys::zkb_sp, Vec<sys::zkb_mle>) {
        let tabs: Vec<sys::zkb_mle> = self.polys.iter().flat_map(|p| p.evaluation.iter().map(|m| m.upload(c))).collect();
        let mut sp = 0;
        check(c.0, unsafe { sys::zkb_sumpoly_create(c.0, tabs.as_ptr(), self.polys.len() as u32, self.get_degree() as u32, &mut sp) });
        (sp, tabs)
    }
}
