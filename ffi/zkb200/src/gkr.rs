//! `gkr::gkr_circuit::{Gate, Layer, Circuit}` and `gkr::gkr_protocol::{prove, verify, GkrProof}` with the
//! reference's shapes (gkr_circuit.rs:4-143, gkr_protocol.rs:23-227) on top of `zkb_circuit_*` / `zkb_gkr_*`.
//!
//! Differences a caller sees, both forced by scope (DESIGN.md sections 3 and 9):
//! * the KZG commitment / opening of the input layer (gkr_protocol.rs:92-118,157-183) is not part of this
//!   engine: `GkrProof` carries the two input-MLE openings instead of `input_proof`, and `verify` takes the
//!   inputs and recomputes those openings on the device;
//! * `prove` runs the linear-time two-phase form of the per-layer sumcheck; its round polynomials are
//!   bit-identical to the reference's dense construction.
//!
//! NOT COMPILED in the build container (no Rust toolchain there); see lib.rs.
use crate::field::{limbs, limbs_mut, zeroed, Zkb200Field};
use crate::{check, ctx};
use zkb200_sys as sys;

#[derive(Debug, Clone, Copy, PartialEq, Eq)]
pub enum Operation {
    Add,
    Mul,
}
impl Operation {
    pub fn apply<F: Zkb200Field>(self, a: F, b: F) -> F {
        match self {
            Operation::Add => a + b,
            Operation::Mul => a * b,
        }
    }
    fn code(self) -> u8 {
        match self {
            Operation::Add => sys::ZKB_OP_ADD as u8,
            Operation::Mul => sys::ZKB_OP_MUL as u8,
        }
    }
}

pub mod gkr_circuit {
    use super::*;

    #[derive(Debug, Clone)]
    pub struct Gate<F: Zkb200Field> {
        pub l_input: F,
        pub r_input: F,
        pub output: F,
        pub op: Operation,
    }
    impl<F: Zkb200Field> Gate<F> {
        pub fn new(l_input: F, r_input: F, op: Operation) -> Self {
            Self { l_input, r_input, output: op.apply(l_input, r_input), op }
        }
    }

    #[derive(Debug, Clone)]
    pub struct Layer<F: Zkb200Field> {
        pub gates: Vec<Gate<F>>,
    }
    impl<F: Zkb200Field> Layer<F> {
        pub fn new(gates: Vec<Gate<F>>) -> Self {
            Self { gates }
        }
        pub fn get_layer_poly(&self) -> Vec<F> {
            self.gates.iter().map(|g| g.output).collect()
        }
        /// gkr_circuit.rs:39-52: the dense 0/1 indicator over a||b||c (2^(3g+2) entries; small layers only), made on
        /// the device by `zkb_layer_add_mul_i`.
        pub fn get_add_mul_i(&self, op: Operation) -> crate::multilinear_polynomial::MultilinearPoly<F> {
            let c = ctx(F::FIELD_ID, sys::ZKB_MODE_COMPAT);
            let ops: Vec<u8> = self.gates.iter().map(|g| g.op.code()).collect();
            let mut h: sys::zkb_mle = 0;
            check(c.0, unsafe { sys::zkb_layer_add_mul_i(c.0, ops.as_ptr(), ops.len() as u32, op.code() as i32, &mut h) });
            let mut nv = 0u32;
            check(c.0, unsafe { sys::zkb_mle_num_vars(c.0, h, &mut nv) });
            let mut out = zeroed::<F>(1usize << nv);
            check(c.0, unsafe { sys::zkb_mle_download(c.0, h, limbs_mut(&mut out)) });
            unsafe { sys::zkb_mle_free(c.0, h) };
            crate::multilinear_polynomial::MultilinearPoly::new(out)
        }
    }

    #[derive(Debug, Clone)]
    pub struct Circuit<F: Zkb200Field> {
        pub layers: Vec<Layer<F>>,
    }
    impl<F: Zkb200Field> Circuit<F> {
        /// gkr_circuit.rs:113-125: layers listed input side first.
        pub fn new(structure: Vec<Vec<Operation>>) -> Self {
            let layers = structure
                .into_iter()
                .map(|ops| Layer::new(ops.into_iter().map(|op| Gate::new(F::zero(), F::zero(), op)).collect()))
                .collect();
            Self { layers }
        }
        pub(crate) fn shape(&self) -> (Vec<u32>, Vec<u8>) {
            let gates = self.layers.iter().map(|l| l.gates.len() as u32).collect();
            let ops = self.layers.iter().flat_map(|l| l.gates.iter().map(|g| g.op.code())).collect();
            (gates, ops)
        }
        pub(crate) fn device(&self, c: &crate::Ctx) -> sys::zkb_circ {
            let (gates, ops) = self.shape();
            let mut h: sys::zkb_circ = 0;
            check(c.0, unsafe { sys::zkb_circuit_create(c.0, gates.len() as u32, gates.as_ptr(), ops.as_ptr(), &mut h) });
            h
        }
        /// gkr_circuit.rs:127-143: evaluates on the device and, like the reference, records every gate's inputs
        /// and output in `self`.
        pub fn evaluate(&mut self, inputs: &[F]) -> Vec<Vec<F>> {
            let c = ctx(F::FIELD_ID, sys::ZKB_MODE_COMPAT);
            let h = self.device(&c);
            let total: usize = self.layers.iter().map(|l| l.gates.len()).sum();
            let mut flat = zeroed::<F>(total);
            check(c.0, unsafe { sys::zkb_circuit_evaluate(c.0, h, limbs(inputs), inputs.len() as u64, limbs_mut(&mut flat)) });
            unsafe { sys::zkb_circuit_free(c.0, h) };
            let mut result = Vec::with_capacity(self.layers.len());
            let mut below: Vec<F> = inputs.to_vec();
            let mut off = 0;
            for layer in &mut self.layers {
                let out = flat[off..off + layer.gates.len()].to_vec();
                for (i, gate) in layer.gates.iter_mut().enumerate() {
                    gate.l_input = below[2 * i];
                    gate.r_input = below[2 * i + 1];
                    gate.output = out[i];
                }
                off += layer.gates.len();
                below = out.clone();
                result.push(out);
            }
            result
        }
    }
}

pub mod gkr_protocol {
    use super::gkr_circuit::Circuit;
    use super::*;

    #[derive(Debug, Clone)]
    pub struct GkrProof<F: Zkb200Field> {
        pub output_poly: Vec<F>,                  // w_0, padded to two entries (gkr_protocol.rs:34-39)
        pub proof_polynomials: Vec<Vec<Vec<F>>>,  // per layer (output side first), per round: trimmed coefficients
        pub claimed_evaluations: Vec<(F, F)>,
        pub final_openings: (F, F),               // the input MLE at (r_b, r_c): what `input_proof` opens in the reference
    }

    fn rounds_per_layer<F: Zkb200Field>(circuit: &Circuit<F>) -> Vec<usize> {
        circuit.layers.iter().rev().map(|l| 2 * ((2 * l.gates.len()).ilog2() as usize).max(1)).collect()
    }

    /// gkr_protocol.rs:31-126 without the KZG part.
    pub fn prove<F: Zkb200Field>(circuit: &mut Circuit<F>, inputs: &[F]) -> GkrProof<F> {
        let c = ctx(F::FIELD_ID, sys::ZKB_MODE_COMPAT);
        let h = circuit.device(&c);
        let rpl = rounds_per_layer(circuit);
        let (total, n_layers) = (rpl.iter().sum::<usize>(), circuit.layers.len());
        let mut w0 = zeroed::<F>(2);
        let mut coeffs = zeroed::<F>(total * 3);
        let mut lens = vec![0i32; total];
        let mut claimed = zeroed::<F>(2 * (n_layers - 1).max(1));
        let mut fin = zeroed::<F>(2);
        let mut n_rounds = 0u32;
        check(c.0, unsafe {
            sys::zkb_gkr_prove(c.0, h, limbs(inputs), inputs.len() as u64, limbs_mut(&mut w0), limbs_mut(&mut coeffs), lens.as_mut_ptr(),
                               core::ptr::null_mut(), limbs_mut(&mut claimed), limbs_mut(&mut fin), &mut n_rounds)
        });
        unsafe { sys::zkb_circuit_free(c.0, h) };
        assert_eq!(n_rounds as usize, total);
        let mut proof_polynomials = Vec::with_capacity(n_layers);
        let mut k = 0;
        for n in rpl {
            proof_polynomials.push((k..k + n).map(|r| coeffs[3 * r..3 * r + lens[r] as usize].to_vec()).collect());
            k += n;
        }
        GkrProof {
            output_poly: w0,
            proof_polynomials,
            claimed_evaluations: (0..n_layers - 1).map(|i| (claimed[2 * i], claimed[2 * i + 1])).collect(),
            final_openings: (fin[0], fin[1]),
        }
    }

    /// gkr_protocol.rs:128-227 with the input opening replaced by an evaluation of the input MLE on the device.
    pub fn verify<F: Zkb200Field>(proof: GkrProof<F>, circuit: Circuit<F>, inputs: &[F]) -> bool {
        let c = ctx(F::FIELD_ID, sys::ZKB_MODE_COMPAT);
        let rpl = rounds_per_layer(&circuit);
        let (total, n_layers) = (rpl.iter().sum::<usize>(), circuit.layers.len());
        if proof.proof_polynomials.len() != n_layers
            || proof.proof_polynomials.iter().map(|l| l.len()).ne(rpl.iter().copied())
            || proof.claimed_evaluations.len() != n_layers - 1
            || proof.output_poly.len() != 2
        {
            return false;
        }
        let mut coeffs = zeroed::<F>(total * 3);
        let mut lens = vec![0i32; total];
        for (r, q) in proof.proof_polynomials.iter().flatten().enumerate() {
            if q.len() > 3 {
                return false;
            }
            lens[r] = q.len() as i32;
            coeffs[3 * r..3 * r + q.len()].copy_from_slice(q);
        }
        let claimed: Vec<F> = proof.claimed_evaluations.iter().flat_map(|(a, b)| [*a, *b]).chain([F::zero(), F::zero()]).collect();
        let fin = [proof.final_openings.0, proof.final_openings.1];
        let h = circuit.device(&c);
        let mut ok = 0i32;
        check(c.0, unsafe {
            sys::zkb_gkr_verify(c.0, h, limbs(inputs), inputs.len() as u64, limbs(&proof.output_poly), limbs(&coeffs), lens.as_ptr(),
                                limbs(&claimed), limbs(&fin), &mut ok)
        });
        unsafe { sys::zkb_circuit_free(c.0, h) };
        ok != 0
    }
}
