//! Raw bindings of `include/zkb200.h`.  Field elements cross as `[u64; 4]`: the Montgomery limbs of
//! ark-ff 0.5 `Fp<MontBackend<_, 4>>` (`x.0 .0`), so `Vec<F>` is passed as `*const u64` unconverted.
#![allow(non_camel_case_types)]
use core::ffi::{c_char, c_void};

#[repr(C)]
pub struct zkb_ctx {
    _p: [u8; 0],
}
#[repr(C)]
pub struct zkb_transcript {
    _p: [u8; 0],
}
pub type zkb_mle = u64;
pub type zkb_sp = u64;
pub type zkb_circ = u64;
pub type zkb_kzg = u64;
pub type zkb_merkle = u64;

pub const ZKB_OK: i32 = 0;
pub const ZKB_ERR_BAD_ARG: i32 = -1;
pub const ZKB_ERR_NOT_POW2: i32 = -2;
pub const ZKB_ERR_ARITY: i32 = -3;
pub const ZKB_ERR_LENGTH_MISMATCH: i32 = -4;
pub const ZKB_ERR_DEGREE_MISMATCH: i32 = -5;
pub const ZKB_ERR_CUDA: i32 = -6;
pub const ZKB_ERR_NCCL: i32 = -7;
pub const ZKB_ERR_OOM: i32 = -8;
pub const ZKB_ERR_UNSUPPORTED: i32 = -9;
pub const ZKB_ERR_COMPAT_SHAPE: i32 = -10;
pub const ZKB_ERR_CIRCUIT_SHAPE: i32 = -11;

pub const ZKB_FIELD_BN254_FR: i32 = 0;
pub const ZKB_FIELD_BN254_FQ: i32 = 1;
pub const ZKB_FIELD_BLS12_381_FR: i32 = 2;
pub const ZKB_MODE_COMPAT: i32 = 0;
pub const ZKB_MODE_FULL: i32 = 1;
pub const ZKB_OP_ADD: i32 = 0;
pub const ZKB_OP_MUL: i32 = 1;
pub const ZKB_OP_SUB: i32 = 2;

extern "C" {
    pub fn zkb_strerror(status: i32) -> *const c_char;
    pub fn zkb_version() -> *const c_char;
    pub fn zkb_ctx_create(field_id: i32, device: i32, mode: i32, out: *mut *mut zkb_ctx) -> i32;
    pub fn zkb_ctx_destroy(ctx: *mut zkb_ctx) -> i32;
    pub fn zkb_ctx_last_error(ctx: *const zkb_ctx) -> *const c_char;
    pub fn zkb_ctx_stream(ctx: *const zkb_ctx) -> *mut c_void;
    pub fn zkb_ctx_launch_count(ctx: *const zkb_ctx) -> u64;
    pub fn zkb_ctx_sync(ctx: *mut zkb_ctx) -> i32;
    pub fn zkb_ctx_profile(ctx: *mut zkb_ctx, enable: i32) -> i32;
    pub fn zkb_ctx_profile_read(ctx: *mut zkb_ctx, kernel_id: i32, launches: *mut u64, ms: *mut f64, alg_bytes: *mut f64) -> i32;
    pub fn zkb_kernel_name(kernel_id: i32) -> *const c_char;
    pub fn zkb_comm_unique_id(out: *mut u8) -> i32;
    pub fn zkb_ctx_comm_init(ctx: *mut zkb_ctx, rank: i32, world: i32, unique_id: *const u8) -> i32;
    pub fn zkb_ctx_set_gather_threshold(ctx: *mut zkb_ctx, log2_local_entries: u32) -> i32;
    pub fn zkb_ctx_set_tail_threshold(ctx: *mut zkb_ctx, log2_entries: u32) -> i32;
    pub fn zkb_ctx_set_small_threshold(ctx: *mut zkb_ctx, smem_bytes: u32) -> i32;
    pub fn zkb_ctx_set_device_transcript(ctx: *mut zkb_ctx, enable: i32) -> i32;
    pub fn zkb_ctx_device_transcript_stats(ctx: *const zkb_ctx, launches: *mut u64, rounds_checked: *mut u64) -> i32;

    pub fn zkb_mle_upload(ctx: *mut zkb_ctx, aos_mont: *const u64, len: u64, out: *mut zkb_mle) -> i32;
    pub fn zkb_mle_upload_shard(ctx: *mut zkb_ctx, aos_mont_full: *const u64, len_full: u64, out: *mut zkb_mle) -> i32;
    pub fn zkb_mle_generate(ctx: *mut zkb_ctx, seed: u64, table_id: u64, n_vars: u32, out: *mut zkb_mle) -> i32;
    pub fn zkb_mle_download(ctx: *mut zkb_ctx, m: zkb_mle, aos_mont: *mut u64) -> i32;
    pub fn zkb_mle_download_canonical(ctx: *mut zkb_ctx, m: zkb_mle, bytes: *mut u8) -> i32;
    pub fn zkb_mle_clone(ctx: *mut zkb_ctx, m: zkb_mle, out: *mut zkb_mle) -> i32;
    pub fn zkb_mle_free(ctx: *mut zkb_ctx, m: zkb_mle) -> i32;
    pub fn zkb_mle_num_vars(ctx: *mut zkb_ctx, m: zkb_mle, n_vars: *mut u32) -> i32;
    pub fn zkb_mle_partial_evaluate(ctx: *mut zkb_ctx, m: zkb_mle, bit: u32, value: *const u64, out: *mut zkb_mle) -> i32;
    pub fn zkb_mle_multi_partial_evaluate(ctx: *mut zkb_ctx, m: zkb_mle, values: *const u64, k: u32, out: *mut zkb_mle) -> i32;
    pub fn zkb_mle_evaluate(ctx: *mut zkb_ctx, m: zkb_mle, values: *const u64, k: u32, out: *mut u64) -> i32;
    pub fn zkb_mle_sum_halves(ctx: *mut zkb_ctx, m: zkb_mle, out: *mut u64) -> i32;
    pub fn zkb_mle_scale(ctx: *mut zkb_ctx, m: zkb_mle, value: *const u64, out: *mut zkb_mle) -> i32;
    pub fn zkb_mle_binary(ctx: *mut zkb_ctx, a: zkb_mle, b: zkb_mle, op: i32, out: *mut zkb_mle) -> i32;
    pub fn zkb_mle_tensor(ctx: *mut zkb_ctx, a: zkb_mle, b: zkb_mle, op: i32, out: *mut zkb_mle) -> i32;

    pub fn zkb_sumpoly_create(ctx: *mut zkb_ctx, tables: *const zkb_mle, n_products: u32, degree: u32, out: *mut zkb_sp) -> i32;
    pub fn zkb_sumpoly_free(ctx: *mut zkb_ctx, sp: zkb_sp) -> i32;
    pub fn zkb_sumpoly_reset(ctx: *mut zkb_ctx, sp: zkb_sp) -> i32;
    pub fn zkb_sumpoly_evaluate(ctx: *mut zkb_ctx, sp: zkb_sp, values: *const u64, k: u32, out: *mut u64) -> i32;
    pub fn zkb_sc_round_evals(ctx: *mut zkb_ctx, sp: zkb_sp, evals: *mut u64) -> i32;
    pub fn zkb_sc_bind_and_next(ctx: *mut zkb_ctx, sp: zkb_sp, r: *const u64, evals: *mut u64) -> i32;
    pub fn zkb_sc_final_values(ctx: *mut zkb_ctx, sp: zkb_sp, values: *mut u64) -> i32;

    pub fn zkb_transcript_new(field_id: i32, out: *mut *mut zkb_transcript) -> i32;
    pub fn zkb_transcript_free(t: *mut zkb_transcript) -> i32;
    pub fn zkb_transcript_append(t: *mut zkb_transcript, bytes: *const u8, len: usize) -> i32;
    pub fn zkb_transcript_append_elements(t: *mut zkb_transcript, mont: *const u64, n: usize) -> i32;
    pub fn zkb_transcript_challenge(t: *mut zkb_transcript, out_mont: *mut u64) -> i32;
    pub fn zkb_keccak256(bytes: *const u8, len: usize, out: *mut u8) -> i32;
    pub fn zkb_fe_to_mont(field_id: i32, canonical: *const u64, mont: *mut u64, n: usize) -> i32;
    pub fn zkb_fe_from_mont(field_id: i32, mont: *const u64, canonical: *mut u64, n: usize) -> i32;
    pub fn zkb_fe_reduce_wide(field_id: i32, wide: *const u64, out: *mut u64, n: usize) -> i32;
    pub fn zkb_uni_interpolate(field_id: i32, xs: *const u64, ys: *const u64, n: u32, coeffs: *mut u64, len: *mut u32) -> i32;
    pub fn zkb_uni_evaluate(field_id: i32, coeffs: *const u64, len: u32, x: *const u64, out: *mut u64) -> i32;

    pub fn zkb_sumcheck_prove(ctx: *mut zkb_ctx, poly: zkb_mle, flags: u32, claimed_sum: *mut u64, msgs: *mut u64, challenges: *mut u64) -> i32;
    pub fn zkb_sumcheck_verify(ctx: *mut zkb_ctx, poly: zkb_mle, flags: u32, claimed_sum: *const u64, msgs: *const u64, n_msgs: u32, accepted: *mut i32) -> i32;
    pub fn zkb_gkr_sumcheck_prove(ctx: *mut zkb_ctx, t: *mut zkb_transcript, claimed_sum: *const u64, sp: zkb_sp, coeffs: *mut u64, lens: *mut i32, challenges: *mut u64, final_values: *mut u64) -> i32;
    pub fn zkb_gkr_sumcheck_verify(t: *mut zkb_transcript, n_rounds: u32, slots: u32, coeffs: *const u64, lens: *const i32, claimed_sum: *const u64, accepted: *mut i32, final_claim: *mut u64, challenges: *mut u64) -> i32;

    pub fn zkb_circuit_create(ctx: *mut zkb_ctx, n_layers: u32, gates_per_layer: *const u32, ops: *const u8, out: *mut zkb_circ) -> i32;
    pub fn zkb_circuit_free(ctx: *mut zkb_ctx, c: zkb_circ) -> i32;
    pub fn zkb_circuit_evaluate(ctx: *mut zkb_ctx, c: zkb_circ, inputs_mont: *const u64, n_inputs: u64, outputs_mont: *mut u64) -> i32;
    pub fn zkb_layer_add_mul_i(ctx: *mut zkb_ctx, ops: *const u8, n_gates: u32, op: i32, out: *mut zkb_mle) -> i32;
    pub fn zkb_gkr_prove(ctx: *mut zkb_ctx, c: zkb_circ, inputs_mont: *const u64, n_inputs: u64, w0: *mut u64, coeffs: *mut u64, lens: *mut i32, challenges: *mut u64, claimed: *mut u64, final_openings: *mut u64, n_rounds: *mut u32) -> i32;
    pub fn zkb_gkr_verify(ctx: *mut zkb_ctx, c: zkb_circ, inputs_mont: *const u64, n_inputs: u64, w0: *const u64, coeffs: *const u64, lens: *const i32, claimed: *const u64, final_openings: *const u64, accepted: *mut i32) -> i32;
    pub fn zkb_gkr_total_rounds(n_layers: u32, gates_per_layer: *const u32) -> u32;
    // general wiring (extension beyond the reference's fixed (2i, 2i+1) wiring; zkb200.h)
    pub fn zkb_circuit_create_wired(ctx: *mut zkb_ctx, n_layers: u32, gates_per_layer: *const u32, n_inputs: u64, ops: *const u8, in1: *const u32, in2: *const u32, out: *mut zkb_circ) -> i32;
    pub fn zkb_circuit_total_rounds(ctx: *mut zkb_ctx, c: zkb_circ, n_rounds: *mut u32) -> i32;
    pub fn zkb_gkr_prove_wired(ctx: *mut zkb_ctx, c: zkb_circ, inputs_mont: *const u64, n_inputs: u64, w0: *mut u64, n_w0: u64, coeffs: *mut u64, lens: *mut i32, challenges: *mut u64, claimed: *mut u64, final_openings: *mut u64, n_rounds: *mut u32) -> i32;
    pub fn zkb_gkr_verify_wired(ctx: *mut zkb_ctx, c: zkb_circ, inputs_mont: *const u64, n_inputs: u64, w0: *const u64, n_w0: u64, coeffs: *const u64, lens: *const i32, claimed: *const u64, final_openings: *const u64, accepted: *mut i32) -> i32;

    pub fn zkb_proof_encode(field_id: i32, kind: i32, n_rounds: u32, slots: u32, msgs_mont: *const u64, lens: *const i32, claimed_sum: *const u64, out: *mut u8, cap: usize, len: *mut usize) -> i32;
    pub fn zkb_proof_decode(bytes: *const u8, len: usize, field_id: *mut i32, kind: *mut i32, n_rounds: *mut u32, slots: u32, msgs_mont: *mut u64, lens: *mut i32, claimed_sum: *mut u64) -> i32;
    pub fn zkb_kzg_setup(ctx: *mut zkb_ctx, n_vars: u32, taus_mont: *const u64, out: *mut zkb_kzg) -> i32;
    pub fn zkb_kzg_free(ctx: *mut zkb_ctx, k: zkb_kzg) -> i32;
    pub fn zkb_kzg_basis(ctx: *mut zkb_ctx, k: zkb_kzg, level: u32, first: u64, count: u64, out: *mut u8) -> i32;
    pub fn zkb_kzg_commit(ctx: *mut zkb_ctx, k: zkb_kzg, poly: zkb_mle, out: *mut u8) -> i32;
    pub fn zkb_kzg_open(ctx: *mut zkb_ctx, k: zkb_kzg, poly: zkb_mle, opening_values: *const u64, n: u32, out: *mut u64) -> i32;
    pub fn zkb_kzg_get_proof(ctx: *mut zkb_ctx, k: zkb_kzg, poly: zkb_mle, opened_value: *const u64, opening_values: *const u64, n: u32, out: *mut u8) -> i32;
    pub fn zkb_tc_fold_matrices(field_id: i32, r_mont: *const u64, out: *mut u8) -> i32;
    pub fn zkb_ctx_tensor_cores(ctx: *const zkb_ctx, enabled: *mut i32, persistent: *mut i32) -> i32;
    pub fn zkb_ctx_small_cluster_max(ctx: *const zkb_ctx, ctas: *mut i32) -> i32;
    pub fn zkb_fft_evaluate(ctx: *mut zkb_ctx, coeffs_mont: *const u64, n: u64, evals_mont: *mut u64) -> i32;
    pub fn zkb_fft_interpolate(ctx: *mut zkb_ctx, evals_mont: *const u64, n: u64, coeffs_mont: *mut u64) -> i32;
    pub fn zkb_mle_ntt(ctx: *mut zkb_ctx, input: zkb_mle, inverse: i32, out: *mut zkb_mle) -> i32;
    pub fn zkb_merkle_build(ctx: *mut zkb_ctx, inputs_mont: *const u64, n_inputs: u64, depth: u32, out: *mut zkb_merkle) -> i32;
    pub fn zkb_merkle_free(ctx: *mut zkb_ctx, t: zkb_merkle) -> i32;
    pub fn zkb_merkle_root(ctx: *mut zkb_ctx, t: zkb_merkle, out: *mut u64) -> i32;
    pub fn zkb_merkle_nodes(ctx: *mut zkb_ctx, t: zkb_merkle, level: u32, first: u64, count: u64, out_mont: *mut u64) -> i32;
    pub fn zkb_merkle_update_leaf(ctx: *mut zkb_ctx, t: zkb_merkle, leaf_id: u64, data: *const u64, is_hash: i32) -> i32;
    pub fn zkb_merkle_create_proof(ctx: *mut zkb_ctx, t: zkb_merkle, data: *const u64, leaf_id: u64, sibling_hashes: *mut u64, sides: *mut u8) -> i32;
    pub fn zkb_merkle_verify(ctx: *mut zkb_ctx, t: zkb_merkle, data: *const u64, sibling_hashes: *const u64, sides: *const u8, n: u32, ok: *mut i32) -> i32;
    pub fn zkb_bench_modmul(ctx: *mut zkb_ctx, variant: i32, iters: u32, modmuls_per_s: *mut f64) -> i32;
    pub fn zkb_bench_imad(ctx: *mut zkb_ctx, mode: i32, iters: u32, ops_per_s: *mut f64) -> i32;
}
