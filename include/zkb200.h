/*
 * zkb200.h -- C ABI of libzkb200.so: the B200-native sumcheck / GKR prover
 * engine that sits behind the crate API of obah/zk-research-implementations.
 *
 * The reference has no FFI of its own (pure Rust); this header IS
 * the drop-in boundary.  Every entry point names the reference item it replaces
 * (paths under /root/reference/).  The Rust side binds these through the thin
 * `zkb200-sys` crate shown in INTEGRATION.md; nothing but plain pointers and
 * sizes crosses the boundary.
 *
 * Conventions
 *  - Field elements cross as `uint64_t[4]`: the little-endian limbs of the
 *    MONTGOMERY residue (R = 2^256), i.e. byte-for-byte ark-ff 0.5's
 *    `Fp<MontBackend<_,4>>` (`.0.0`), so a `Vec<F>` is passed as `*const u64`
 *    with no conversion.  Canonical 32-byte little-endian encodings appear only
 *    where the reference serialises (`fq_vec_to_bytes`,
 *    fiat_shamir/src/fiat_shamir_transcript.rs:32-37).
 *  - Every function returns 0 on success or a negative zkb_status; nothing
 *    unwinds across the boundary.  The reference's panics map to error codes
 *    (the Rust shim re-raises the same panic strings).
 *  - A ctx owns one CUDA device, one stream and its scratch memory; it is not
 *    internally synchronised: one ctx per host thread.  Calls return after
 *    their (tiny) results are in host memory.
 *  - There is NO CPU fallback: creating a ctx without a usable CUDA device
 *    fails with ZKB_ERR_CUDA.
 */
#ifndef ZKB200_H
#define ZKB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct zkb_ctx zkb_ctx;
typedef struct zkb_transcript zkb_transcript;
typedef uint64_t zkb_mle;  /* opaque handle of a device-resident MultilinearPoly */
typedef uint64_t zkb_sp;   /* opaque handle of a device-resident SumPoly        */
typedef uint64_t zkb_circ; /* opaque handle of a device-resident Circuit        */
typedef uint64_t zkb_kzg;  /* opaque handle of a multilinear-KZG setup (G1 side)   */
typedef uint64_t zkb_merkle; /* opaque handle of a Keccak Merkle tree on the device */

typedef enum {
    ZKB_OK = 0,
    ZKB_ERR_BAD_ARG = -1,
    ZKB_ERR_NOT_POW2 = -2,        /* panic "Invalid evaluations"                    multilinear_polynomial_evaluation.rs:30 */
    ZKB_ERR_ARITY = -3,           /* panic "Invalid number of values"               :67, :81 */
    ZKB_ERR_LENGTH_MISMATCH = -4, /* panic "all evaluations must have same length"  composed_polynomial.rs:20 */
    ZKB_ERR_DEGREE_MISMATCH = -5, /* panic "all product polys must have same degree" :65 */
    ZKB_ERR_CUDA = -6,
    ZKB_ERR_NCCL = -7,
    ZKB_ERR_OOM = -8,
    ZKB_ERR_UNSUPPORTED = -9,     /* shape outside the instantiated kernels (degree > 4, > 16 tables) */
    ZKB_ERR_COMPAT_SHAPE = -10,   /* compat mode needs >= 2 products of >= 2 factors: the reference indexes
                                     polys[1] / evaluation[1] and panics otherwise (composed_polynomial.rs:53,90) */
    ZKB_ERR_CIRCUIT_SHAPE = -11   /* layer sizes the reference's fixed wiring cannot express */
} zkb_status;

enum { ZKB_FIELD_BN254_FR = 0, ZKB_FIELD_BN254_FQ = 1, ZKB_FIELD_BLS12_381_FR = 2 };
/* SumPoly::reduce semantics: COMPAT reproduces the reference exactly (only factors 0,1 of products 0,1
 * contribute, composed_polynomial.rs:52-54,88-99); FULL is the true sum of products. */
enum { ZKB_MODE_COMPAT = 0, ZKB_MODE_FULL = 1 };
enum { ZKB_OP_ADD = 0, ZKB_OP_MUL = 1, ZKB_OP_SUB = 2 }; /* Operation, multilinear_polynomial_evaluation.rs:4-17 */

const char* zkb_strerror(int32_t status);
const char* zkb_version(void);

/* ------------------------------------------------------------------ context */
int32_t zkb_ctx_create(int32_t field_id, int32_t device, int32_t mode, zkb_ctx** out);
int32_t zkb_ctx_destroy(zkb_ctx* ctx);
const char* zkb_ctx_last_error(const zkb_ctx* ctx);
/* The ctx's CUDA stream (cudaStream_t) for callers that time with CUDA events. */
void* zkb_ctx_stream(const zkb_ctx* ctx);
/* Number of this library's kernels launched through ctx so far (bench.py's `gpu_launches`). */
uint64_t zkb_ctx_launch_count(const zkb_ctx* ctx);
int32_t zkb_ctx_sync(zkb_ctx* ctx);
/* Per-launch device timing for the roofline report: while enabled, every table-kernel launch is bracketed by
 * CUDA events on the ctx's stream.  zkb_ctx_profile_read synchronises and returns, for one kernel id
 * (ZKB_K_*), the number of launches, their summed duration and their summed ALGORITHMIC bytes
 * (compulsory reads + writes, DESIGN.md "Kernels") since profiling was enabled. */
enum { ZKB_K_SC_EVAL = 0, ZKB_K_SC_FOLD_EVAL = 1, ZKB_K_FOLD_TABLES = 2, ZKB_K_FINAL_BIND = 3, ZKB_K_FOLD = 4,
       ZKB_K_LAYOUT = 5, ZKB_K_GKR_BUILD = 6, ZKB_K_OTHER = 7, ZKB_K_SC_TAIL = 8, ZKB_K_SC_SMALL = 9, ZKB_K_SC_TAIL_MID = 10, ZKB_K_COUNT = 11 };
int32_t zkb_ctx_profile(zkb_ctx* ctx, int32_t enable);
int32_t zkb_ctx_profile_read(zkb_ctx* ctx, int32_t kernel_id, uint64_t* launches, double* ms, double* alg_bytes);
const char* zkb_kernel_name(int32_t kernel_id);

/* Multi-GPU (one process per GPU).  The table is sharded on LOW index bits: rank j of 2^g holds entries
 * i with i mod 2^g == j, so both halves of every bound variable stay local.  Per round the
 * (d+1) partial sums of the ranks are combined by their host threads through a POSIX shared-memory segment
 * (ranks of one box; about 1 us) or, when that segment cannot be opened (ZKB200_NO_SHM=1, several hosts), with one
 * ncclAllReduce over zero-extended limbs; NCCL over NVLink carries the all-gather of the shrunken tables.
 * `unique_id` is the 128-byte ncclUniqueId produced by zkb_comm_unique_id on rank 0 and broadcast by the launcher. */
int32_t zkb_comm_unique_id(uint8_t out[128]);
int32_t zkb_ctx_comm_init(zkb_ctx* ctx, int32_t rank, int32_t world, const uint8_t unique_id[128]);
/* Local (per-rank) table size, as log2 entries, at which the shards are all-gathered and the remaining rounds run on
 * every rank alone.  0 (default) = automatic: as soon as the replicated table fits the on-chip kernel, or 2^14 when
 * the ranks' hosts cannot exchange the round sums through shared memory. */
int32_t zkb_ctx_set_gather_threshold(zkb_ctx* ctx, uint32_t log2_local_entries);
/* Rounds whose tables have at most 2^log2_entries entries run inside ONE persistent cooperative kernel that
 * exchanges round sums / challenges with the host transcript through a mailbox in mapped host memory
 * (no launch per round).  0 disables it (one launch per round).  Default 40 (always). */
/* Host only (no ctx, no GPU): the two byte matrices the tensor-core fold uses for a challenge r (Montgomery limbs in): out[0..1024)
 * for the (1 - r) rows, out[1024..2048) for the r rows; byte n of T_k = (1 - r) 2^(8 k + 32) mod p resp. r 2^(8 k + 32) mod p at
 * (k / 16) * 512 + n * 16 + k % 16 (csrc/tcfold.cuh).  Exposed so that the construction can be checked without a device. */
int32_t zkb_tc_fold_matrices(int32_t field_id, const uint64_t r_mont[4], uint8_t out[2048]);
/* Tensor-core paths (csrc/tcfold.cuh): *enabled = 1 unless ZKB200_NO_TC was set when the ctx was created; *persistent = 1 if
 * the persistent round kernel uses them too (two of its CTAs per SM were found co-resident by the creation-time probe;
 * otherwise the large rounds of the persistent kernel stay on the CUDA cores and only the per-round launches use them). */
int32_t zkb_ctx_tensor_cores(const zkb_ctx* ctx, int32_t* enabled, int32_t* persistent);
int32_t zkb_ctx_set_tail_threshold(zkb_ctx* ctx, uint32_t log2_entries);
/* CTAs of the largest thread-block cluster the on-chip kernel (next entry) is launched with: 16 on a B200 unless
 * ZKB200_CLUSTER_MAX (a power of two, 1 = single CTA) was set when the ctx was created or the device cannot place it. */
int32_t zkb_ctx_small_cluster_max(const zkb_ctx* ctx, int32_t* ctas);
/* Once all tables of a sumcheck fit in `smem_bytes` of shared memory per CTA (default and maximum 200 KiB) the remaining
 * rounds run in ONE launch that keeps the tables on chip -- a single CTA, or a thread-block cluster of up to 16 CTAs with
 * the tables sharded over their shared memories (16 x the size); 0 disables it. */
int32_t zkb_ctx_set_small_threshold(zkb_ctx* ctx, uint32_t smem_bytes);
/* Device-side Fiat-Shamir transcript (default OFF, ZKB200_DT=1 in the environment turns it on).  In the
 * on-chip kernel above, for round polynomials of degree <= 2 (the GKR layer sumcheck and products of two factors),
 * the GPU itself interpolates and trims the round message (univariate_polynomial_dense.rs:14-18,48-74), serialises
 * it (fiat_shamir_transcript.rs:32-37), runs Keccak-256 and reduces the digest mod p (:23-29), so no round waits for
 * PCIe.  The host transcript remains the source of truth: it replays every round from the sums the device reports
 * and the call fails with ZKB_ERR_CUDA ("device transcript diverged") if a challenge differs.  Bit-identical to the
 * host path, but measured slower on B200 (one warp needs 3.8 us per Keccak-f), hence off.
 * _stats: how many launches ran with the device transcript and how many device challenges the host has checked. */
int32_t zkb_ctx_set_device_transcript(zkb_ctx* ctx, int32_t enable);
int32_t zkb_ctx_device_transcript_stats(const zkb_ctx* ctx, uint64_t* launches, uint64_t* rounds_checked);

/* --------------------------------------------- MultilinearPoly (device table) */
/* MultilinearPoly::new (multilinear_polynomial_evaluation.rs:26-37): `len` must be a power of two.
 * Source-buffer lifetime: the host-to-device copy is queued, not awaited.  From PAGEABLE memory (a Rust Vec) the
 * driver stages the data before the call returns, so the buffer may be reused at once; from PINNED memory the
 * buffer must stay unchanged until the next call that returns a result from this table (any prove / evaluate /
 * download) or zkb_ctx_sync.  The same holds for zkb_mle_upload_shard and the `inputs` of zkb_circuit_evaluate /
 * zkb_gkr_prove*.
 * Table lifetime: a SumPoly holds a reference to the tables it was created from (the reference's SumPoly owns clones):
 * zkb_mle_free on such a table only marks it, the memory is released when the last SumPoly using it is freed. */
int32_t zkb_mle_upload(zkb_ctx* ctx, const uint64_t* aos_mont, uint64_t len, zkb_mle* out);
/* Rank-local shard of a host table that every rank holds in full (strided gather on upload). */
int32_t zkb_mle_upload_shard(zkb_ctx* ctx, const uint64_t* aos_mont_full, uint64_t len_full, zkb_mle* out);
/* Synthetic table generated on the device (bench inputs).  With a communicator attached the
 * call creates this rank's shard of the 2^n_vars-entry global table. */
int32_t zkb_mle_generate(zkb_ctx* ctx, uint64_t seed, uint64_t table_id, uint32_t n_vars, zkb_mle* out);
int32_t zkb_mle_download(zkb_ctx* ctx, zkb_mle m, uint64_t* aos_mont);
/* canonical 32-byte LE encodings, the bytes fq_vec_to_bytes would produce */
int32_t zkb_mle_download_canonical(zkb_ctx* ctx, zkb_mle m, uint8_t* bytes);
int32_t zkb_mle_clone(zkb_ctx* ctx, zkb_mle m, zkb_mle* out);
int32_t zkb_mle_free(zkb_ctx* ctx, zkb_mle m);
int32_t zkb_mle_num_vars(zkb_ctx* ctx, zkb_mle m, uint32_t* n_vars); /* local shard size */
/* partial_evaluate(bit, value) :52-63 */
int32_t zkb_mle_partial_evaluate(zkb_ctx* ctx, zkb_mle in, uint32_t bit, const uint64_t value[4], zkb_mle* out);
/* multi_partial_evaluate(values) :65-77 */
int32_t zkb_mle_multi_partial_evaluate(zkb_ctx* ctx, zkb_mle in, const uint64_t* values, uint32_t k, zkb_mle* out);
/* evaluate(values) :79-91 */
int32_t zkb_mle_evaluate(zkb_ctx* ctx, zkb_mle m, const uint64_t* values, uint32_t k, uint64_t out[4]);
/* [sum of low half, sum of high half]: get_round_partial_polynomial_proof, sum_check_protocol.rs:168-175 */
int32_t zkb_mle_sum_halves(zkb_ctx* ctx, zkb_mle m, uint64_t out[8]);
/* scale :93-97 */
int32_t zkb_mle_scale(zkb_ctx* ctx, zkb_mle in, const uint64_t value[4], zkb_mle* out);
/* impl Add / Mul / Sub :113-156 (op = ZKB_OP_*) */
int32_t zkb_mle_binary(zkb_ctx* ctx, zkb_mle a, zkb_mle b, int32_t op, zkb_mle* out);
/* tensor_add_mul_polynomials :99-111 (op = ZKB_OP_ADD | ZKB_OP_MUL) */
int32_t zkb_mle_tensor(zkb_ctx* ctx, zkb_mle a, zkb_mle b, int32_t op, zkb_mle* out);

/* ------------------------------------------ ProductPoly / SumPoly (composed) */
/* SumPoly::new over ProductPoly::new (composed_polynomial.rs:16-29,62-69): `tables` holds n_products *
 * degree handles, product-major.  The SumPoly takes its own copy of nothing: it references the tables
 * read-only and folds into private scratch, so the caller's tables stay intact (the reference clones). */
int32_t zkb_sumpoly_create(zkb_ctx* ctx, const zkb_mle* tables, uint32_t n_products, uint32_t degree, zkb_sp* out);
int32_t zkb_sumpoly_free(zkb_ctx* ctx, zkb_sp sp);
/* Back to the unbound state (the caller's tables were never written). */
int32_t zkb_sumpoly_reset(zkb_ctx* ctx, zkb_sp sp);
/* ProductPoly/SumPoly::evaluate (:31-36, :71-76): sum of products of the factors' evaluations. */
int32_t zkb_sumpoly_evaluate(zkb_ctx* ctx, zkb_sp sp, const uint64_t* values, uint32_t k, uint64_t out[4]);
/* Step API of the composed sumcheck (the body of gkr_prove's loop, sum_check_protocol.rs:96-108):
 *   round_evals : s(0..d) of the current round          (get_round_partial_polynomial_proof_gkr :152-166)
 *   bind_and_next: fold every table with r and return the next round's s(0..d) in ONE pass over HBM
 *   final_values: the n_products*degree fully bound table values after the last bind                */
int32_t zkb_sc_round_evals(zkb_ctx* ctx, zkb_sp sp, uint64_t* evals /* (d+1)*4 */);
int32_t zkb_sc_bind_and_next(zkb_ctx* ctx, zkb_sp sp, const uint64_t r[4], uint64_t* evals /* (d+1)*4, NULL on the last bind */);
int32_t zkb_sc_final_values(zkb_ctx* ctx, zkb_sp sp, uint64_t* values /* T*4 */);

/* ------------------------------------------------ Transcript (host, Keccak-256) */
/* fiat_shamir/src/fiat_shamir_transcript.rs:5-37 */
int32_t zkb_transcript_new(int32_t field_id, zkb_transcript** out);
int32_t zkb_transcript_free(zkb_transcript* t);
int32_t zkb_transcript_append(zkb_transcript* t, const uint8_t* bytes, size_t len);
int32_t zkb_transcript_append_elements(zkb_transcript* t, const uint64_t* mont, size_t n); /* append(&fq_vec_to_bytes(..)) */
int32_t zkb_transcript_challenge(zkb_transcript* t, uint64_t out_mont[4]);
int32_t zkb_keccak256(const uint8_t* bytes, size_t len, uint8_t out[32]);
/* Host conversion between canonical little-endian limbs and the Montgomery residues that cross this ABI
 * (what ark-ff's `F::from(BigInt)` / `into_bigint()` do; used by non-Rust callers). */
int32_t zkb_fe_to_mont(int32_t field_id, const uint64_t* canonical, uint64_t* mont, size_t n);
int32_t zkb_fe_from_mont(int32_t field_id, const uint64_t* mont, uint64_t* canonical, size_t n);
/* The host half of the per-round multi-GPU combine (C1): `wide` holds, for each of n values, the 8 32-bit limbs
 * of the ranks' residues summed as plain integers in 8 uint64 slots; carry-propagate and reduce mod p. */
int32_t zkb_fe_reduce_wide(int32_t field_id, const uint64_t* wide, uint64_t* out, size_t n);

/* ------------------------------------------------ UnivariatePoly (host, tiny) */
/* interpolate + trim (univariate_polynomial_dense.rs:48-74,14-18): returns the trimmed length in *len. */
int32_t zkb_uni_interpolate(int32_t field_id, const uint64_t* xs, const uint64_t* ys, uint32_t n, uint64_t* coeffs,
                            uint32_t* len);
/* evaluate :20-26 */
int32_t zkb_uni_evaluate(int32_t field_id, const uint64_t* coeffs, uint32_t len, const uint64_t x[4], uint64_t out[4]);

/* ------------------------------------------------ proof wire format (host only)
 * A stable byte encoding of sum_check's two proof structs (sum_check_protocol.rs:8-17); every field element is written
 * exactly as `fq_vec_to_bytes` writes it into the transcript (32-byte little-endian canonical integer,
 * fiat_shamir_transcript.rs:32-37):
 *   "ZKBP" | u8 version = 1 | u8 field_id | u8 kind | u8 0 | u32 n_rounds (LE) | claimed_sum (32 B) |
 *   per round: u8 len, len x 32 B
 * kind 1 = `Proof` (sum_check::prove): len = 2, the evaluations [s(0), s(1)]; `slots` must be 2, `lens` is ignored.
 * kind 2 = `GkrProof` (gkr_prove): the trimmed ascending coefficients (len <= slots; `msgs` is n_rounds x slots).
 * encode: out == NULL only returns the size in *len.  decode: msgs == NULL only checks the bytes and returns
 * field / kind / n_rounds; it rejects (ZKB_ERR_BAD_ARG) bad magic, truncation, trailing bytes, len > slots and any
 * element that is not a canonical residue, before writing anything. */
int32_t zkb_proof_encode(int32_t field_id, int32_t kind, uint32_t n_rounds, uint32_t slots, const uint64_t* msgs_mont,
                         const int32_t* lens, const uint64_t claimed_sum[4], uint8_t* out, size_t cap, size_t* len);
int32_t zkb_proof_decode(const uint8_t* bytes, size_t len, int32_t* field_id, int32_t* kind, uint32_t* n_rounds, uint32_t slots,
                         uint64_t* msgs_mont, int32_t* lens, uint64_t claimed_sum[4]);

/* -------------------------------------------------------- sum_check_protocol */
/* prove (sum_check_protocol.rs:25-52).  msgs: n_vars x 2 elements; challenges: n_vars elements (not part
 * of the reference's Proof, returned for callers that want them; may be NULL).
 * flags bit 0: absorb the whole table into the transcript first, as the reference does (:27).  Clear it
 * for the "seeded-transcript prover core" (SURVEY F9); the reference-faithful setting is 1. */
int32_t zkb_sumcheck_prove(zkb_ctx* ctx, zkb_mle poly, uint32_t flags, uint64_t claimed_sum[4], uint64_t* msgs,
                           uint64_t* challenges);
/* verify (:54-84): *accepted = 1/0. */
int32_t zkb_sumcheck_verify(zkb_ctx* ctx, zkb_mle poly, uint32_t flags, const uint64_t claimed_sum[4],
                            const uint64_t* msgs, uint32_t n_msgs, int32_t* accepted);
/* gkr_prove (:86-115).  coeffs: n_vars x (d+1) slots, lens[k] = trimmed length of round k's coefficient
 * vector (possibly 0); challenges: n_vars; final_values (may be NULL): the T bound table values. */
int32_t zkb_gkr_sumcheck_prove(zkb_ctx* ctx, zkb_transcript* t, const uint64_t claimed_sum[4], zkb_sp sp,
                               uint64_t* coeffs, int32_t* lens, uint64_t* challenges, uint64_t* final_values);
/* gkr_verify (:117-150).  slots = stride (in elements) between rounds in coeffs.  On rejection
 * final_claim = 0 and challenges[0] = 0 (:129-133). */
int32_t zkb_gkr_sumcheck_verify(zkb_transcript* t, uint32_t n_rounds, uint32_t slots, const uint64_t* coeffs,
                                const int32_t* lens, const uint64_t claimed_sum[4], int32_t* accepted,
                                uint64_t final_claim[4], uint64_t* challenges);

/* --------------------------------------------------------- gkr_circuit / gkr */
/* Circuit::new (gkr_circuit.rs:113-125): layers listed input side first; gate i of a layer reads wires
 * 2i, 2i+1 of the layer below (:76-78,132).  ops: concatenated ZKB_OP_ADD/ZKB_OP_MUL bytes. */
int32_t zkb_circuit_create(zkb_ctx* ctx, uint32_t n_layers, const uint32_t* gates_per_layer, const uint8_t* ops,
                           zkb_circ* out);
int32_t zkb_circuit_free(zkb_ctx* ctx, zkb_circ c);
/* Circuit::evaluate (:127-143): layer outputs stay on the device; `outputs` (may be NULL) receives them
 * concatenated, input-side layer first. */
int32_t zkb_circuit_evaluate(zkb_ctx* ctx, zkb_circ c, const uint64_t* inputs_mont, uint64_t n_inputs,
                             uint64_t* outputs_mont);
/* Layer::get_add_mul_i (gkr_circuit.rs:39-104): the DENSE indicator table of one layer's gates of operation `op`
 * (2^(3w+2) entries for 2^w > 1 gates, 8 for one gate).  For the reference's dense constructions
 * (get_fbc_poly, gkr_protocol.rs:243-292) on small layers; zkb_gkr_prove never materialises it.
 * n_gates must be a power of two and 3*log2(n_gates)+2 <= 30. */
int32_t zkb_layer_add_mul_i(zkb_ctx* ctx, const uint8_t* ops, uint32_t n_gates, int32_t op, zkb_mle* out);
/* gkr_protocol::prove (gkr_protocol.rs:31-91; the KZG input opening :92-118 is out of scope).
 * Outputs: w0[2] (output_poly); per layer (output side first) 2*(log2(2G)) rounds of 3 coefficient slots
 * with trimmed lengths; claimed[(L-1)][2] (claimed_evaluations); final_openings[2] = the input MLE at
 * (r_b, r_c) -- what the input-layer commitment opens; challenges (may be NULL).  *n_rounds receives the
 * total number of sumcheck rounds. */
int32_t zkb_gkr_prove(zkb_ctx* ctx, zkb_circ c, const uint64_t* inputs_mont, uint64_t n_inputs, uint64_t w0[8],
                      uint64_t* coeffs, int32_t* lens, uint64_t* challenges, uint64_t* claimed, uint64_t final_openings[8],
                      uint32_t* n_rounds);
/* gkr_protocol::verify (:128-227) with the input opening replaced by a direct evaluation of the input
 * MLE on the device and the wiring predicates evaluated in O(G) through eq tables. */
int32_t zkb_gkr_verify(zkb_ctx* ctx, zkb_circ c, const uint64_t* inputs_mont, uint64_t n_inputs, const uint64_t w0[8],
                       const uint64_t* coeffs, const int32_t* lens, const uint64_t* claimed, const uint64_t final_openings[8],
                       int32_t* accepted);
uint32_t zkb_gkr_total_rounds(uint32_t n_layers, const uint32_t* gates_per_layer);

/* ------------------------------------------- general wiring (EXTENSION beyond the reference) */
/* The reference wires gate i to (2i, 2i+1) of the layer below (gkr_circuit.rs:76-78,132), so every layer halves and
 * the output layer has 1-2 gates (gkr_protocol.rs:235).  BASELINE.json configs[2] ("2^20 gates per layer and 16
 * layers") needs gates that read arbitrary wires in1[g], in2[g] and a wide output layer.  These entry points run
 * the reference's protocol (gkr_protocol.rs:31-91,128-227: same transcript order, round polynomials, claim merge
 * alpha*o1+beta*o2 and openings) on such circuits; the only change is that initiate_protocol (:229-241) draws
 * log2(max(outputs,2)) consecutive challenges for the output MLE instead of one.  With in1 = 2g, in2 = 2g+1 and
 * <= 2 outputs the proof bytes equal zkb_gkr_prove's.
 * gates_per_layer: input side first, each a power of two; layer l reads the n_inputs inputs (l = 0) or the
 * gates_per_layer[l-1] outputs of layer l-1; ops / in1 / in2 are concatenated per gate in layer order.
 * The handle works with zkb_circuit_evaluate / zkb_circuit_free. */
int32_t zkb_circuit_create_wired(zkb_ctx* ctx, uint32_t n_layers, const uint32_t* gates_per_layer, uint64_t n_inputs,
                                 const uint8_t* ops, const uint32_t* in1, const uint32_t* in2, zkb_circ* out);
/* Total sumcheck rounds of a proof over circuit `c` (either kind). */
int32_t zkb_circuit_total_rounds(zkb_ctx* ctx, zkb_circ c, uint32_t* n_rounds);
/* As zkb_gkr_prove; w0 receives n_w0 = max(outputs, 2) elements (the whole output layer). */
int32_t zkb_gkr_prove_wired(zkb_ctx* ctx, zkb_circ c, const uint64_t* inputs_mont, uint64_t n_inputs, uint64_t* w0,
                            uint64_t n_w0, uint64_t* coeffs, int32_t* lens, uint64_t* challenges, uint64_t* claimed,
                            uint64_t final_openings[8], uint32_t* n_rounds);
int32_t zkb_gkr_verify_wired(zkb_ctx* ctx, zkb_circ c, const uint64_t* inputs_mont, uint64_t n_inputs, const uint64_t* w0,
                             uint64_t n_w0, const uint64_t* coeffs, const int32_t* lens, const uint64_t* claimed,
                             const uint64_t final_openings[8], int32_t* accepted);

/* ------------------------------------------------------------ microbenchmarks */
/* ------------------------------------------------ input-layer commitment: multilinear KZG over BLS12-381 G1
 * Prover side of pcs/src/kzg_pcs/kzg.rs, as gkr/src/gkr_protocol.rs:92-118 runs it on the input MLE.  The ctx must be
 * created with ZKB_FIELD_BLS12_381_FR (ZKB_ERR_UNSUPPORTED otherwise).  G1 points cross the ABI as 96 bytes: the affine
 * coordinates x, y as 48-byte little-endian canonical integers; 96 zero bytes = the point at infinity (the reference
 * compares projective points for equality, which is equality of these bytes).
 *
 * zkb_kzg_setup    KZG::new / run_trusted_setup / get_lagrange_basis (kzg.rs:17-49,183-212), G1 side: basis[i] =
 *                  g1 * eq(taus, i).  `taus`: n_vars Montgomery Fr elements -- an INPUT here (the reference draws them
 *                  from OS entropy inside gkr_protocol::prove, :95-101, which no implementation can reproduce).
 * zkb_kzg_basis    canonical bytes of `count` basis points from `first` (level 0 = the Lagrange basis; level k = the
 *                  basis folded over its k leading variables, see csrc/kzg_impl.cuh)
 * zkb_kzg_commit   KZG::commit (:51-53) = evaluate_poly_with_l_basis_in_g1 (:131-144), as a bucket MSM
 * zkb_kzg_open     KZG::open (:55-57) = poly.evaluate(opening_values)
 * zkb_kzg_get_proof  KZG::get_proof (:59-95): `n` quotient commitments (n x 96 bytes), one per opening value
 * The G2 powers and KZG::verify (:97-129, pairings) are verifier-side and not provided. */
int32_t zkb_kzg_setup(zkb_ctx* ctx, uint32_t n_vars, const uint64_t* taus_mont, zkb_kzg* out);
int32_t zkb_kzg_free(zkb_ctx* ctx, zkb_kzg k);
int32_t zkb_kzg_basis(zkb_ctx* ctx, zkb_kzg k, uint32_t level, uint64_t first, uint64_t count, uint8_t* out);
int32_t zkb_kzg_commit(zkb_ctx* ctx, zkb_kzg k, zkb_mle poly, uint8_t out[96]);
int32_t zkb_kzg_open(zkb_ctx* ctx, zkb_kzg k, zkb_mle poly, const uint64_t* opening_values, uint32_t n, uint64_t out[4]);
int32_t zkb_kzg_get_proof(zkb_ctx* ctx, zkb_kzg k, zkb_mle poly, const uint64_t opened_value[4], const uint64_t* opening_values, uint32_t n,
                          uint8_t* out);

/* ------------------------------------------------ NTT: fft/src/fft.rs
 * zkb_fft_evaluate     fft_evaluate (:31-41): evals[j] = sum_i coeffs[i] w^(i j), w = F::get_root_of_unity(n) (ark-ff: the field's
 *                      two-adic root squared down to order n), natural order in and out; n not a power of two:
 *                      ZKB_ERR_NOT_POW2 ("Length must be a power of 2", :34); n above the field's two-adic order (2^28 for
 *                      BN254 Fr, 2^32 for BLS12-381 Fr, 2 for BN254 Fq): ZKB_ERR_UNSUPPORTED (the reference unwraps None, :38)
 * zkb_fft_interpolate  fft_interpolate (:43-60): the inverse transform (w^-1, then n^-1); the coefficient vector comes back
 *                      at full length n (UnivariatePoly::new does not trim)
 * zkb_mle_ntt          the same transform on a device-resident table (inverse != 0: interpolate) */
int32_t zkb_fft_evaluate(zkb_ctx* ctx, const uint64_t* coeffs_mont, uint64_t n, uint64_t* evals_mont);
int32_t zkb_fft_interpolate(zkb_ctx* ctx, const uint64_t* evals_mont, uint64_t n, uint64_t* coeffs_mont);
int32_t zkb_mle_ntt(zkb_ctx* ctx, zkb_mle in, int32_t inverse, zkb_mle* out);

/* ------------------------------------------------ Keccak Merkle tree: merkle_tree/src/merkle_tree.rs
 * Nodes are field elements (Montgomery limbs at the ABI); compute_hash(x) = Keccak256(fq_vec_to_bytes([x])) and
 * hash_pair(l, r) = Keccak256(bytes(l) || bytes(r)), both mapped back with from_le_bytes_mod_order (:201-214).
 * zkb_merkle_build         MerkleTree::new (:31-50, n_inputs = 0) / new_with_inputs (:52-84): leaves[i] = compute_hash(input_i),
 *                          the other leaves are zero (not hashed); "Too many inputs for tree depth" -> ZKB_ERR_BAD_ARG
 * zkb_merkle_root          get_root_hash (:134-136)
 * zkb_merkle_nodes         `count` nodes of `level` from `first` (level 0 = the leaves, level depth = the root)
 * zkb_merkle_update_leaf   update_leaf + recompute_path (:86-132); "Invalid leaf ID" -> ZKB_ERR_BAD_ARG
 * zkb_merkle_create_proof  create_proof (:138-183): depth sibling hashes and their sides (0 = Left, 1 = Right);
 *                          "Data does not match the leaf hash" / "Invalid leaf ID" -> ZKB_ERR_BAD_ARG
 * zkb_merkle_verify        verify (:185-199): ok = 1 iff the path hashes to the stored root */
int32_t zkb_merkle_build(zkb_ctx* ctx, const uint64_t* inputs_mont, uint64_t n_inputs, uint32_t depth, zkb_merkle* out);
int32_t zkb_merkle_free(zkb_ctx* ctx, zkb_merkle t);
int32_t zkb_merkle_root(zkb_ctx* ctx, zkb_merkle t, uint64_t out[4]);
int32_t zkb_merkle_nodes(zkb_ctx* ctx, zkb_merkle t, uint32_t level, uint64_t first, uint64_t count, uint64_t* out_mont);
int32_t zkb_merkle_update_leaf(zkb_ctx* ctx, zkb_merkle t, uint64_t leaf_id, const uint64_t data[4], int32_t is_hash);
int32_t zkb_merkle_create_proof(zkb_ctx* ctx, zkb_merkle t, const uint64_t data[4], uint64_t leaf_id, uint64_t* sibling_hashes, uint8_t* sides);
int32_t zkb_merkle_verify(zkb_ctx* ctx, zkb_merkle t, const uint64_t data[4], const uint64_t* sibling_hashes, const uint8_t* sides, uint32_t n,
                          int32_t* ok);

/* Device-timed multiplier throughput (fills the IMAD-roofline denominator).
 * variant: 0/1 = IMAD.WIDE multiplier with 1/2 independent chains per thread, 2/3 = 32-bit lo/hi
 * multiplier.  Returns modmuls per second. */
int32_t zkb_bench_modmul(zkb_ctx* ctx, int32_t variant, uint32_t iters, double* modmuls_per_s);
/* mode 0 IMAD, 1 IMAD.HI, 2 mad.wide, 3 IMAD.WIDE.X carry chain; returns multiply instructions/s */
int32_t zkb_bench_imad(zkb_ctx* ctx, int32_t mode, uint32_t iters, double* ops_per_s);

#ifdef __cplusplus
}
#endif
#endif /* ZKB200_H */
