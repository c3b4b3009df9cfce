// launch.h -- per-field launcher table.  Each field (BN254 Fr, BN254 Fq,
// BLS12-381 Fr) is one translation unit instantiating kernels.cuh with its
// modulus as compile-time immediates; the host driver picks a table by id.
#pragma once
#include <cuda_runtime.h>

#include "kernels.cuh"

namespace zkb {

struct FieldKernels {
    int field_id;
    // composed-sumcheck kernels; kind/D/npts select the instantiation. Return false if unsupported.
    bool (*sc_eval)(int kind, int D, int npts, const ScArgs& a, int grid, cudaStream_t s);
    bool (*sc_fold_eval)(int kind, int D, int npts, const ScArgs& a, int grid, cudaStream_t s);
    // round 0 of a product of two or three factors with the sums of products as a Gram matrix on the tensor cores (tcfold.cuh)
    bool (*sc_eval_tc)(int kind, int D, int npts, const ScArgs& a, int grid, cudaStream_t s);
    // the same round with the folds on the tensor cores (tcfold.cuh); false if the shape is not instantiated
    bool (*sc_fold_eval_tc)(int kind, int D, int npts, const ScArgsTc& a, int grid, cudaStream_t s);
    // persistent round kernel (cooperative launch; tc: tensor-core variant, plain launch of a grid the caller sized to be
    // co-resident); returns a cudaError_t, or -1 if the shape is not instantiated
    int (*sc_tail)(int kind, int D, int npts, bool tc, const TailArgs& a, int grid, cudaStream_t s);
    // single-CTA shared-memory kernel for small tables; returns a cudaError_t, or -1 if not instantiated
    int (*sc_small)(int kind, int D, int npts, const SmallArgs& a, cudaStream_t s);
    int (*sc_small_max_cluster)();  // largest thread-block cluster of k_sc_small the device places (1 = none)
    // resident CTAs per SM for that instantiation (0 = unsupported); fused: 0 = k_sc_eval, 1 = k_sc_fold_eval, 2 = k_sc_tail,
    // 3 = k_sc_fold_eval_tc, 4 = k_sc_tail<TC>, 5 = k_sc_eval_tc / k_sc_eval_gram
    int (*sc_occupancy)(int fused, int kind, int D, int npts);
    void (*fold_tables)(const FoldTablesArgs& a, int grid, cudaStream_t s);
    void (*final_bind)(const FoldTablesArgs& a, Fe* out, volatile unsigned int* flag, unsigned int seq, cudaStream_t s);
    void (*multifold)(int k, const MultiFoldArgs& a, int grid, cudaStream_t s);
    void (*multifold_tc)(const MultiFoldTcArgs& a, int grid, cudaStream_t s);  // three variables per pass on the tensor cores
    void (*fold)(TabRef in, TabRef out, uint64_t n_out, uint32_t shift, const FixedMul& rt, int grid, cudaStream_t s);
    void (*aos_to_planar)(const void* aos, TabRef out, uint64_t n, uint64_t first, uint64_t stride, int conv, int grid, cudaStream_t s);
    void (*planar_to_aos)(TabRef in, void* aos, uint64_t n, int conv, int grid, cudaStream_t s);
    void (*interleave_shards)(TabRef gathered, uint64_t pitch, TabRef out, uint64_t n_local, uint32_t log2g, int grid, cudaStream_t s);
    void (*generate)(TabRef out, uint64_t n, uint64_t seed, uint64_t table, uint64_t first, uint64_t stride, int grid, cudaStream_t s);
    void (*vec_op)(TabRef x, TabRef y, TabRef out, uint64_t n, int op, int grid, cudaStream_t s);
    void (*axpby)(TabRef x, TabRef y, TabRef out, uint64_t n, const Fe& alpha, const Fe& beta, int grid, cudaStream_t s);
    void (*tensor)(TabRef x, TabRef y, TabRef out, uint64_t na, uint64_t nb, int op, int grid, cudaStream_t s);
    void (*layer_eval)(TabRef in, TabRef out, const uint8_t* ops, uint64_t n_gates, int grid, cudaStream_t s);
    void (*add_mul_i)(const uint8_t* ops, uint32_t n_gates, int op, int w, TabRef out, cudaStream_t s);
    void (*eq_split)(const ChalList& r, int n, int n_hi, TabRef hi, TabRef lo, int grid, cudaStream_t s);
    void (*gkr_phase1)(const GkrP1Args& a, int grid, cudaStream_t s);
    void (*gkr_phase2)(const GkrP2Args& a, int grid, cudaStream_t s);
    void (*gkr_wiring)(const GkrWiringArgs& a, int grid, cudaStream_t s);
    void (*gkr_w_phase1)(const GkrW1Args& a, int grid_gates, int grid_wires, cudaStream_t s);  // two launches each
    void (*gkr_w_phase2)(const GkrW2Args& a, int grid_gates, int grid_wires, cudaStream_t s);
    void (*gkr_w_wiring)(const GkrWWiringArgs& a, int grid, cudaStream_t s);
    void (*layer_eval_w)(TabRef in, TabRef out, const uint8_t* ops, const uint32_t* in1, const uint32_t* in2, uint64_t n_gates, int grid, cudaStream_t s);
    void (*bench_mul)(int variant, Fe* out, uint32_t iters, int grid, cudaStream_t s);
    // fft/src/fft.rs and merkle_tree/src/merkle_tree.rs (ntt_merkle.cuh)
    void (*ntt_twiddles)(TabRef lo, uint32_t lo_bits, TabRef hi, uint64_t n_hi, const NttPows& pw, int grid, cudaStream_t s);
    void (*ntt_pass)(const NttArgs& a, int grid, cudaStream_t s);
    void (*merkle_leaves)(const Fe* inputs, uint64_t n_inputs, Fe* leaves, uint64_t n_leaves, int grid, cudaStream_t s);
    void (*merkle_level)(const Fe* prev, Fe* next, uint64_t n_next, int grid, cudaStream_t s);
    void (*merkle_path)(Fe* tree, uint32_t depth, uint64_t leaf_id, const Fe& data, int is_hash, int mode, Fe* siblings, unsigned int* status, cudaStream_t s);
    // host-side arithmetic on Montgomery residues (fr.cuh compiled for the host)
    void (*h_add)(const Fe& a, const Fe& b, Fe& r);
    void (*h_sub)(const Fe& a, const Fe& b, Fe& r);
    void (*h_mul)(const Fe& a, const Fe& b, Fe& r);
    void (*h_to_mont)(const Fe& a, Fe& r);
    void (*h_from_mont)(const Fe& a, Fe& r);
    void (*h_modulus)(Fe& p);
};

const FieldKernels* field_kernels_bn254_fr();
const FieldKernels* field_kernels_bn254_fq();
const FieldKernels* field_kernels_bls12_381_fr();
void launch_gather_elems(const GatherArgs& a, cudaStream_t s);
int launch_tc_probe(unsigned int* counter, unsigned int* failed, int grid, int smem, long long timeout_clocks, cudaStream_t s);  // cudaError_t
void launch_bench_imad(int mode, uint64_t* out, uint32_t iters, int grid, cudaStream_t s);

}  // namespace zkb
