// field_impl.cuh -- included once per field TU with ZKB_FIELD / ZKB_FIELD_FN set.
#include "launch.h"

namespace zkb {
namespace {
typedef ZKB_FIELD FT;

// (kind, D, npts) instantiations: PROD(1,2) plain sumcheck; PROD(2,3) the GKR /
// compat shape; PROD(2,4), PROD(2,5) compat with 3 or 4 declared factors
// (reference quirk: only factors 0,1 multiply); PROD(3,4), PROD(4,5)
// full products; XYZ(2,3) two-phase GKR.
#define ZKB_SC_CASES(X) \
    X(KIND_PROD, 1, 2) X(KIND_PROD, 2, 3) X(KIND_PROD, 2, 4) X(KIND_PROD, 2, 5) X(KIND_PROD, 3, 4) X(KIND_PROD, 4, 5) \
    X(KIND_XYZ, 2, 3)

bool l_sc_eval(int kind, int D, int npts, const ScArgs& a, int grid, cudaStream_t s) {
#define X(K, DD, NP) \
    if (kind == K && D == DD && npts == NP) { \
        static bool once = (cudaFuncSetAttribute(k_sc_eval<FT, K, DD, NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, STAGE_BYTES), true); \
        (void)once; \
        k_sc_eval<FT, K, DD, NP><<<grid, BLOCK, STAGE_BYTES, s>>>(a); \
        return true; \
    }
    ZKB_SC_CASES(X)
#undef X
    return false;
}
bool l_sc_fold_eval(int kind, int D, int npts, const ScArgs& a, int grid, cudaStream_t s) {
#define X(K, DD, NP) \
    if (kind == K && D == DD && npts == NP) { \
        constexpr int SM = FoldSmem<K, DD, NP>::bytes; \
        static bool once = (cudaFuncSetAttribute(k_sc_fold_eval<FT, K, DD, NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM), true); \
        (void)once; \
        k_sc_fold_eval<FT, K, DD, NP><<<grid, BLOCK, SM, s>>>(a); \
        return true; \
    }
    ZKB_SC_CASES(X)
#undef X
    return false;
}
#define ZKB_TC_CASES(X) X(KIND_PROD, 2, 3) X(KIND_PROD, 2, 4) X(KIND_PROD, 3, 4)
bool l_sc_eval_tc(int kind, int D, int npts, const ScArgs& a, int grid, cudaStream_t s) {
#define X(NP) \
    if (kind == KIND_PROD && D == 2 && npts == NP) { \
        static bool once = (cudaFuncSetAttribute(k_sc_eval_tc<FT, NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, TCG_SMEM), cudaFuncSetAttribute(k_sc_eval_tc<FT, NP>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared), true); \
        (void)once; \
        k_sc_eval_tc<FT, NP><<<grid, BLOCK, TCG_SMEM, s>>>(a); \
        return true; \
    }
    X(3) X(4) X(5)
#undef X
    if (kind == KIND_PROD && D == 3 && npts == 4) {
        constexpr int SM = TcGramEvalSmem<4>::bytes;
        static bool once = (cudaFuncSetAttribute(k_sc_eval_gram<FT, 3, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM), cudaFuncSetAttribute(k_sc_eval_gram<FT, 3, 4>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared), true);
        (void)once;
        k_sc_eval_gram<FT, 3, 4><<<grid, BLOCK, SM, s>>>(a);
        return true;
    }
    return false;
}
bool l_sc_fold_eval_tc(int kind, int D, int npts, const ScArgsTc& a, int grid, cudaStream_t s) {
#define X(K, DD, NP) \
    if (kind == K && D == DD && npts == NP) { \
        constexpr int SM = TcFoldEvalCfg<DD, NP>::smem; \
        static bool once = (cudaFuncSetAttribute(k_sc_fold_eval_tc<FT, K, DD, NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM), cudaFuncSetAttribute(k_sc_fold_eval_tc<FT, K, DD, NP>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared), true); \
        (void)once; \
        k_sc_fold_eval_tc<FT, K, DD, NP><<<grid, BLOCK, SM, s>>>(a); \
        return true; \
    }
    ZKB_TC_CASES(X)
#undef X
    return false;
}
int l_sc_tail(int kind, int D, int npts, bool tc, const TailArgs& a, int grid, cudaStream_t s) {
    void* params[1] = {const_cast<TailArgs*>(&a)};
    if (tc) {
#define X(K, DD, NP) \
    if (kind == K && D == DD && npts == NP) { \
        constexpr int SM = TailSmemTc<NP>::bytes; \
        static bool once = (cudaFuncSetAttribute(k_sc_tail<FT, K, DD, NP, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM), cudaFuncSetAttribute(k_sc_tail<FT, K, DD, NP, true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared), true); \
        (void)once; \
        k_sc_tail<FT, K, DD, NP, true><<<grid, BLOCK, SM, s>>>(a); \
        return (int)cudaGetLastError(); \
    }
        ZKB_TC_CASES(X)
#undef X
        return -1;
    }
#define X(K, DD, NP) \
    if (kind == K && D == DD && npts == NP) { \
        constexpr int SM = TailSmem<K, DD, NP>::bytes; \
        static bool once = (cudaFuncSetAttribute(k_sc_tail<FT, K, DD, NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM), true); \
        (void)once; \
        return (int)cudaLaunchCooperativeKernel((const void*)k_sc_tail<FT, K, DD, NP>, dim3(grid), dim3(BLOCK), params, SM, s); \
    }
    ZKB_SC_CASES(X)
#undef X
    return -1;
}
int l_sc_small(int kind, int D, int npts, const SmallArgs& a, cudaStream_t s) {
    const int T = a.n_tables;
    const int nc = a.nc > 1 ? a.nc : 1;
    const size_t smem = (size_t)T * (a.n_in / nc) * 32;
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    cfg.gridDim = dim3(nc);
    cfg.blockDim = dim3(SMALL_BLOCK);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = nc;
    attr[0].val.clusterDim.y = attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = nc > 1 ? 1 : 0;
#define X(K, DD, NP) \
    if (kind == K && D == DD && npts == NP) { \
        static bool once = (cudaFuncSetAttribute(k_sc_small<FT, K, DD, NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMALL_SMEM_MAX), \
                            cudaFuncSetAttribute(k_sc_small<FT, K, DD, NP>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1), true); \
        (void)once; \
        return (int)cudaLaunchKernelEx(&cfg, k_sc_small<FT, K, DD, NP>, a); \
    }
    ZKB_SC_CASES(X)
#undef X
    return -1;
}
// largest cluster (<= SMALL_MAX_CLUSTER CTAs, each with the full shared-memory budget) this device can place
int l_sc_small_max_cluster() {
    auto kern = k_sc_small<FT, KIND_PROD, 2, 3>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMALL_SMEM_MAX);
    cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    for (int nc = SMALL_MAX_CLUSTER; nc > 1; nc >>= 1) {
        cudaLaunchConfig_t cfg = {};
        cudaLaunchAttribute attr[1];
        cfg.gridDim = dim3(nc);
        cfg.blockDim = dim3(SMALL_BLOCK);
        cfg.dynamicSmemBytes = SMALL_SMEM_MAX;
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = nc;
        attr[0].val.clusterDim.y = attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) == cudaSuccess && n >= 1) return nc;
        cudaGetLastError();
    }
    return 1;
}
int l_sc_occupancy(int fused, int kind, int D, int npts) {
    int nb = 0;
    if (fused == 5) return kind == KIND_PROD && D == 2 && npts <= 5 ? 1 : (kind == KIND_PROD && D == 3 && npts == 4 ? 2 : 0);  // k_sc_eval_tc / _gram
    if (fused >= 3) {
        // cudaOccupancyMaxActiveBlocksPerMultiprocessor answers 1 for every kernel that allocates tensor memory; two CTAs
        // of 128 columns each are co-resident (measured: 2 x 148 CTAs run in the time of one wave), so the count
        // follows from shared memory and registers (<= 128 by __launch_bounds__(BLOCK, 2)) alone
#define X(K, DD, NP) \
    if (kind == K && D == DD && npts == NP) { \
        const int sm = fused == 3 ? TcFoldEvalCfg<DD, NP>::smem : TailSmemTc<NP>::bytes; \
        return 2 * (sm + 1024) + 2048 <= 227 * 1024 ? 2 : 1; \
    }
        ZKB_TC_CASES(X)
#undef X
        return 0;
    }
#define X(K, DD, NP) \
    if (kind == K && D == DD && npts == NP) { \
        if (fused == 2) { \
            constexpr int SM = TailSmem<K, DD, NP>::bytes; \
            cudaFuncSetAttribute(k_sc_tail<FT, K, DD, NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM); \
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_sc_tail<FT, K, DD, NP>, BLOCK, SM); \
        } \
        else if (fused) { \
            constexpr int SM = FoldSmem<K, DD, NP>::bytes; \
            cudaFuncSetAttribute(k_sc_fold_eval<FT, K, DD, NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM); \
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_sc_fold_eval<FT, K, DD, NP>, BLOCK, SM); \
        } else { \
            cudaFuncSetAttribute(k_sc_eval<FT, K, DD, NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, STAGE_BYTES); \
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_sc_eval<FT, K, DD, NP>, BLOCK, STAGE_BYTES); \
        } \
        return nb; \
    }
    ZKB_SC_CASES(X)
#undef X
    return 0;
}
void l_fold_tables(const FoldTablesArgs& a, int grid, cudaStream_t s) { k_fold_tables<FT><<<grid, BLOCK, 0, s>>>(a); }
void l_final_bind(const FoldTablesArgs& a, Fe* out, volatile unsigned int* flag, unsigned int seq, cudaStream_t s) {
    k_final_bind<FT><<<1, 32, 0, s>>>(a, out, flag, seq);
}
void l_multifold(int k, const MultiFoldArgs& a, int grid, cudaStream_t s) {
    if (k == 3) k_multifold<FT, 3><<<grid, BLOCK, 0, s>>>(a);
    else if (k == 2) k_multifold<FT, 2><<<grid, BLOCK, 0, s>>>(a);
    else k_multifold<FT, 1><<<grid, BLOCK, 0, s>>>(a);
}
void l_multifold_tc(const MultiFoldTcArgs& a, int grid, cudaStream_t s) {
    static bool once = (cudaFuncSetAttribute(k_multifold_tc<FT>, cudaFuncAttributeMaxDynamicSharedMemorySize, TCM_SMEM), cudaFuncSetAttribute(k_multifold_tc<FT>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared), true);
    (void)once;
    k_multifold_tc<FT><<<grid, TCM_THREADS, TCM_SMEM, s>>>(a);
}
void l_fold(TabRef in, TabRef out, uint64_t n_out, uint32_t shift, const FixedMul& rt, int grid, cudaStream_t s) {
    k_fold<FT><<<grid, BLOCK, 0, s>>>(in, out, n_out, shift, rt);
}
void l_aos_to_planar(const void* aos, TabRef out, uint64_t n, uint64_t first, uint64_t stride, int conv, int grid, cudaStream_t s) {
    k_aos_to_planar<FT><<<grid, BLOCK, 0, s>>>((const uint4*)aos, out, n, first, stride, conv);
}
void l_planar_to_aos(TabRef in, void* aos, uint64_t n, int conv, int grid, cudaStream_t s) {
    k_planar_to_aos<FT><<<grid, BLOCK, 0, s>>>(in, (uint4*)aos, n, conv);
}
void l_interleave(TabRef g, uint64_t pitch, TabRef out, uint64_t n_local, uint32_t log2g, int grid, cudaStream_t s) {
    k_interleave_shards<FT><<<grid, BLOCK, 0, s>>>(g, pitch, out, n_local, log2g);
}
void l_generate(TabRef out, uint64_t n, uint64_t seed, uint64_t table, uint64_t first, uint64_t stride, int grid, cudaStream_t s) {
    k_generate<FT><<<grid, BLOCK, 0, s>>>(out, n, seed, table, first, stride);
}
void l_vec_op(TabRef x, TabRef y, TabRef out, uint64_t n, int op, int grid, cudaStream_t s) {
    k_vec_op<FT><<<grid, BLOCK, 0, s>>>(x, y, out, n, op);
}
void l_axpby(TabRef x, TabRef y, TabRef out, uint64_t n, const Fe& al, const Fe& be, int grid, cudaStream_t s) {
    k_axpby<FT><<<grid, BLOCK, 0, s>>>(x, y, out, n, al, be);
}
void l_tensor(TabRef x, TabRef y, TabRef out, uint64_t na, uint64_t nb, int op, int grid, cudaStream_t s) {
    k_tensor<FT><<<grid, BLOCK, 0, s>>>(x, y, out, na, nb, op);
}
void l_layer_eval(TabRef in, TabRef out, const uint8_t* ops, uint64_t n_gates, int grid, cudaStream_t s) {
    k_layer_eval<FT><<<grid, BLOCK, 0, s>>>(in, out, ops, n_gates);
}
void l_add_mul_i(const uint8_t* ops, uint32_t n_gates, int op, int w, TabRef out, cudaStream_t s) {
    k_add_mul_i<FT><<<(n_gates + BLOCK - 1) / BLOCK, BLOCK, 0, s>>>(ops, n_gates, op, w, out);
}
void l_eq_split(const ChalList& r, int n, int n_hi, TabRef hi, TabRef lo, int grid, cudaStream_t s) {
    k_eq_split<FT><<<grid, BLOCK, 0, s>>>(r, n, n_hi, hi, lo);
}
void l_gkr_phase1(const GkrP1Args& a, int grid, cudaStream_t s) { k_gkr_phase1<FT><<<grid, BLOCK, 0, s>>>(a); }
void l_gkr_phase2(const GkrP2Args& a, int grid, cudaStream_t s) { k_gkr_phase2<FT><<<grid, BLOCK, 0, s>>>(a); }
void l_gkr_wiring(const GkrWiringArgs& a, int grid, cudaStream_t s) { k_gkr_wiring<FT><<<grid, BLOCK, 0, s>>>(a); }
void l_gkr_w_phase1(const GkrW1Args& a, int grid_g, int grid_w, cudaStream_t s) {
    k_gkr_w_gates1<FT><<<grid_g, BLOCK, 0, s>>>(a);
    k_gkr_w_phase1<FT><<<grid_w, BLOCK, 0, s>>>(a);
}
void l_gkr_w_phase2(const GkrW2Args& a, int grid_g, int grid_w, cudaStream_t s) {
    k_gkr_w_gates2<FT><<<grid_g, BLOCK, 0, s>>>(a);
    k_gkr_w_phase2<FT><<<grid_w, BLOCK, 0, s>>>(a);
}
void l_gkr_w_wiring(const GkrWWiringArgs& a, int grid, cudaStream_t s) { k_gkr_w_wiring<FT><<<grid, BLOCK, 0, s>>>(a); }
void l_layer_eval_w(TabRef in, TabRef out, const uint8_t* ops, const uint32_t* in1, const uint32_t* in2, uint64_t n_gates, int grid, cudaStream_t s) {
    k_layer_eval_w<FT><<<grid, BLOCK, 0, s>>>(in, out, ops, in1, in2, n_gates);
}
void l_bench_mul(int variant, Fe* out, uint32_t iters, int grid, cudaStream_t s) {
    Fe seed = Field<FT>::r2();
    if (variant == 0) k_bench_mul<FT, 1, false><<<grid, BLOCK, 0, s>>>(out, iters, seed);
    else if (variant == 1) k_bench_mul<FT, 2, false><<<grid, BLOCK, 0, s>>>(out, iters, seed);
    else if (variant == 2) k_bench_mul<FT, 1, true><<<grid, BLOCK, 0, s>>>(out, iters, seed);
    else k_bench_mul<FT, 2, true><<<grid, BLOCK, 0, s>>>(out, iters, seed);
}
void l_ntt_twiddles(TabRef lo, uint32_t lo_bits, TabRef hi, uint64_t n_hi, const NttPows& pw, int grid, cudaStream_t s) {
    k_ntt_twiddles<FT><<<grid, BLOCK, 0, s>>>(lo, lo_bits, hi, n_hi, pw);
}
void l_ntt_pass(const NttArgs& a, int grid, cudaStream_t s) { k_ntt_pass<FT><<<grid, BLOCK, 0, s>>>(a); }
void l_merkle_leaves(const Fe* inputs, uint64_t n_inputs, Fe* leaves, uint64_t n_leaves, int grid, cudaStream_t s) {
    k_merkle_leaves<FT><<<grid, BLOCK, 0, s>>>(inputs, n_inputs, leaves, n_leaves);
}
void l_merkle_level(const Fe* prev, Fe* next, uint64_t n_next, int grid, cudaStream_t s) { k_merkle_level<FT><<<grid, BLOCK, 0, s>>>(prev, next, n_next); }
void l_merkle_path(Fe* tree, uint32_t depth, uint64_t leaf_id, const Fe& data, int is_hash, int mode, Fe* siblings, unsigned int* status, cudaStream_t s) {
    k_merkle_path<FT><<<1, 32, 0, s>>>(tree, depth, leaf_id, data, is_hash, mode, siblings, status);
}
void h_add(const Fe& a, const Fe& b, Fe& r) { r = Field<FT>::add(a, b); }
void h_sub(const Fe& a, const Fe& b, Fe& r) { r = Field<FT>::sub(a, b); }
void h_mul(const Fe& a, const Fe& b, Fe& r) { r = Field<FT>::mul(a, b); }
void h_to_mont(const Fe& a, Fe& r) { r = Field<FT>::to_mont(a); }
void h_from_mont(const Fe& a, Fe& r) { r = Field<FT>::from_mont(a); }
void h_modulus(Fe& p) {
    for (int i = 0; i < 8; ++i) p.l[i] = FT::P(i);
}

const FieldKernels TABLE = {
    FT::ID,      l_sc_eval,   l_sc_fold_eval, l_sc_eval_tc, l_sc_fold_eval_tc, l_sc_tail, l_sc_small, l_sc_small_max_cluster, l_sc_occupancy, l_fold_tables, l_final_bind, l_multifold, l_multifold_tc, l_fold,      l_aos_to_planar, l_planar_to_aos,
    l_interleave, l_generate, l_vec_op,       l_axpby,        l_tensor,      l_layer_eval, l_add_mul_i, l_eq_split,     l_gkr_phase1,
    l_gkr_phase2, l_gkr_wiring, l_gkr_w_phase1, l_gkr_w_phase2, l_gkr_w_wiring, l_layer_eval_w, l_bench_mul, l_ntt_twiddles, l_ntt_pass, l_merkle_leaves, l_merkle_level, l_merkle_path, h_add,         h_sub,          h_mul,         h_to_mont,   h_from_mont,     h_modulus,
};
}  // namespace

const FieldKernels* ZKB_FIELD_FN() { return &TABLE; }
}  // namespace zkb
