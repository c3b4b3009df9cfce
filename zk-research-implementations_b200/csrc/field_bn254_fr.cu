#define ZKB_FIELD Bn254Fr
#define ZKB_FIELD_FN field_kernels_bn254_fr
#include "field_impl.cuh"

namespace zkb {
void launch_gather_elems(const GatherArgs& a, cudaStream_t s) { k_gather_elems<<<1, 32, 0, s>>>(a); }
int launch_tc_probe(unsigned int* counter, unsigned int* failed, int grid, int smem, long long timeout_clocks, cudaStream_t s) {
    cudaFuncSetAttribute(k_tc_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k_tc_probe, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    k_tc_probe<<<grid, BLOCK, smem, s>>>(counter, failed, timeout_clocks);
    return (int)cudaGetLastError();
}
void launch_bench_imad(int mode, uint64_t* out, uint32_t iters, int grid, cudaStream_t s) {
    if (mode == 0) k_bench_imad<0><<<grid, BLOCK, 0, s>>>(out, iters, 3u, 5u);
    else if (mode == 1) k_bench_imad<1><<<grid, BLOCK, 0, s>>>(out, iters, 3u, 5u);
    else if (mode == 2) k_bench_imad<2><<<grid, BLOCK, 0, s>>>(out, iters, 3u, 5u);
    else k_bench_imad<3><<<grid, BLOCK, 0, s>>>(out, iters, 3u, 5u);
}
}  // namespace zkb
