// tcfold.cu -- prototype + check of the tensor-core fold (tcgen05.mma kind::i8): one challenge bound in one table,
//   out[j] = a[j] + r (a[j + n_out] - a[j]) = a[j] (1 - r) + a[j + n_out] r,
// as TWO u8 x u8 -> s32 matrix products per 128 entries: rows = the 32 bytes of an element exactly as they lie in the
// planar table (plane 0 = bytes 0..15, plane 1 = bytes 16..31 -- which IS the K-major no-swizzle UMMA operand layout with
// LBO = plane distance, SBO = 128), columns = the 32 bytes of T1_i = (1 - r) 2^(8 i + 32) mod p resp. T2_i = r 2^(8 i + 32)
// mod p.  The 32 column sums (< 2^22 each) are carried into 9 limbs, one 32-bit Montgomery row divides by 2^32, one
// conditional subtraction makes the result canonical: bit-identical to Field::fold_fixed.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I zk-research-implementations_b200/csrc -I include tools/tcfold.cu -o build/kb/tcfold
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "kernels.cuh"
#include "tcfold.cuh"

using namespace zkb;
typedef Bn254Fr FT;
#define CK(x)                                                                          \
    do {                                                                               \
        cudaError_t e_ = (x);                                                          \
        if (e_ != cudaSuccess) {                                                       \
            fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
            exit(1);                                                                   \
        }                                                                              \
    } while (0)

// reference: the CUDA-core fold of kernels.cuh
__global__ void k_ref_fold(TabRef in, TabRef out, uint64_t n_out, const FixedMul rt) {
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n_out; j += (uint64_t)gridDim.x * blockDim.x)
        st_fe(out, j, Field<FT>::fold_fixed(ld_fe(in, j), ld_fe(in, j + n_out), rt));
}

int main(int argc, char** argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 20;
    const int reps = argc > 2 ? atoi(argv[2]) : 3;
    const uint64_t N = 1ull << n, n_out = N / 2;
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    void *p_in, *p_ref, *p_tc;
    CK(cudaMalloc(&p_in, N * 32));
    CK(cudaMalloc(&p_ref, n_out * 32));
    CK(cudaMalloc(&p_tc, n_out * 32));
    CK(cudaMemset(p_tc, 0xff, n_out * 32));
    TabRef in{(uint4*)p_in, N}, ref{(uint4*)p_ref, n_out}, tc{(uint4*)p_tc, n_out};
    k_generate<FT><<<sms * 8, BLOCK>>>(in, N, 0xB2000002ull, 0, 0, 1);
    // challenge (Montgomery form), its FixedMul table and the two byte matrices
    Fe r = Field<FT>::r2();
    r.l[0] ^= 0x1234567u;
    r = Field<FT>::mul(r, Field<FT>::r2());
    FixedMul rt;
    {
        Fe v = Field<FT>::zero();
        v.l[0] = 1;
        for (int k = 0; k < 64; ++k) v = Field<FT>::add(v, v);
        for (int i = 0; i < 8; ++i) {
            Fe t = Field<FT>::mul(r, v);
            memcpy(rt.t[i], t.l, 32);
            for (int k = 0; k < 32; ++k) v = Field<FT>::add(v, v);
        }
    }
    TcFoldMats hm;
    tc_fold_mats<FT>(r, &hm);
    TcFoldMats* dm;
    CK(cudaMalloc(&dm, sizeof(TcFoldMats)));
    CK(cudaMemcpy(dm, &hm, sizeof hm, cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    float best_ref = 1e30f, best_tc = 1e30f;
    for (int it = 0; it < reps; ++it) {
        CK(cudaEventRecord(e0));
        k_ref_fold<<<sms * 8, 256>>>(in, ref, n_out, rt);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best_ref) best_ref = ms;
    }
    auto kern = k_tc_fold<FT>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_FOLD_SMEM));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, TC_FOLD_THREADS, TC_FOLD_SMEM));
    const uint64_t tiles = (n_out + 127) / 128;
    const int grid = (int)(tiles < (uint64_t)sms * occ ? tiles : (uint64_t)sms * occ);
    for (int it = 0; it < reps; ++it) {
        CK(cudaEventRecord(e0));
        kern<<<grid, TC_FOLD_THREADS, TC_FOLD_SMEM>>>(in, tc, n_out, dm);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaGetLastError());
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best_tc) best_tc = ms;
    }
    std::vector<uint32_t> a(n_out * 8), b(n_out * 8);
    CK(cudaMemcpy(a.data(), p_ref, n_out * 32, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(b.data(), p_tc, n_out * 32, cudaMemcpyDeviceToHost));
    uint64_t bad = 0, first = ~0ull;
    for (uint64_t i = 0; i < n_out * 8; ++i)
        if (a[i] != b[i]) {
            if (!bad) first = i;
            ++bad;
        }
    const double bytes = 48.0 * (double)N;
    printf("fold 2^%d -> 2^%d: CUDA cores %.3f ms (%.0f GB/s), tensor cores %.3f ms (%.0f GB/s), occ=%d grid=%d, mismatching words %llu of %llu (first %llu)\n",
           n, n - 1, best_ref, bytes / best_ref / 1e6, best_tc, bytes / best_tc / 1e6, occ, grid, (unsigned long long)bad,
           (unsigned long long)(n_out * 8), (unsigned long long)first);
    if (bad) {
        const uint64_t w = first % (n_out * 4), pl = first / (n_out * 4);
        printf("  first mismatch: plane %llu entry %llu word %llu: ref %08x tc %08x\n", (unsigned long long)pl, (unsigned long long)(w / 4),
               (unsigned long long)(w % 4), a[first], b[first]);
        for (int k = 0; k < 8; ++k) printf("  ref[%d]=%08x tc[%d]=%08x\n", k, a[(k / 4) * n_out * 4 + (w / 4) * 4 + k % 4], k, b[(k / 4) * n_out * 4 + (w / 4) * 4 + k % 4]);
    }
    return bad ? 1 : 0;
}
