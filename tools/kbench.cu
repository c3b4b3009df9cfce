// kbench.cu -- stand-alone timing harness for ONE shape of the round kernels (k_sc_eval / k_sc_fold_eval), used to
// iterate on a kernel variant without rebuilding the library (the three field TUs take two minutes):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I zk-research-implementations_b200/csrc
//        -I include [-DKB_D=3] [-DZKB_...=..] tools/kbench.cu -o gpurun_out/kbench_<variant>
//   gpurun -- 'gpurun_out/kbench_<variant> 26 5'      (n_vars, repetitions)
// Prints the CUDA-event time of the first-round kernel and of the largest fold round, the achieved GB/s on the
// algorithmic bytes, and a hash of the sums and of the folded tables: two variants of a kernel must print the same hash
// (bit-exactness against the oracle is the parity suite's job, through the library).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "kernels.cuh"
#include "tcfold.cuh"

#ifndef KB_D
#define KB_D 3
#endif
#ifndef KB_P
#define KB_P 1
#endif
using namespace zkb;
typedef Bn254Fr FT;
constexpr int D = KB_D, NPTS = KB_D + 1, P = KB_P, T = KB_P * KB_D;

#define CK(x)                                                                          \
    do {                                                                               \
        cudaError_t e_ = (x);                                                          \
        if (e_ != cudaSuccess) {                                                       \
            fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
            exit(1);                                                                   \
        }                                                                              \
    } while (0)

static uint64_t fnv(uint64_t h, const void* p, size_t n) {
    const unsigned char* b = (const unsigned char*)p;
    for (size_t i = 0; i < n; ++i) h = (h ^ b[i]) * 1099511628211ull;
    return h;
}

int main(int argc, char** argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 24;
    const int reps = argc > 2 ? atoi(argv[2]) : 5;
    const uint64_t N = 1ull << n;
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    TabRef tab[T], work[T];
    for (int t = 0; t < T; ++t) {
        void* p;
        CK(cudaMalloc(&p, N * 32));
        tab[t] = TabRef{(uint4*)p, N};
        CK(cudaMalloc(&p, N * 16));
        work[t] = TabRef{(uint4*)p, N / 2};
        k_generate<FT><<<sms * 8, BLOCK>>>(tab[t], N, 0xB2000002ull + 3, (uint64_t)t, 0, 1);
    }
    CK(cudaDeviceSynchronize());
    Fe *partials, *result;
    unsigned int* ticket;
    CK(cudaMalloc(&partials, sizeof(Fe) * MAXPTS * sms * 4));
    CK(cudaMalloc(&result, sizeof(Fe) * MAXPTS));
    CK(cudaMalloc(&ticket, 4));
    CK(cudaMemset(ticket, 0, 4));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    uint64_t h = 1469598103934665603ull;

    // ---- round 0
    {
        ScArgs a;
        memset(&a, 0, sizeof a);
        for (int t = 0; t < T; ++t) a.in[t] = tab[t];
        a.n_tables = T;
        a.n_products = P;
        a.n_out = N;
        a.fin.partials = partials;
        a.fin.ticket = ticket;
        a.fin.result = result;
#if defined(KB_TC) && KB_D == 2
        auto kern = k_sc_eval_tc<FT, NPTS>;
        constexpr int ESM = TCG_SMEM;
#elif defined(KB_TC) && KB_D == 3
        auto kern = k_sc_eval_gram<FT, D, NPTS>;
        constexpr int ESM = TcGramEvalSmem<NPTS>::bytes;
#else
        auto kern = k_sc_eval<FT, KIND_PROD, D, NPTS>;
        constexpr int ESM = STAGE_BYTES;
#endif
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ESM));
        int occ = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, BLOCK, ESM));
        if (getenv("KB_OCC_EVAL")) occ = atoi(getenv("KB_OCC_EVAL"));
        const int grid = sms * occ;
        float best = 1e30f;
        for (int r = 0; r < reps + 1; ++r) {
            CK(cudaEventRecord(e0));
            kern<<<grid, BLOCK, ESM>>>(a);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (r > 0 && ms < best) best = ms;
        }
        Fe res[MAXPTS];
        CK(cudaMemcpy(res, result, sizeof(Fe) * NPTS, cudaMemcpyDeviceToHost));
        h = fnv(h, res, sizeof(Fe) * NPTS);
        const double bytes = 32.0 * T * (double)N;
        printf("k_sc_eval      D=%d P=%d n=%d occ=%d: %.3f ms  %.1f GB/s\n", D, P, n, occ, best, bytes / best / 1e6);
    }
    // ---- the largest fold round (tables 2^n -> 2^(n-1)); source tables are not modified
    {
        ScArgs a;
        memset(&a, 0, sizeof a);
        for (int t = 0; t < T; ++t) {
            a.in[t] = tab[t];
            a.out[t] = work[t];
        }
        a.n_tables = T;
        a.n_products = P;
        a.n_out = N / 2;
        a.fin.partials = partials;
        a.fin.ticket = ticket;
        a.fin.result = result;
        // fixed-multiplicand table of an arbitrary challenge r (fr.cuh FixedMul): t[i] = mul(r, 2^(32 i + 64) mod p)
        Fe r = Field<FT>::r2();
        r.l[0] ^= 0x1234567u;
        r = Field<FT>::mul(r, Field<FT>::r2());
        Fe v = Field<FT>::zero();
        v.l[0] = 1;
        for (int k = 0; k < 64; ++k) v = Field<FT>::add(v, v);
        for (int i = 0; i < 8; ++i) {
            Fe t = Field<FT>::mul(r, v);
            memcpy(a.rt.t[i], t.l, 32);
            for (int k = 0; k < 32; ++k) v = Field<FT>::add(v, v);
        }
#ifdef KB_TC
        ScArgsTc atc;
        atc.s = a;
        tc_fold_mats<FT>(r, &atc.mats);
        auto kern = k_sc_fold_eval_tc<FT, KIND_PROD, D, NPTS>;
        constexpr int SM = TcFoldEvalCfg<D, NPTS>::smem;
#define a atc
#else
        auto kern = k_sc_fold_eval<FT, KIND_PROD, D, NPTS>;
        constexpr int SM = FoldSmem<KIND_PROD, D, NPTS>::bytes;
#endif
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SM));
        int occ = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, BLOCK, SM));
        if (getenv("KB_OCC")) occ = atoi(getenv("KB_OCC"));
        const int grid = sms * occ;
        float best = 1e30f;
        for (int r2 = 0; r2 < reps + 1; ++r2) {
            CK(cudaEventRecord(e0));
            kern<<<grid, BLOCK, SM>>>(a);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (r2 > 0 && ms < best) best = ms;
        }
        Fe res[MAXPTS];
        CK(cudaMemcpy(res, result, sizeof(Fe) * (NPTS - 1), cudaMemcpyDeviceToHost));
        h = fnv(h, res, sizeof(Fe) * (NPTS - 1));
        // hash a slice of every folded table
        std::vector<unsigned char> buf(1 << 20);
        for (int t = 0; t < T; ++t) {
            CK(cudaMemcpy(buf.data(), work[t].base + (N / 4 > 1000 ? (N / 4) - 1000 : 0), buf.size() < N * 8 ? buf.size() : N * 8, cudaMemcpyDeviceToHost));
            h = fnv(h, buf.data(), buf.size());
            CK(cudaMemcpy(buf.data(), work[t].base + work[t].stride, buf.size() < N * 8 ? buf.size() : N * 8, cudaMemcpyDeviceToHost));
            h = fnv(h, buf.data(), buf.size());
        }
        const double bytes = 96.0 * T * (double)(N / 2);
        printf("k_sc_fold_eval D=%d P=%d n=%d occ=%d smem=%d: %.3f ms  %.1f GB/s\n", D, P, n, occ, SM, best, bytes / best / 1e6);
    }
    printf("hash %016llx\n", (unsigned long long)h);
    return 0;
}
