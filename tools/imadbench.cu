// imadbench.cu -- multiplier-pipe microbenchmarks behind the IMAD roofline: issue rate and
// dependent-issue latency of the 32x32->64 multiply-accumulate forms the field arithmetic is built from.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/imadbench.cu -o build/kb/imadbench
// Every mode runs CHAINS independent carry chains of LEN multiply-accumulates per thread and iteration, at full
// occupancy (throughput) and with one warp per scheduler (latency of the chain's dependent issue).
#include <cstdio>
#include <cstdlib>
#include <cstdint>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1);} } while (0)

// MODE 0: mad.wide (no carry)            1: mul.wide + add.cc.u64 (carry out only, every op)
// MODE 2: chain cc, c.cc x (LEN-1)        3: mad.lo.cc / madc.hi.cc pairs (32-bit halves, one chain)
// MODE 4: mad.wide + separate 64-bit carry chain through IADD3 (add.cc.u32 / addc.u32 on the halves)
template <int MODE, int CHAINS, int LEN>
__global__ void __launch_bounds__(256) k(uint64_t* out, uint32_t iters, uint32_t a, uint32_t b) {
    uint64_t acc[CHAINS][LEN];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c)
#pragma unroll
        for (int k = 0; k < LEN; ++k) acc[c][k] = threadIdx.x + c * 17 + k;
    uint32_t x = a + threadIdx.x, y = b;
    for (uint32_t it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (MODE == 2 && CHAINS == 2) {
                // two chains interleaved instruction by instruction is impossible with one CC flag in PTX; leave the
                // order to ptxas by writing each chain as ONE asm block
#pragma unroll
                for (int c = 0; c < CHAINS; ++c) {
                    static_assert(LEN == 4 || MODE != 2 || CHAINS != 2, "");
                    asm("{\n\t.reg .u64 p;\n\t"
                        "mul.wide.u32 p, %4, %5;\n\tadd.cc.u64 %0, %0, p;\n\t"
                        "mul.wide.u32 p, %4, %5;\n\taddc.cc.u64 %1, %1, p;\n\t"
                        "mul.wide.u32 p, %4, %5;\n\taddc.cc.u64 %2, %2, p;\n\t"
                        "mul.wide.u32 p, %4, %5;\n\taddc.u64 %3, %3, p;\n\t}"
                        : "+l"(acc[c][0]), "+l"(acc[c][1]), "+l"(acc[c][2]), "+l"(acc[c][3]) : "r"(x), "r"(y));
                }
                continue;
            }
#pragma unroll
            for (int c = 0; c < CHAINS; ++c) {
#pragma unroll
                for (int k = 0; k < LEN; ++k) {
                    if (MODE == 0) {
                        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[c][k]) : "r"(x), "r"(y));
                    } else if (MODE == 1) {
                        asm volatile("{\n\t.reg .u64 p;\n\tmul.wide.u32 p, %1, %2;\n\tadd.cc.u64 %0, %0, p;\n\t}" : "+l"(acc[c][k]) : "r"(x), "r"(y));
                    } else if (MODE == 2) {
                        if (k == 0) asm volatile("{\n\t.reg .u64 p;\n\tmul.wide.u32 p, %1, %2;\n\tadd.cc.u64 %0, %0, p;\n\t}" : "+l"(acc[c][k]) : "r"(x), "r"(y));
                        else asm volatile("{\n\t.reg .u64 p;\n\tmul.wide.u32 p, %1, %2;\n\taddc.cc.u64 %0, %0, p;\n\t}" : "+l"(acc[c][k]) : "r"(x), "r"(y));
                    } else if (MODE == 3) {
                        uint32_t lo = (uint32_t)acc[c][k], hi = (uint32_t)(acc[c][k] >> 32);
                        if (k == 0) asm volatile("mad.lo.cc.u32 %0, %1, %2, %0;" : "+r"(lo) : "r"(x), "r"(y));
                        else asm volatile("madc.lo.cc.u32 %0, %1, %2, %0;" : "+r"(lo) : "r"(x), "r"(y));
                        asm volatile("madc.hi.cc.u32 %0, %1, %2, %0;" : "+r"(hi) : "r"(x), "r"(y));
                        acc[c][k] = ((uint64_t)hi << 32) | lo;
                    } else if (MODE == 4) {
                        // product without carry, carries collected by 32-bit adds on the ALU pipe
                        uint64_t p;
                        asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(x), "r"(y));
                        uint32_t lo = (uint32_t)acc[c][k], hi = (uint32_t)(acc[c][k] >> 32);
                        if (k == 0) asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(lo) : "r"((uint32_t)p));
                        else asm volatile("addc.cc.u32 %0, %0, %1;" : "+r"(lo) : "r"((uint32_t)p));
                        asm volatile("addc.cc.u32 %0, %0, %1;" : "+r"(hi) : "r"((uint32_t)(p >> 32)));
                        acc[c][k] = ((uint64_t)hi << 32) | lo;
                    }
                }
            }
        }
    }
    uint64_t s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c)
#pragma unroll
        for (int k = 0; k < LEN; ++k) s += acc[c][k];
    if (s == 0x123456789abcdef0ull) out[blockIdx.x * 256 + threadIdx.x] = s;
}

template <int MODE, int CHAINS, int LEN>
static void run(const char* name, int sms, uint64_t* buf) {
    const uint32_t iters = 4096;
    for (int occ = 0; occ < 3; ++occ) {
        // occ 0: 8 CTAs x 256 threads per SM (16 warps per scheduler); 1: one CTA of 128 threads (1 warp per scheduler);
        // 2: one CTA of 512 threads (4 warps per scheduler, the round kernels' occupancy)
        const int grid = occ == 0 ? sms * 8 : sms, block = occ == 0 ? 256 : (occ == 1 ? 128 : 512);
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0));
        CK(cudaEventCreate(&e1));
        k<MODE, CHAINS, LEN><<<grid, block>>>(buf, 64, 3u, 5u);
        CK(cudaEventRecord(e0));
        k<MODE, CHAINS, LEN><<<grid, block>>>(buf, iters, 3u, 5u);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        const double macs = (double)grid * block * iters * 4.0 * CHAINS * LEN;
        const double per_s = macs / (ms * 1e-3);
        // cycles per warp-instruction per scheduler at 1.965 GHz: (4 schedulers x sms x clk) / (macs/32 per second)
        const double cyc = 4.0 * sms * 1.965e9 / (per_s / 32.0);
        printf("%-44s warps/sched %2d: %8.3f T MAC/s  %.2f sched-cycles per warp-MAC\n", name, occ == 0 ? 16 : (occ == 1 ? 1 : 4), per_s / 1e12, cyc);
    }
}

int main() {
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    uint64_t* buf;
    CK(cudaMalloc(&buf, 8ull * 256 * sms * 8));
    run<0, 2, 4>("mad.wide no carry, 8 accumulators", sms, buf);
    run<1, 2, 4>("mul.wide+add.cc (carry out only)", sms, buf);
    run<2, 1, 4>("chain of 4 (cc, c.cc x3), 1 chain", sms, buf);
    run<2, 1, 8>("chain of 8, 1 chain", sms, buf);
    run<2, 2, 4>("2 chains of 4, one asm block each", sms, buf);
    run<3, 1, 4>("mad.lo.cc/madc.hi.cc chain of 4 pairs", sms, buf);
    run<4, 1, 4>("mul.wide + 32-bit add carry chain", sms, buf);
    return 0;
}
