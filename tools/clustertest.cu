// clustertest.cu -- which distributed-shared-memory primitive works how on this device: generic mapa + store, remote
// atomicAdd, spin on the local counter, with 200 KiB of dynamic shared memory per CTA and clusters of 2..16 CTAs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/clustertest.cu -o build/kb/clustertest
#include <cstdio>
#include <cstdint>
#include <ctime>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t cta_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void csync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
template <class T>
__device__ __forceinline__ T* cmap(T* p, uint32_t rank) {
    uint64_t out;
    asm volatile("mapa.u64 %0, %1, %2;" : "=l"(out) : "l"((uint64_t)(uintptr_t)p), "r"(rank));
    return (T*)(uintptr_t)out;
}

// mode 0: sync only; 1: remote plain store + final sync; 2: remote store + fence + remote atomic, CTA 0 spins; 3: as 2, others exit at once
__global__ void k_test(int mode, int nc, unsigned int* out) {
    extern __shared__ uint4 dyn[];
    __shared__ unsigned int s_arr;
    __shared__ uint4 s_slot[16];
    const uint32_t r = cta_rank();
    if (threadIdx.x == 0) s_arr = 0;
    dyn[threadIdx.x] = make_uint4(r, 0, 0, 0);
    __syncthreads();
    csync();
    if (mode == 0) {
        if (threadIdx.x == 0) out[r] = 100 + r;
        return;
    }
    if (mode == 1) {
        if (threadIdx.x == 0 && r != 0) *cmap(&s_slot[r], 0) = make_uint4(r, r, r, r);
        csync();
        if (r == 0 && threadIdx.x < nc) out[threadIdx.x] = threadIdx.x == 0 ? 100 : 100 + s_slot[threadIdx.x].x;
        return;
    }
    if (r != 0) {
        if (threadIdx.x == 0) {
            *cmap(&s_slot[r], 0) = make_uint4(r, r, r, r);
            cmap(dyn, 0)[1024 + r] = make_uint4(7 * r, 0, 0, 0);
            __threadfence();
            atomicAdd(cmap(&s_arr, 0), 1u);
        }
        if (mode == 2) csync();
        return;
    }
    if (threadIdx.x == 0) {
        const long long t0 = clock64();
        while (*(volatile unsigned int*)&s_arr < (unsigned)(nc - 1))
            if (clock64() - t0 > 2000000000ll) break;
        __threadfence();
        out[16] = s_arr;
    }
    __syncthreads();
    if (threadIdx.x < nc) out[threadIdx.x] = threadIdx.x == 0 ? 100 : 100 + s_slot[threadIdx.x].x + dyn[1024 + threadIdx.x].x;
    if (mode == 2) csync();
}

__device__ __forceinline__ void cfence() { asm volatile("fence.acq_rel.cluster;" ::: "memory"); }
// ping-pong between CTA 0 and CTA 1 of a cluster through distributed shared memory, `iters` round trips, cycles per trip:
//  variant 0: volatile remote store of a sequence number / volatile local spin, no fence
//  variant 1: + payload (32 B remote store) and fence.acq_rel.cluster on both sides
//  variant 2: as 1 with __threadfence() instead
//  variant 3: payload + fence + remote atomicAdd as the flag (CTA 1 -> CTA 0), remote store back
__global__ void k_pingpong(int variant, int iters, long long* out) {
    __shared__ unsigned int s_flag;
    __shared__ uint4 s_pay[2];
    const uint32_t r = cta_rank();
    if (threadIdx.x == 0) s_flag = 0;
    __syncthreads();
    csync();
    if (threadIdx.x == 0 && r < 2) {
        unsigned int* peer_flag = cmap(&s_flag, r ^ 1);
        uint4* peer_pay = cmap(&s_pay[0], r ^ 1);
        const long long t0 = clock64();
        for (int i = 1; i <= iters; ++i) {
            if (r == 0) {
                if (variant >= 1) { peer_pay[0] = make_uint4(i, i, i, i); peer_pay[1] = make_uint4(i, i, i, i); }
                if (variant == 1 || variant == 3) cfence();
                if (variant == 2) __threadfence();
                *(volatile unsigned int*)peer_flag = (unsigned)i;
                while (*(volatile unsigned int*)&s_flag < (unsigned)i) {}
                if (variant == 1 || variant == 3) cfence();
                if (variant == 2) __threadfence();
            } else {
                while (*(volatile unsigned int*)&s_flag < (unsigned)i) {}
                if (variant == 1 || variant == 3) cfence();
                if (variant == 2) __threadfence();
                if (variant >= 1) { peer_pay[0] = s_pay[0]; peer_pay[1] = s_pay[1]; }
                if (variant == 1 || variant == 3) cfence();
                if (variant == 2) __threadfence();
                if (variant == 3) atomicAdd(peer_flag, 1u);
                else *(volatile unsigned int*)peer_flag = (unsigned)i;
            }
        }
        if (r == 0) out[variant] = (clock64() - t0) / iters;
    }
    __syncthreads();
    csync();
}

int main() {
    unsigned int* d;
    cudaMalloc(&d, 32 * 4);
    const int smem = 200 * 1024;
    cudaFuncSetAttribute(k_test, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k_test, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    for (int nc = 2; nc <= 16; nc *= 2)
        for (int mode = 0; mode < 4; ++mode) {
            cudaMemset(d, 0, 32 * 4);
            cudaLaunchConfig_t cfg = {};
            cudaLaunchAttribute attr[1];
            cfg.gridDim = dim3(nc);
            cfg.blockDim = dim3(512);
            cfg.dynamicSmemBytes = smem;
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = nc;
            attr[0].val.clusterDim.y = attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            int ncl = -1;
            cudaError_t eo = cudaOccupancyMaxActiveClusters(&ncl, k_test, &cfg);
            cudaError_t e = cudaLaunchKernelEx(&cfg, k_test, mode, nc, d);
            cudaError_t e2 = cudaDeviceSynchronize();
            unsigned int h[32] = {0};
            cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
            printf("nc=%2d mode=%d: occupancy %s clusters=%d, launch %s, sync %s; out:", nc, mode, cudaGetErrorName(eo), ncl, cudaGetErrorName(e), cudaGetErrorName(e2));
            for (int i = 0; i < nc; ++i) printf(" %u", h[i]);
            printf(" arr=%u\n", h[16]);
            if (e2 != cudaSuccess) return 1;
        }
    {
        long long* dl;
        cudaMalloc(&dl, 8 * 8);
        cudaMemset(dl, 0, 64);
        for (int v = 0; v < 4; ++v) {
            cudaLaunchConfig_t cfg = {};
            cudaLaunchAttribute attr[1];
            cfg.gridDim = dim3(2);
            cfg.blockDim = dim3(512);
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = 2;
            attr[0].val.clusterDim.y = attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            cudaLaunchKernelEx(&cfg, k_pingpong, v, 2000, dl);
            cudaError_t e = cudaDeviceSynchronize();
            long long h[8];
            cudaMemcpy(h, dl, 64, cudaMemcpyDeviceToHost);
            printf("DSMEM ping-pong variant %d: %s, %lld cycles per round trip\n", v, cudaGetErrorName(e), h[v]);
        }
    }
    // launch cost: an empty kernel (mode 0) as a plain launch and as a cluster, launch call and launch + completion
    for (int smem_kb : {8, 200})
        for (int nc : {1, 2, 4, 16}) {
            cudaLaunchConfig_t cfg = {};
            cudaLaunchAttribute attr[1];
            cfg.gridDim = dim3(nc);
            cfg.blockDim = dim3(512);
            cfg.dynamicSmemBytes = smem_kb * 1024;
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = nc;
            attr[0].val.clusterDim.y = attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr;
            cfg.numAttrs = nc > 1 ? 1 : 0;
            double t_call = 0, t_all = 0;
            const int reps = 200;
            for (int i = 0; i < reps + 10; ++i) {
                timespec a, b, c2;
                clock_gettime(CLOCK_MONOTONIC, &a);
                cudaLaunchKernelEx(&cfg, k_test, 0, nc, d);
                clock_gettime(CLOCK_MONOTONIC, &b);
                cudaDeviceSynchronize();
                clock_gettime(CLOCK_MONOTONIC, &c2);
                if (i >= 10) {
                    t_call += (b.tv_sec - a.tv_sec) * 1e6 + (b.tv_nsec - a.tv_nsec) * 1e-3;
                    t_all += (c2.tv_sec - a.tv_sec) * 1e6 + (c2.tv_nsec - a.tv_nsec) * 1e-3;
                }
            }
            printf("empty kernel, %3d KiB smem, %2d CTA%s: launch call %.2f us, launch + completion %.2f us\n", smem_kb, nc,
                   nc > 1 ? "s (cluster)" : "", t_call / reps, t_all / reps);
        }
    return 0;
}
