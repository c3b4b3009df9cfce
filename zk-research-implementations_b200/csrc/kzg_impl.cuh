// kzg_impl.cuh -- the input-layer commitment of the GKR prover on the device: multilinear KZG over BLS12-381 G1,
// prover side of pcs/src/kzg_pcs/kzg.rs (setup of the G1 Lagrange basis :36-49,183-212; commit :51-53,131-144;
// open :55-57; get_proof :59-95), which gkr/src/gkr_protocol.rs:92-118 runs on the input MLE.  Included by zkb200.cu
// (it needs the context and the table handles); scalars are BLS12-381 Fr tables in the planar layout of kernels.cuh.
//
// What changes against the reference's loops, with identical results (group elements, compared as affine points):
//  * g1 * scalar for every Lagrange scalar (kzg.rs:209-212): fixed-base windows -- 32 mixed additions from a
//    256 x 32 table of multiples of the generator instead of 255 doublings + additions per scalar;
//  * sum_i basis[i] * poly[i] (kzg.rs:131-144, one scalar multiplication per entry): a bucket (Pippenger) MSM --
//    signed-free c-bit windows, (window, digit) keys sorted with cub::DeviceRadixSort, one thread per bucket (one CTA
//    for a bucket with many points), per-window running sums, Horner over the windows;
//  * the quotient of variable k (kzg.rs:152-163) is blown up to the full length by tiling (:165-171) and multiplied
//    with the whole basis; tiling means basis entries with equal low index bits share a scalar, so the same point is
//    the MSM of the 2^(n-k-1)-entry quotient against the FOLDED basis  fold_k[j] = sum_i basis[i * 2^(n-k-1) + j],
//    made once per setup (n additions of halves).  Total MSM work per opening: N instead of n * N;
//  * poly - opened_value (:65-71) is never formed: quotients f(1,.) - f(0,.) and the remainders' quotients do not
//    depend on a constant shift.
// The G2 side of the setup and KZG::verify (:97-129, pairings) are verifier-only and not part of this engine.
#pragma once
#include <cub/device/device_radix_sort.cuh>

#include "g1.cuh"

namespace zkb {
typedef Bls12381Fr KF;  // the scalar field of the commitment

// L[i] = eq(taus, i) from the split eq tables (k_eq_split): the Lagrange scalars of kzg.rs:183-207.
__global__ void __launch_bounds__(BLOCK) k_kzg_eq_full(TabRef hi, TabRef lo, int n_lo, TabRef out, uint64_t n) {
    const uint64_t step = (uint64_t)gridDim.x * BLOCK;
    for (uint64_t i = (uint64_t)blockIdx.x * BLOCK + threadIdx.x; i < n; i += step) st_fe(out, i, eq_lookup<KF>(hi, lo, n_lo, i));
}
// q[j] = f[j + half] - f[j]  (get_quotient, kzg.rs:152-163: f(1, .) - f(0, .))
__global__ void __launch_bounds__(BLOCK) k_kzg_quotient(TabRef f, TabRef q, uint64_t half) {
    const uint64_t step = (uint64_t)gridDim.x * BLOCK;
    for (uint64_t j = (uint64_t)blockIdx.x * BLOCK + threadIdx.x; j < half; j += step)
        st_fe(q, j, Field<KF>::sub(ld_fe(f, j + half), ld_fe(f, j)));
}

// T[w][d] = d * 2^(8 w) * G, affine, w < 32, d < 256 (d = 0: infinity).  One thread per entry.
__global__ void __launch_bounds__(128) k_g1_window_table(G1Affine* table) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= 32 * 256) return;
    const int w = idx >> 8, d = idx & 255;
    G1Jac p = G1::from_affine(G1::generator());
    for (int k = 0; k < 8 * w; ++k) p = G1::dbl(p);
    table[idx] = G1::to_affine(G1::mul_small(p, (uint32_t)d));
}
// out[i] = scalar[i] * G through the window table (kzg.rs:209-212); scalars: Montgomery Fr table.
__global__ void __launch_bounds__(128) k_g1_fixed_base(TabRef scalars, const G1Affine* __restrict__ table, G1Jac* out, uint64_t n) {
    const uint64_t step = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
        const Fe s = Field<KF>::from_mont(ld_fe(scalars, i));
        G1Jac acc = G1::inf();
#pragma unroll 1
        for (int w = 0; w < 32; ++w) {
            const uint32_t d = (s.l[w >> 2] >> (8 * (w & 3))) & 255u;
            if (d) acc = G1::madd(acc, table[w * 256 + d]);
        }
        out[i] = acc;
    }
}
// Jacobian -> affine with one inversion per KZG_BATCH points (Montgomery's trick inside a thread).
constexpr int KZG_BATCH = 8;
__global__ void __launch_bounds__(128) k_g1_batch_affine(const G1Jac* __restrict__ in, G1Affine* out, uint64_t n) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t first = t * KZG_BATCH;
    if (first >= n) return;
    const int m = (int)(n - first < KZG_BATCH ? n - first : KZG_BATCH);
    Fq pre[KZG_BATCH];
    Fq acc = Fqf::one();
    for (int k = 0; k < m; ++k) {
        pre[k] = acc;
        const Fq z = in[first + k].z;
        if (!Fqf::is_zero(z)) acc = Fqf::mul(acc, z);
    }
    Fq inv = Fqf::inv(acc);
    for (int k = m - 1; k >= 0; --k) {
        const G1Jac p = in[first + k];
        G1Affine a;
        if (Fqf::is_zero(p.z)) {
            a.x = Fqf::zero();
            a.y = Fqf::zero();
        } else {
            const Fq zi = Fqf::mul(inv, pre[k]);
            inv = Fqf::mul(inv, p.z);
            const Fq zi2 = Fqf::sqr(zi);
            a.x = Fqf::mul(p.x, zi2);
            a.y = Fqf::mul(p.y, Fqf::mul(zi2, zi));
        }
        out[first + k] = a;
    }
}
// next[j] = cur[j] + cur[j + half]  (the folded basis of the next variable)
__global__ void __launch_bounds__(128) k_g1_fold_basis(const G1Affine* __restrict__ cur, G1Jac* next, uint64_t half) {
    const uint64_t step = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < half; j += step)
        next[j] = G1::madd(G1::from_affine(cur[j]), cur[j + half]);
}

// ---------------------------------------------------------------------------------------- bucket MSM
struct MsmPlan {
    uint32_t c;        // window bits
    uint32_t windows;  // ceil(255 / c)
    uint32_t chunks;   // threads per window in the running-sum reduction
};
// keys[w * n + i] = (w << c) | digit_w(scalar_i), vals = i
__global__ void __launch_bounds__(BLOCK) k_msm_digits(TabRef scalars, uint64_t n, MsmPlan pl, uint32_t* keys, uint32_t* vals) {
    const uint64_t step = (uint64_t)gridDim.x * BLOCK;
    for (uint64_t i = (uint64_t)blockIdx.x * BLOCK + threadIdx.x; i < n; i += step) {
        const Fe s = Field<KF>::from_mont(ld_fe(scalars, i));
        for (uint32_t w = 0; w < pl.windows; ++w) {
            const uint32_t bit = w * pl.c, limb = bit >> 5, sh = bit & 31;
            uint64_t v = s.l[limb];
            if (limb + 1 < 8) v |= (uint64_t)s.l[limb + 1] << 32;
            const uint32_t d = (uint32_t)(v >> sh) & ((1u << pl.c) - 1u);
            keys[(uint64_t)w * n + i] = (w << pl.c) | d;
            vals[(uint64_t)w * n + i] = (uint32_t)i;
        }
    }
}
// first / one-past-last position of every key in the sorted array; buckets above `heavy_min` points are listed
__global__ void __launch_bounds__(BLOCK) k_msm_bounds(const uint32_t* __restrict__ keys, uint64_t m, uint32_t* start, uint32_t* end) {
    const uint64_t step = (uint64_t)gridDim.x * BLOCK;
    for (uint64_t j = (uint64_t)blockIdx.x * BLOCK + threadIdx.x; j < m; j += step) {
        const uint32_t k = keys[j];
        if (j == 0 || keys[j - 1] != k) start[k] = (uint32_t)j;
        if (j + 1 == m || keys[j + 1] != k) end[k] = (uint32_t)(j + 1);
    }
}
constexpr uint32_t MSM_HEAVY = 512;  // buckets with more points get a whole CTA
__global__ void __launch_bounds__(128) k_msm_buckets(const uint32_t* __restrict__ vals, const uint32_t* __restrict__ start,
                                                     const uint32_t* __restrict__ end, const G1Affine* __restrict__ bases, MsmPlan pl,
                                                     G1Jac* buckets, uint32_t* heavy_list, uint32_t* heavy_count) {
    const uint32_t nb = pl.windows << pl.c;
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    G1Jac acc = G1::inf();
    if ((b & ((1u << pl.c) - 1u)) != 0) {  // digit 0 contributes nothing
        const uint32_t s = start[b], e = end[b];
        if (e - s > MSM_HEAVY) {
            heavy_list[atomicAdd(heavy_count, 1u)] = b;
        } else {
            for (uint32_t j = s; j < e; ++j) acc = G1::madd(acc, bases[vals[j]]);
        }
    }
    buckets[b] = acc;
}
__global__ void __launch_bounds__(256) k_msm_heavy(const uint32_t* __restrict__ vals, const uint32_t* __restrict__ start,
                                                   const uint32_t* __restrict__ end, const G1Affine* __restrict__ bases, G1Jac* buckets,
                                                   const uint32_t* __restrict__ heavy_list, const uint32_t* __restrict__ heavy_count) {
    extern __shared__ unsigned char kz_smem[];
    G1Jac* red = reinterpret_cast<G1Jac*>(kz_smem);
    const uint32_t nh = *heavy_count;
    for (uint32_t h = blockIdx.x; h < nh; h += gridDim.x) {
        const uint32_t b = heavy_list[h];
        G1Jac acc = G1::inf();
        for (uint32_t j = start[b] + threadIdx.x; j < end[b]; j += blockDim.x) acc = G1::madd(acc, bases[vals[j]]);
        red[threadIdx.x] = acc;
        __syncthreads();
        for (uint32_t off = blockDim.x / 2; off > 0; off >>= 1) {
            if (threadIdx.x < off) red[threadIdx.x] = G1::add(red[threadIdx.x], red[threadIdx.x + off]);
            __syncthreads();
        }
        if (threadIdx.x == 0) buckets[b] = red[0];
        __syncthreads();
    }
}
// Per window and chunk of digits [lo, hi): L = sum_d (d - lo + 1) B_d and R = sum_d B_d by the running-sum trick,
// then part = L + (lo - 1) R = sum_d d * B_d over the chunk.
__global__ void __launch_bounds__(128) k_msm_window_chunks(const G1Jac* __restrict__ buckets, MsmPlan pl, G1Jac* parts) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= pl.windows * pl.chunks) return;
    const uint32_t w = t / pl.chunks, ch = t % pl.chunks;
    const uint32_t per = ((1u << pl.c) + pl.chunks - 1) / pl.chunks;
    uint32_t lo = ch * per, hi = lo + per;
    if (hi > (1u << pl.c)) hi = 1u << pl.c;
    if (lo == 0) lo = 1;
    G1Jac run = G1::inf(), L = G1::inf();
    for (uint32_t d = hi; d-- > lo;) {
        run = G1::add(run, buckets[(w << pl.c) + d]);
        L = G1::add(L, run);
    }
    if (lo < hi && lo > 1) L = G1::add(L, G1::mul_small(run, lo - 1));
    parts[t] = L;
}
// S_w = sum of the window's chunk results: one CTA per window, shared-memory tree.
__global__ void __launch_bounds__(128) k_msm_window_sum(const G1Jac* __restrict__ parts, MsmPlan pl, G1Jac* wsum) {
    __shared__ G1Jac red[128];
    const uint32_t w = blockIdx.x;
    G1Jac s = G1::inf();
    for (uint32_t ch = threadIdx.x; ch < pl.chunks; ch += blockDim.x) s = G1::add(s, parts[w * pl.chunks + ch]);
    red[threadIdx.x] = s;
    __syncthreads();
    for (uint32_t off = blockDim.x / 2; off > 0; off >>= 1) {
        if (threadIdx.x < off) red[threadIdx.x] = G1::add(red[threadIdx.x], red[threadIdx.x + off]);
        __syncthreads();
    }
    if (threadIdx.x == 0) wsum[w] = red[0];
}
// Horner over the windows: result = sum_w 2^(c w) S_w.  Affine output (canonical, 2 x 48 bytes little-endian;
// infinity = zeros).
__global__ void k_msm_horner(const G1Jac* __restrict__ wsum, MsmPlan pl, uint8_t* result_bytes) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    G1Jac acc = G1::inf();
    for (uint32_t ww = pl.windows; ww-- > 0;) {
        for (uint32_t k = 0; k < pl.c; ++k) acc = G1::dbl(acc);
        acc = G1::add(acc, wsum[ww]);
    }
    G1Affine a = G1::to_affine(acc);
    if (!G1::is_inf(a)) {
        a.x = Fqf::from_mont(a.x);
        a.y = Fqf::from_mont(a.y);
    }
    memcpy(result_bytes, a.x.l, 48);
    memcpy(result_bytes + 48, a.y.l, 48);
}
// Small problems: every thread multiplies its own pair (4-bit windows), a tree over global memory adds them up.
__global__ void __launch_bounds__(256) k_msm_small(TabRef scalars, const G1Affine* __restrict__ bases, uint32_t n, G1Jac* scratch, uint8_t* result_bytes) {
    const uint32_t i = threadIdx.x;
    G1Jac acc = G1::inf();
    for (uint32_t k = i; k < n; k += blockDim.x) {
        const Fe s = Field<KF>::from_mont(ld_fe(scalars, k));
        const G1Affine b = bases[k];
        G1Jac tab[16];
        tab[0] = G1::inf();
        for (int d = 1; d < 16; ++d) tab[d] = G1::madd(tab[d - 1], b);
        G1Jac p = G1::inf();
        for (int nib = 63; nib >= 0; --nib) {
            for (int q = 0; q < 4; ++q) p = G1::dbl(p);
            const uint32_t d = (s.l[nib >> 3] >> (4 * (nib & 7))) & 15u;
            if (d) p = G1::add(p, tab[d]);
        }
        acc = G1::add(acc, p);
    }
    scratch[i] = acc;
    __syncthreads();
    for (uint32_t off = blockDim.x / 2; off > 0; off >>= 1) {
        if (i < off) scratch[i] = G1::add(scratch[i], scratch[i + off]);
        __syncthreads();
    }
    if (i == 0) {
        G1Affine a = G1::to_affine(scratch[0]);
        if (!G1::is_inf(a)) {
            a.x = Fqf::from_mont(a.x);
            a.y = Fqf::from_mont(a.y);
        }
        memcpy(result_bytes, a.x.l, 48);
        memcpy(result_bytes + 48, a.y.l, 48);
    }
}
// canonical affine bytes of stored basis points (tests / export)
__global__ void k_g1_export(const G1Affine* __restrict__ pts, uint64_t n, uint8_t* out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    G1Affine a = pts[i];
    if (!G1::is_inf(a)) {
        a.x = Fqf::from_mont(a.x);
        a.y = Fqf::from_mont(a.y);
    }
    memcpy(out + 96 * i, a.x.l, 48);
    memcpy(out + 96 * i + 48, a.y.l, 48);
}
}  // namespace zkb
