// kernels.cuh -- the table kernels of the sumcheck / GKR hot path (sm_100a).
//
// Data layout in HBM ("planar", limb-interleaved): a table of N field elements
// is two planes of N uint4 each -- plane 0 holds limbs 0..3 (the low 128 bits)
// of every element, plane 1 holds limbs 4..7.  Consecutive threads read
// consecutive uint4 of one plane, so every 128-bit load of a warp is one
// contiguous 512-byte request.  Values are Montgomery residues (R = 2^256),
// i.e. exactly ark-ff's in-memory representation re-ordered.
//
// Index convention of the reference (multilinear_polynomial_evaluation.rs:39-50,
// :158-164): variable 0 is the MOST significant index bit, so binding variable 0
// pairs entry i with entry i + N/2.
//
// Kernel inventory (SURVEY.md section 2.2 numbering):
//   K1  k_fold            partial_evaluate, any `bit`
//   K2/K7  k_sc_eval      round-0 evaluations  (plain sumcheck = PROD<D=1>)
//   K8  k_sc_fold_eval    fused: bind the previous challenge in all tables,
//                         write the folded tables once, and accumulate the next
//                         round's evaluations from the freshly folded values
//   K9  (in-kernel)       last-block-done second-stage reduction
//   K3  k_fold_tables     fold a list of (small) tables; final bind
//   K4  k_vec_op / k_scale, K5 k_tensor, K6 k_aos_to_planar / k_planar_to_aos
//   K10 k_layer_eval, K11 k_eq_split, K12 k_gkr_phase1, K13 k_gkr_phase2,
//   K14 = k_sc_* with KIND_XYZ.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

#include "fr.cuh"

namespace zkb {

constexpr int MAXT = 16;      // tables per composed polynomial
constexpr int MAXPTS = 5;     // evaluation points per round (degree <= 4)
#ifndef ZKB_BLOCK
#define ZKB_BLOCK 256
#endif
#ifndef ZKB_MINB
#define ZKB_MINB 2
#endif
constexpr int BLOCK = ZKB_BLOCK;    // threads per CTA for every table kernel (tools/kbench.cu may override both)
constexpr int MAXCHAL = 40;   // challenges per eq table

struct TabRef {
    uint4* base;       // plane 0; plane 1 starts at base + stride
    uint64_t stride;   // in uint4 units
};

// Where a reducing kernel leaves its NPTS sums.
struct FinishArgs {
    Fe* partials;      // [gridDim.x][NPTS] scratch
    unsigned int* ticket;
    Fe* result;        // [NPTS] Montgomery (device or mapped host memory)
    unsigned long long* result_wide;  // optional [NPTS][8] limbs zero-extended to u64 (NCCL sum operand)
    volatile unsigned int* flag;      // optional mailbox flag (mapped host memory): set to `seq` after `result`
    unsigned int seq;
    unsigned long long* stamp;        // optional: %globaltimer right before the result is published
    volatile unsigned int* chk;       // optional (mapped host memory, 2 words): checksum of result + seq.  With it the
                                      // message is published WITHOUT a system fence (1.4 us on B200): the host accepts the
                                      // sums only when the checksum matches what it reads, and re-reads otherwise
};
// Checksum of a device -> host round message: n field elements and the sequence number they belong to.
__host__ __device__ __forceinline__ void msg_checksum(const Fe* v, int n, unsigned int seq, unsigned int* c0, unsigned int* c1) {
    unsigned int x = seq * 0x9E3779B9u, y = seq ^ 0x85EBCA6Bu;
    for (int p = 0; p < n; ++p)
        for (int k = 0; k < 8; ++k) {
            const unsigned int w = v[p].l[k];
            x ^= w;
            y = ((y << 5) | (y >> 27)) + w;
        }
    *c0 = x;
    *c1 = y;
}
__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
template <int NPTS>
__device__ __forceinline__ void publish_result(const Fe* tot, const FinishArgs& a) {
#pragma unroll
    for (int p = 0; p < NPTS; ++p) {
        a.result[p] = tot[p];
        if (a.result_wide) {
#pragma unroll
            for (int k = 0; k < 8; ++k) a.result_wide[p * 8 + k] = tot[p].l[k];
        }
    }
    if (a.stamp) *a.stamp = gtime();
    if (a.chk && a.flag && !a.result_wide) {
        unsigned int c0, c1;
        msg_checksum(tot, NPTS, a.seq, &c0, &c1);
        a.chk[0] = c0;
        a.chk[1] = c1;
        *a.flag = a.seq;
        return;
    }
    __threadfence_system();
    if (a.flag) *a.flag = a.seq;
}

struct ScArgs {
    TabRef in[MAXT];
    TabRef out[MAXT];
    int n_tables;      // tables to fold (product-major: table = p*D + f)
    int n_products;    // P (KIND_PROD)
    uint64_t n_out;    // entries per table AFTER the fold (k_sc_fold_eval) / table size (k_sc_eval)
    FixedMul rt;       // multiplication table of the challenge being bound (k_sc_fold_eval)
    FinishArgs fin;
};

struct ChalList {
    Fe r[MAXCHAL];
};

enum { KIND_PROD = 0, KIND_XYZ = 1 };

// ----------------------------------------------------------- load / store
__device__ __forceinline__ Fe ld_fe(const TabRef& t, uint64_t i) {
    uint4 a = t.base[i];
    uint4 b = t.base[t.stride + i];
    Fe r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
__device__ __forceinline__ void st_fe(const TabRef& t, uint64_t i, const Fe& v) {
    t.base[i] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    t.base[t.stride + i] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}

// ------------------------------------------------------------- reductions
template <class F>
__device__ __forceinline__ Fe warp_sum(Fe v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        Fe o;
#pragma unroll
        for (int k = 0; k < 8; ++k) o.l[k] = __shfl_down_sync(0xffffffffu, v.l[k], off);
        v = Field<F>::add(v, o);
    }
    return v;
}

// Sum NPTS accumulators over the CTA; valid in thread 0.
template <class F, int NPTS>
__device__ __forceinline__ void block_sum(Fe* acc, Fe (*smem)[NPTS]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int p = 0; p < NPTS; ++p) {
        acc[p] = warp_sum<F>(acc[p]);
        if (lane == 0) smem[warp][p] = acc[p];
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int p = 0; p < NPTS; ++p) {
            Fe v = (lane < BLOCK / 32) ? smem[lane][p] : Field<F>::zero();
            acc[p] = warp_sum<F>(v);
        }
    }
    __syncthreads();
}

// Per-CTA partials -> global; the last CTA to arrive reduces them all and
// publishes the round's NPTS evaluations (K9 folded into the producer).
template <class F, int NPTS>
__device__ __forceinline__ void finish_round(Fe* acc, const FinishArgs& a, unsigned int n_active = 0) {
    __shared__ Fe smem[BLOCK / 32][NPTS];
    __shared__ bool is_last;
    if (n_active == 0) n_active = gridDim.x;  // CTAs [0, n_active) take part; the others must not call this
    block_sum<F, NPTS>(acc, smem);
    if (n_active == 1) {  // small tables: the CTA's sums are the round's sums
        if (threadIdx.x == 0) publish_result<NPTS>(acc, a);
        return;
    }
    if (threadIdx.x == 0) {
#pragma unroll
        for (int p = 0; p < NPTS; ++p) a.partials[(uint64_t)blockIdx.x * NPTS + p] = acc[p];
        __threadfence();
        unsigned int t = atomicAdd(a.ticket, 1u);
        is_last = (t == n_active - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    Fe tot[NPTS];
#pragma unroll
    for (int p = 0; p < NPTS; ++p) tot[p] = Field<F>::zero();
    for (unsigned int b = threadIdx.x; b < n_active; b += BLOCK) {
#pragma unroll
        for (int p = 0; p < NPTS; ++p) {
            const Fe* src = a.partials + (uint64_t)b * NPTS + p;
            Fe v;
            const uint4* s4 = reinterpret_cast<const uint4*>(src);
            uint4 x = __ldcg(s4), y = __ldcg(s4 + 1);
            v.l[0] = x.x; v.l[1] = x.y; v.l[2] = x.z; v.l[3] = x.w;
            v.l[4] = y.x; v.l[5] = y.y; v.l[6] = y.z; v.l[7] = y.w;
            tot[p] = Field<F>::add(tot[p], v);
        }
    }
    block_sum<F, NPTS>(tot, smem);
    if (threadIdx.x == 0) {
        *a.ticket = 0u;
        publish_result<NPTS>(tot, a);
    }
}

// --------------------------------------------- round-polynomial evaluation
// The round polynomial s(t) = sum over pair positions of the integrand at
// lo + t*(hi - lo), t = 0..NPTS-1 (get_round_partial_polynomial_proof_gkr,
// sum_check_protocol.rs:152-166, without materialising the (d+1) folded copies).
// KIND_PROD: integrand = sum over products p of prod_f table[p][f]
// KIND_XYZ : integrand = X*Y + Z  (two-phase GKR)
//
// Accumulation is LAZY: the raw 512-bit products of Montgomery residues are
// summed in Wide accumulators and reduced once per thread (fr.cuh mac_wide /
// reduce_wide), which halves the multiplier work of every product.
// SKIP1: s(1) is not computed -- the host derives it from the running claim,
// s(1) = claim - s(0); slot layout is then {s(0), s(2), s(3), ...}.
// Accumulators parked in shared memory (one private slot per thread): the 17-word sums are
// only needed for the 33 additions that end each product, so they need not occupy registers
// during the folds.  Volatile asm keeps the load AFTER the product in program order.
// Slot layout: [acc][5][BLOCK] uint4 (words 17..19 unused).
__device__ __forceinline__ uint4 lds128(const uint4* p) {
    uint4 v;
    const unsigned int sa = static_cast<unsigned int>(__cvta_generic_to_shared(p));
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(sa));
    return v;
}
__device__ __forceinline__ void sts128(uint4* p, const uint4& v) {
    const unsigned int sa = static_cast<unsigned int>(__cvta_generic_to_shared(p));
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(sa), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
constexpr int ACC_VECS = 5;
__device__ __forceinline__ Wide wide_load(const uint4* slot) {
    Wide w;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint4 v = lds128(slot + k * BLOCK);
        w.l[4 * k] = v.x; w.l[4 * k + 1] = v.y; w.l[4 * k + 2] = v.z; w.l[4 * k + 3] = v.w;
    }
    w.l[16] = lds128(slot + 4 * BLOCK).x;
    return w;
}
__device__ __forceinline__ void wide_store(uint4* slot, const Wide& w) {
#pragma unroll
    for (int k = 0; k < 4; ++k) sts128(slot + k * BLOCK, make_uint4(w.l[4 * k], w.l[4 * k + 1], w.l[4 * k + 2], w.l[4 * k + 3]));
    sts128(slot + 4 * BLOCK, make_uint4(w.l[16], 0u, 0u, 0u));
}
template <class F>
__device__ __forceinline__ void mac_wide_sm(uint4* slot, const Fe& a, const Fe& b) {
    uint64_t ev[8], od[8];
    Field<F>::mul_wide16(a, b, ev, od);
    Wide w = wide_load(slot);
    Field<F>::wide_add(w, ev, od);
    wide_store(slot, w);
}

template <int NPTS, bool SKIP1>
struct Slots {
    static constexpr int N = SKIP1 ? NPTS - 1 : NPTS;
    __device__ __forceinline__ static constexpr int of(int t) { return SKIP1 ? (t == 0 ? 0 : t - 1) : t; }
};

template <class F, int D, int NPTS, bool SKIP1, bool SM = false>
struct RoundAcc {
    typedef Field<F> Fd;
    typedef Slots<NPTS, SKIP1> S;
    Wide w[SM ? 1 : S::N];
    uint4* sm;  // SM: this thread's slot 0 (accumulator p at sm + p * ACC_VECS * BLOCK)
    __device__ __forceinline__ void init(uint4* slots = nullptr) {
        sm = slots;
#pragma unroll
        for (int p = 0; p < S::N; ++p) {
            if (SM) wide_store(sm + p * ACC_VECS * BLOCK, Fd::wide_zero());
            else w[p] = Fd::wide_zero();
        }
    }
    __device__ __forceinline__ void mac(int p, const Fe& x, const Fe& y) {
        if (SM) mac_wide_sm<F>(sm + p * ACC_VECS * BLOCK, x, y);
        else Fd::mac_wide(w[p], x, y);
    }
    // one product of D factors at this pair position
    __device__ __forceinline__ void add_product(const Fe* lo, const Fe* hi) {
        if (!SKIP1 && D == 3 && NPTS == 4) {
            // Round 0 of a 3-factor product needs all four points.  q(t) = (lo0 + t d0)(lo1 + t d1) is a quadratic in t:
            // three full products (t = 0, 1, infinity) give it at every point by second differences, instead of four.
            const Fe q0 = Fd::mul(lo[0], lo[1]), q1 = Fd::mul(hi[0], hi[1]);
            const Fe d2 = Fd::sub(hi[2], lo[2]);
            Fe e = Fd::mul(Fd::sub(hi[0], lo[0]), Fd::sub(hi[1], lo[1]));  // q(inf)
            mac(0, q0, lo[2]);
            mac(1, q1, hi[2]);
            e = Fd::dbl(e);                       // the constant second difference 2 q(inf)
            Fe dd = Fd::add(Fd::sub(q1, q0), e);  // q(2) - q(1)
            Fe q = Fd::add(q1, dd), c = Fd::add(hi[2], d2);
            mac(2, q, c);
            dd = Fd::add(dd, e);
            q = Fd::add(q, dd);
            c = Fd::add(c, d2);
            mac(3, q, c);
            return;
        }
        {
            Fe m = lo[0];
#pragma unroll
            for (int f = 1; f < D - 1; ++f) m = Fd::mul(m, lo[f]);
            mac(0, m, lo[D - 1]);
        }
        if (!SKIP1) {
            Fe m = hi[0];
#pragma unroll
            for (int f = 1; f < D - 1; ++f) m = Fd::mul(m, hi[f]);
            mac(S::of(1), m, hi[D - 1]);
        }
        if (NPTS == 3 && D == 2 && F::SLACK3P) {
            // the only extra point is t = 2 and both factors go straight into the lazy product: take the
            // unreduced 2*hi - lo + p (< 3p < 2^256) and skip two modular subtractions and two additions
            mac(S::of(2), Fd::line2_lazy(lo[0], hi[0]), Fd::line2_lazy(lo[D - 1], hi[D - 1]));
        } else {
            Fe cur[D], dl[D];
#pragma unroll
            for (int f = 0; f < D; ++f) {
                dl[f] = Fd::sub(hi[f], lo[f]);
                cur[f] = hi[f];
            }
#pragma unroll
            for (int t = 2; t < NPTS; ++t) {
#pragma unroll
                for (int f = 0; f < D; ++f) cur[f] = Fd::add(cur[f], dl[f]);
                Fe m = cur[0];
#pragma unroll
                for (int f = 1; f < D - 1; ++f) m = Fd::mul(m, cur[f]);
                mac(S::of(t), m, cur[D - 1]);
            }
        }
    }
    // The same product taken one factor at a time (SKIP1 rounds): m[k] carries the partial product at point
    // k of {0, 2, 3, ..} over the factors seen so far, so only S::N values stay live between two folds instead of
    // every folded factor (what keeps the >= 3-factor kernels inside 128 registers).
    __device__ __forceinline__ void factor(int f, const Fe& lo, const Fe& hi, Fe* m) {
        Fe cur = lo;
        const Fe d = Fd::sub(hi, lo);
#pragma unroll
        for (int k = 0; k < S::N; ++k) {
            if (k == 1) cur = Fd::add(hi, d);                 // t = 2 (t = 1 is skipped)
            else if (k > 1) cur = Fd::add(cur, d);
            if (f == 0) m[k] = cur;
            else if (f < D - 1) m[k] = Fd::mul(m[k], cur);
            else mac(k, m[k], cur);
        }
    }
    __device__ __forceinline__ void finish(Fe* out) {
#pragma unroll
        for (int p = 0; p < S::N; ++p) out[p] = Fd::reduce_wide(SM ? wide_load(sm + p * ACC_VECS * BLOCK) : w[p]);
    }
};
// D == 1 (plain sumcheck, sum_check_protocol.rs:168-175): sums of table values, modular adds.
template <class F, int NPTS, bool SKIP1, bool SM>
struct RoundAcc<F, 1, NPTS, SKIP1, SM> {
    typedef Field<F> Fd;
    typedef Slots<NPTS, SKIP1> S;
    Fe v[S::N];
    __device__ __forceinline__ void init(uint4* = nullptr) {
#pragma unroll
        for (int p = 0; p < S::N; ++p) v[p] = Fd::zero();
    }
    __device__ __forceinline__ void add_product(const Fe* lo, const Fe* hi) {
        v[0] = Fd::add(v[0], lo[0]);
        if (!SKIP1) v[S::of(1)] = Fd::add(v[S::of(1)], hi[0]);
        Fe cur = hi[0];
        Fe dl = Fd::sub(hi[0], lo[0]);
#pragma unroll
        for (int t = 2; t < NPTS; ++t) {
            cur = Fd::add(cur, dl);
            v[S::of(t)] = Fd::add(v[S::of(t)], cur);
        }
    }
    __device__ __forceinline__ void finish(Fe* out) {
#pragma unroll
        for (int p = 0; p < S::N; ++p) out[p] = v[p];
    }
};
// X*Y + Z at t = 0, (1), 2: products lazily, the Z column with modular adds.
template <class F, bool SKIP1, bool SM = false>
struct XyzAcc {
    typedef Field<F> Fd;
    typedef Slots<3, SKIP1> S;
    Wide w[SM ? 1 : S::N];
    Fe z[S::N];
    uint4* sm;
    __device__ __forceinline__ void init(uint4* slots = nullptr) {
        sm = slots;
#pragma unroll
        for (int p = 0; p < S::N; ++p) {
            if (SM) wide_store(sm + p * ACC_VECS * BLOCK, Fd::wide_zero());
            else w[p] = Fd::wide_zero();
            z[p] = Fd::zero();
        }
    }
    __device__ __forceinline__ void mac(int p, const Fe& x, const Fe& y) {
        if (SM) mac_wide_sm<F>(sm + p * ACC_VECS * BLOCK, x, y);
        else Fd::mac_wide(w[p], x, y);
    }
    __device__ __forceinline__ void add(const Fe* lo, const Fe* hi) {
        mac(0, lo[0], lo[1]);
        z[0] = Fd::add(z[0], lo[2]);
        if (!SKIP1) {
            mac(S::of(1), hi[0], hi[1]);
            z[S::of(1)] = Fd::add(z[S::of(1)], hi[2]);
        }
        Fe x2, y2;
        if (F::SLACK3P) {
            x2 = Fd::line2_lazy(lo[0], hi[0]);
            y2 = Fd::line2_lazy(lo[1], hi[1]);
        } else {
            x2 = Fd::sub(Fd::dbl(hi[0]), lo[0]);
            y2 = Fd::sub(Fd::dbl(hi[1]), lo[1]);
        }
        Fe z2 = Fd::sub(Fd::dbl(hi[2]), lo[2]);
        mac(S::of(2), x2, y2);
        z[S::of(2)] = Fd::add(z[S::of(2)], z2);
    }
    __device__ __forceinline__ void finish(Fe* out) {
#pragma unroll
        for (int p = 0; p < S::N; ++p) out[p] = Fd::add(Fd::reduce_wide(SM ? wide_load(sm + p * ACC_VECS * BLOCK) : w[p]), z[p]);
    }
};

// ------------------------------------------------------- staged table reads
// The streaming kernels read each thread's entries through a private slot of
// shared memory filled by cp.async (LDGSTS): the loads of the NEXT table /
// iteration are in flight while the current one is multiplied, without
// holding 64 registers of not-yet-arrived data per thread.  A slot is only ever
// touched by its own thread, so there is no CTA-wide synchronisation.
// Layout: stage[buf][vec][thread] of uint4 -> consecutive threads hit
// consecutive 16-byte words (bank-conflict free), global reads stay coalesced.
__device__ __forceinline__ void cp_async16(uint4* smem, const uint4* gmem) {
    const unsigned int sa = static_cast<unsigned int>(__cvta_generic_to_shared(smem));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ Fe fe_from_smem(const uint4* slot_lo, const uint4* slot_hi) {
    const uint4 a = *slot_lo, b = *slot_hi;
    Fe r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
// ---- TMA 1-D bulk copies (cp.async.bulk -> UBLKCP in SASS) completing on an mbarrier.  One lane issues a copy for
// the whole warp: no thread holds an address or a register for data in flight, which is what lets the >= 3-factor
// round kernel (at the 128-register limit) prefetch at all.
__device__ __forceinline__ unsigned int smem_u32(const void* p) { return static_cast<unsigned int>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_inval(uint64_t* bar) {
    asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned int bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned int parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "ZKB_MBAR_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra ZKB_MBAR_WAIT_%=;\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem, const void* gmem, unsigned int bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem)), "l"(gmem),
                 "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
constexpr int EVAL_BUFS = 4;   // k_sc_eval: 4 buffers x 4 vectors (lo, hi of one table)
constexpr int FOLD_BUFS = 2;   // k_sc_fold_eval: 2 buffers x 8 vectors (the quad of one table)
constexpr int STAGE_BYTES = 2 * 8 * BLOCK * 16;  // 64 KiB per CTA of staging for both kernels
constexpr int FOLD_SMEM_BYTES = STAGE_BYTES + (MAXPTS - 1) * ACC_VECS * BLOCK * 16;  // + parked accumulators

// K7: evaluations of the first round (no challenge to bind yet): all NPTS points.  Device function shared by
// the stand-alone kernel and by the persistent kernel (which can start with this pass).
template <class F, int KIND, int D, int NPTS>
__device__ __forceinline__ void eval_pass(const TabRef* __restrict__ in, int n_products, uint64_t n, uint4* stage, Fe* out) {
    const uint64_t half = n >> 1;
    const uint64_t step = (uint64_t)gridDim.x * BLOCK;
    const uint64_t j0 = (uint64_t)blockIdx.x * BLOCK + threadIdx.x;
    const int T = KIND == KIND_XYZ ? 3 : n_products * D;
    uint4* my = stage + threadIdx.x;
    // flattened prefetch sequence q = iteration * T + table
    uint64_t pj = j0;
    int pt = 0, pbuf = 0;
    auto issue = [&]() {
        if (pj < half) {
            const TabRef& t = in[pt];
            uint4* dst = my + (size_t)pbuf * 4 * BLOCK;
            cp_async16(dst, t.base + pj);
            cp_async16(dst + BLOCK, t.base + t.stride + pj);
            cp_async16(dst + 2 * BLOCK, t.base + pj + half);
            cp_async16(dst + 3 * BLOCK, t.base + t.stride + pj + half);
        }
        cp_async_commit();
        pbuf = (pbuf + 1) & (EVAL_BUFS - 1);
        if (++pt == T) {
            pt = 0;
            pj += step;
        }
    };
#pragma unroll
    for (int k = 0; k < EVAL_BUFS; ++k) issue();
    int cbuf = 0;
    auto take = [&](Fe& lo, Fe& hi) {
        cp_async_wait<EVAL_BUFS - 1>();
        const uint4* src = my + (size_t)cbuf * 4 * BLOCK;
        lo = fe_from_smem(src, src + BLOCK);
        hi = fe_from_smem(src + 2 * BLOCK, src + 3 * BLOCK);
        cbuf = (cbuf + 1) & (EVAL_BUFS - 1);
        issue();
    };
    if (KIND == KIND_XYZ) {
        XyzAcc<F, false> acc;
        acc.init();
        for (uint64_t j = j0; j < half; j += step) {
            Fe lo[3], hi[3];
#pragma unroll
            for (int f = 0; f < 3; ++f) take(lo[f], hi[f]);
            acc.add(lo, hi);
        }
        acc.finish(out);
    } else {
        RoundAcc<F, D, NPTS, false> acc;
        acc.init();
        for (uint64_t j = j0; j < half; j += step) {
            for (int p = 0; p < n_products; ++p) {
                Fe lo[D], hi[D];
#pragma unroll
                for (int f = 0; f < D; ++f) take(lo[f], hi[f]);
                acc.add_product(lo, hi);
            }
        }
        acc.finish(out);
    }
    cp_async_wait<0>();
}
template <class F, int KIND, int D, int NPTS>
__global__ void __launch_bounds__(BLOCK, ZKB_MINB) k_sc_eval(const __grid_constant__ ScArgs a) {
    extern __shared__ uint4 stage[];
    Fe out[NPTS];
    eval_pass<F, KIND, D, NPTS>(a.in, a.n_products, a.n_out, stage, out);
    finish_round<F, NPTS>(out, a.fin);
}

// K8: one HBM pass per round.  Thread j owns the quad
//   (j, j + n_out/2, j + n_out, j + n_out + n_out/2) of every input table
// (input size 2*n_out): it folds (j, j+n_out) -> new[j] and
// (j+n_out/2, j+n_out+n_out/2) -> new[j+n_out/2], stores both, and the pair
// (new[j], new[j+n_out/2]) is exactly the next round's (lo, hi).
// In-place operation (out == in) is safe: a thread only overwrites entries
// that no other thread reads.  Produces NPTS-1 per-thread sums: s(0), s(2), .. (SKIP1).
// Shared by the one-launch-per-round kernel and the persistent kernel; `rt` may live in
// the kernel parameters (constant bank) or in shared memory.
// STAGED: table reads go through the cp.async slots (memory-latency-bound shapes, D <= 2).  Products of
// three or more factors are multiplier-bound (5x the work per byte), so they read straight into
// registers and spend the shared memory on the parked accumulators only (2 CTAs/SM either way).
template <int KIND, int D>
struct Staged {
    static constexpr bool value = KIND == KIND_XYZ || D <= 2;
};
// UNITS: products of >= 3 factors take their tables one fold at a time (round_pass_async3): a unit is one fold of one table,
// 4 vectors per thread (2 elements x 2 limb planes); UnitBufs units are in flight per thread.
template <int KIND, int D>
struct Bulk {
    static constexpr bool value = KIND == KIND_PROD && D >= 3;
};
template <int NPTS>
struct BulkBufs {
    static constexpr int value = NPTS <= 4 ? 3 : 2;  // bounded by 2 CTAs/SM next to the parked accumulators
};
template <int NPTS>
struct BulkWarpBytes {  // per warp: 32 threads x 4 vectors x 16 bytes per unit buffer
    static constexpr int value = BulkBufs<NPTS>::value * 4 * 32 * 16;
};
template <int KIND, int D, int NPTS>
struct FoldSmem {
    static constexpr int stage_bytes =
        Staged<KIND, D>::value ? STAGE_BYTES : (Bulk<KIND, D>::value ? (BLOCK / 32) * BulkWarpBytes<NPTS>::value : 0);
    static constexpr int bytes = stage_bytes + (NPTS - 1) * ACC_VECS * BLOCK * 16;
};
template <int KIND, int D, int NPTS>
struct TailSmem {  // the persistent kernel may start with the evaluation pass, which stages through 64 KiB
    static constexpr int bytes = FoldSmem<KIND, D, NPTS>::bytes > STAGE_BYTES ? FoldSmem<KIND, D, NPTS>::bytes : STAGE_BYTES;
};
__device__ __forceinline__ Fe lds_fe(unsigned int a_lo, unsigned int a_hi) {
    Fe r;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.l[0]), "=r"(r.l[1]), "=r"(r.l[2]), "=r"(r.l[3]) : "r"(a_lo));
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.l[4]), "=r"(r.l[5]), "=r"(r.l[6]), "=r"(r.l[7]) : "r"(a_hi));
    return r;
}

// The >= 3-factor round pass for the sizes the tensor-core path does not take (below 2^16 entries): a unit is one fold of one
// table; NB units are in flight through per-THREAD cp.async (LDGSTS) slots, stage[buf][vec][thread] -- a slot is private to its
// thread, so cp.async.wait_group is the only synchronisation -- and the product is taken one factor at a time (RoundAcc::factor).
// (A per-warp TMA bulk-copy version of the same pipeline measured 5 % slower and was dropped.)
template <class F, int D, int NPTS>
__device__ __forceinline__ void round_pass_async3(const TabRef* __restrict__ in, const TabRef* __restrict__ outp, int n_products,
                                                  uint64_t n_out, const FixedMul& rt, uint4* stage, uint4* accs, Fe* out) {
    typedef Field<F> Fd;
    constexpr int NB = BulkBufs<NPTS>::value;
    constexpr unsigned int UNIT = 4 * BLOCK * 16;  // bytes per unit (CTA-wide): 4 vectors per thread
    const uint64_t half = n_out >> 1;
    const uint64_t step = (uint64_t)gridDim.x * BLOCK;
    const uint64_t j0 = (uint64_t)blockIdx.x * BLOCK + threadIdx.x;
    const int units = 2 * n_products * D;
    const unsigned int sb = smem_u32(stage) + threadIdx.x * 16u;
    auto issue = [&](uint64_t pj, int pu, unsigned int buf) {
        if (pj < half) {
            const TabRef& t = in[pu >> 1];
            const uint4* g0 = t.base + pj + ((pu & 1) ? half : 0);
            const uint4* g1 = g0 + t.stride;
            const unsigned int dst = sb + buf * UNIT;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(g0) : "memory");
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + BLOCK * 16u), "l"(g1) : "memory");
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 2u * BLOCK * 16u), "l"(g0 + n_out) : "memory");
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 3u * BLOCK * 16u), "l"(g1 + n_out) : "memory");
        }
        cp_async_commit();
    };
#pragma unroll
    for (int b = 0; b < NB; ++b) issue(j0, b, b);
    unsigned int cb = 0;
    RoundAcc<F, D, NPTS, true, true> acc;
    acc.init(accs);
    for (uint64_t j = j0; j < half; j += step) {
        for (int p = 0; p < n_products; ++p) {
            Fe m[NPTS - 1];
            // NOT unrolled over the factors: the body (two folds + the three forms of factor()) is 1 550 instructions;
            // unrolled it is 2 700 (43 KB), misses the instruction cache (83 % hit rate, 12 % of the stall samples
            // "no instruction") and spills 270 bytes.  Measured 3.53 -> 3.07 ms on the 2^26 launch.
#pragma unroll 1
            for (int f = 0; f < D; ++f) {
                Fe lo, hi;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    cp_async_wait<NB - 1>();
                    const unsigned int src = sb + cb * UNIT;
                    const Fe x0 = lds_fe(src, src + BLOCK * 16u), x1 = lds_fe(src + 2u * BLOCK * 16u, src + 3u * BLOCK * 16u);
                    {
                        int pu = 2 * (p * D + f) + h + NB;
                        uint64_t pj = j;
                        if (pu >= units) {
                            pu -= units;
                            pj += step;
                        }
                        issue(pj, pu, cb);
                    }
                    cb = cb + 1 == NB ? 0u : cb + 1;
                    Fe& dst = h ? hi : lo;
                    dst = Fd::fold_fixed(x0, x1, rt);
                    st_fe(outp[p * D + f], h ? j + half : j, dst);
                }
                acc.factor(f, lo, hi, m);
            }
        }
    }
    acc.finish(out);
    cp_async_wait<0>();
}

template <class F, int KIND, int D, int NPTS>
__device__ __forceinline__ void round_pass(const TabRef* __restrict__ in, const TabRef* __restrict__ outp, int n_products,
                                           uint64_t n_out, const FixedMul& rt, uint4* stage, Fe* out) {
    typedef Field<F> Fd;
    constexpr bool STAGED = Staged<KIND, D>::value;
    const uint64_t half = n_out >> 1;
    const uint64_t step = (uint64_t)gridDim.x * BLOCK;
    const uint64_t j0 = (uint64_t)blockIdx.x * BLOCK + threadIdx.x;
    const int T = KIND == KIND_XYZ ? 3 : n_products * D;
    uint4* my = stage + threadIdx.x;
    uint4* accs = stage + FoldSmem<KIND, D, NPTS>::stage_bytes / 16 + threadIdx.x;  // accumulators after the staging buffers
    if constexpr (Bulk<KIND, D>::value) {
        round_pass_async3<F, D, NPTS>(in, outp, n_products, n_out, rt, stage, accs, out);
        return;
    }
    uint64_t pj = j0;
    int pt = 0, pbuf = 0;
    auto issue = [&]() {
        if (!STAGED) return;
        if (pj < half) {
            const TabRef& t = in[pt];
            uint4* dst = my + (size_t)pbuf * 8 * BLOCK;
            const uint4* g0 = t.base + pj;
            const uint4* g1 = t.base + t.stride + pj;
            cp_async16(dst, g0);
            cp_async16(dst + BLOCK, g1);
            cp_async16(dst + 2 * BLOCK, g0 + n_out);
            cp_async16(dst + 3 * BLOCK, g1 + n_out);
            cp_async16(dst + 4 * BLOCK, g0 + half);
            cp_async16(dst + 5 * BLOCK, g1 + half);
            cp_async16(dst + 6 * BLOCK, g0 + half + n_out);
            cp_async16(dst + 7 * BLOCK, g1 + half + n_out);
        }
        cp_async_commit();
        pbuf ^= 1;
        if (++pt == T) {
            pt = 0;
            pj += step;
        }
    };
    issue();
    issue();
    int cbuf = 0;
    // fold the quad of table `t` at position j: lo = new[j], hi = new[j + half]
    auto fold_table = [&](int t, uint64_t j, Fe& lo, Fe& hi) {
        if (!STAGED) {
            const TabRef& ti = in[t];
            {
                const Fe x0 = ld_fe(ti, j), x1 = ld_fe(ti, j + n_out);
                lo = Fd::fold_fixed(x0, x1, rt);
                st_fe(outp[t], j, lo);
            }
            {
                const Fe y0 = ld_fe(ti, j + half), y1 = ld_fe(ti, j + half + n_out);
                hi = Fd::fold_fixed(y0, y1, rt);
                st_fe(outp[t], j + half, hi);
            }
            return;
        }
        cp_async_wait<FOLD_BUFS - 1>();
        const uint4* src = my + (size_t)cbuf * 8 * BLOCK;
        {
            const Fe x0 = fe_from_smem(src, src + BLOCK), x1 = fe_from_smem(src + 2 * BLOCK, src + 3 * BLOCK);
            lo = Fd::fold_fixed(x0, x1, rt);
            st_fe(outp[t], j, lo);
        }
        {
            const Fe y0 = fe_from_smem(src + 4 * BLOCK, src + 5 * BLOCK), y1 = fe_from_smem(src + 6 * BLOCK, src + 7 * BLOCK);
            cbuf ^= 1;
            issue();
            hi = Fd::fold_fixed(y0, y1, rt);
            st_fe(outp[t], j + half, hi);
        }
    };
    if (KIND == KIND_XYZ) {
        XyzAcc<F, true, true> acc;
        acc.init(accs);
        for (uint64_t j = j0; j < half; j += step) {
            Fe lo[3], hi[3];
#pragma unroll
            for (int f = 0; f < 3; ++f) fold_table(f, j, lo[f], hi[f]);
            acc.add(lo, hi);
        }
        acc.finish(out);
    } else {
        RoundAcc<F, D, NPTS, true, true> acc;
        acc.init(accs);
        for (uint64_t j = j0; j < half; j += step) {
            for (int p = 0; p < n_products; ++p) {
                Fe lo[D], hi[D];
#pragma unroll
                for (int f = 0; f < D; ++f) fold_table(p * D + f, j, lo[f], hi[f]);
                acc.add_product(lo, hi);
            }
        }
        acc.finish(out);
    }
    if (STAGED) cp_async_wait<0>();
}

template <class F, int KIND, int D, int NPTS>
__global__ void __launch_bounds__(BLOCK, ZKB_MINB) k_sc_fold_eval(const __grid_constant__ ScArgs a) {
    extern __shared__ uint4 stage[];
    Fe out[NPTS - 1];
    round_pass<F, KIND, D, NPTS>(a.in, a.out, a.n_products, a.n_out, a.rt, stage, out);
    finish_round<F, NPTS - 1>(out, a.fin);
}

}  // namespace zkb
#include "tcfold.cuh"
namespace zkb {

// ------------------------------------------------- persistent round kernel
// All remaining rounds of a sumcheck in ONE cooperative launch.  The transcript
// stays on the host: per round the kernel publishes its NPTS-1 sums to a mailbox
// in mapped host memory, the host answers with the next challenge in the same
// mailbox, and the exchange doubles as the grid-wide barrier between rounds
// (a round's sums are complete only when every CTA has stored its folded
// entries).  Per round this costs two PCIe one-way trips instead of a kernel
// launch, a stream wait and a second-stage reduction launch gap.
struct alignas(64) TailMailbox {
    // host -> device: ONE 64-byte line, polled with one coalesced read.  The line may arrive as two 32-byte
    // sectors read at different instants, so the second sector carries a checksum of the first
    // (xor of the 8 words of r, mixed with seq): a torn read (new seq, stale r) fails it and is retried.
    uint32_t r[8];
    volatile unsigned int host_seq;
    volatile unsigned int abort;
    volatile unsigned int chk;
    unsigned int pad0[5];
    // device -> host
    Fe evals[MAXPTS];
    Fe finals[MAXT];
    volatile unsigned int dev_seq;
    volatile unsigned int dev_error;  // 1 = timed out waiting for the host
    volatile unsigned int dev_chk[2]; // checksum of evals + dev_seq (messages published without a system fence)
    unsigned int pad1[12];
    // diagnostics (ZKB200_TRACE=1): device %globaltimer stamps of the last round, see RoundDriver::trace
    unsigned long long ts[8];
};
struct TailRelay {  // device memory: CTA 0 re-publishes the host's message for the other CTAs
    FixedMul rt;
    unsigned int seq;
    unsigned int abort;
    unsigned int pad[2];
    __align__(64) uint32_t line[16];  // the host's mailbox line as CTA 0 read it: challenge [0..8), number [8], checksum [10]
};
struct TailArgs {
    TabRef in[MAXT];
    TabRef out[MAXT];
    int n_tables;
    int n_products;
    uint64_t n_in;             // entries per table at entry
    FixedMul rt0;              // table of the first challenge to bind
    Fe cpow[8];                // 2^(32 i + 64) mod p: FixedMul rows are mul(r, cpow[i])
    const Fe* cpow8;           // tensor-core variant (device memory): 2^(8 i + 32) mod p, i < 32, and ONE (Montgomery) at [32]
    TcFoldMats mats0;          // tensor-core variant: the matrices of the first challenge to bind
    TailMailbox* mb;
    TailRelay* relay;
    Fe* partials;
    unsigned int* ticket;
    unsigned int base_seq;
    long long timeout_clocks;
    uint64_t stop_n;           // leave after publishing the round whose tables have <= stop_n entries (0: run to the end)
    int first_eval;            // 1: start with round 0 (all NPTS sums of the unbound tables) and take rt0 from the mailbox
    unsigned long long* dbg;   // optional [2 * gridDim]: per-CTA start/end %globaltimer of the pass of round `it == 1`
};

// TC: the folds of every round pass run on the tensor cores (round_pass_tc); the host only launches this variant when
// every round of the launch has at least 256 output entries per table (n_out / 2 a multiple of 128).
template <class F, int KIND, int D, int NPTS, bool TC>
__device__ __forceinline__ void tail_body(const TailArgs& a, uint4* stage, FixedMul& s_rt, unsigned int& s_abort, uint32_t tmem);

template <class F, int KIND, int D, int NPTS, bool TC = false>
__global__ void __launch_bounds__(BLOCK, 2) k_sc_tail(const __grid_constant__ TailArgs a) {
    extern __shared__ __align__(128) uint4 stage[];
    __shared__ FixedMul s_rt;
    __shared__ unsigned int s_abort;
    __shared__ uint32_t s_tmem;
    uint32_t tmem = 0;
    if (TC) {
        if (threadIdx.x < 32) tmem_alloc(&s_tmem, TcCfg<NPTS>::tmem_cols);
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        tmem = s_tmem;
    }
    tail_body<F, KIND, D, NPTS, TC>(a, stage, s_rt, s_abort, tmem);
    if (TC) {  // every exit of the body is CTA-uniform
        tc_fence_before();
        __syncthreads();
        if (threadIdx.x < 32) tmem_dealloc(tmem, TcCfg<NPTS>::tmem_cols);
    }
}
template <class F, int KIND, int D, int NPTS, bool TC>
__device__ __forceinline__ void tail_body(const TailArgs& a, uint4* stage, FixedMul& s_rt, unsigned int& s_abort, uint32_t tmem) {
    typedef Field<F> Fd;
    uint64_t n_in = a.n_in;
    if (threadIdx.x == 0) s_abort = 0;
    if (a.first_eval) {  // round 0 inside the launch: message 1 = s(0..d) of the unbound tables
        const uint64_t ctas = ((n_in >> 1) + BLOCK - 1) / BLOCK;
        const unsigned int n_active = ctas < 1 ? 1u : (ctas < gridDim.x ? (unsigned int)ctas : gridDim.x);
        if (blockIdx.x < n_active) {
            Fe out[NPTS];
            eval_pass<F, KIND, D, NPTS>(a.in, a.n_products, n_in, stage, out);
            FinishArgs fin;
            fin.partials = a.partials;
            fin.ticket = a.ticket;
            fin.result = a.mb->evals;
            fin.result_wide = nullptr;
            fin.flag = &a.mb->dev_seq;
            fin.seq = a.base_seq + 1;
            fin.stamp = nullptr;
            fin.chk = a.mb->dev_chk;
            finish_round<F, NPTS>(out, fin, n_active);
        }
    }
    const unsigned int it0 = a.first_eval ? 1u : 0u;
    for (unsigned int it = it0;; ++it) {
        // ---- the challenge table of this round -> shared memory
        if (it == 0) {
            for (int w = threadIdx.x; w < 64; w += BLOCK) (&s_rt.t[0][0])[w] = (&a.rt0.t[0][0])[w];
            if (TC) {
                for (int w = threadIdx.x; w < 128; w += BLOCK)
                    reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(stage) + TcRoundSmem<NPTS>::mats_off)[w] = reinterpret_cast<const uint4*>(&a.mats0)[w];
            }
            if (threadIdx.x == 0) s_abort = 0;
        } else {
            const unsigned int want = a.base_seq + it;
            if (blockIdx.x == 0 && threadIdx.x < 32) {  // relay warp: host mailbox -> device memory
                const int lane = threadIdx.x;
                const volatile uint32_t* line = reinterpret_cast<const volatile uint32_t*>(a.mb);
                const long long t0 = clock64();
                uint32_t word = 0;
                unsigned int status = 0;  // 1 = got it, 2 = abort, 3 = timeout
                while (status == 0) {
                    word = lane < 16 ? line[lane] : 0u;  // one coalesced 64-byte read over PCIe
                    const uint32_t seq = __shfl_sync(0xffffffffu, word, 8);
                    const uint32_t ab = __shfl_sync(0xffffffffu, word, 9);
                    const uint32_t chk = __shfl_sync(0xffffffffu, word, 10);
                    const uint32_t x = __reduce_xor_sync(0xffffffffu, lane < 8 ? word : 0u);
                    if (ab) status = 2;
                    else if (seq == want && (x ^ (want * 0x9E3779B9u)) == chk) status = 1;
                    else if (clock64() - t0 > a.timeout_clocks) status = 3;
                }
                if (status == 1) {
                    // The line goes on as it was read -- challenge, number and checksum: every CTA validates it the same way
                    // (no fence, no second flag) and builds what it needs from the challenge itself: 64 (tensor cores) or 8
                    // Montgomery products per CTA are cheaper than one warp building 2 KiB and 296 CTAs re-reading them
                    // (relay 2.4 -> 0.4 us per round, ZKB200_TRACE=2).
                    if (lane == 0) a.mb->ts[0] = gtime();
                    if (lane < 16) __stcg(&a.relay->line[lane], word);
                    if (lane == 0) a.mb->ts[1] = gtime();
                } else if (lane == 0) {
                    if (status == 3) {
                        a.mb->dev_error = 1;
                        __threadfence_system();
                    }
                    *reinterpret_cast<volatile unsigned int*>(&a.relay->abort) = status;
                }
            }
            if (threadIdx.x < 32) {
                // all lanes of warp 0 poll together (one 64-byte read): a spin by a single lane leaves the warp split, and
                // its next shuffles were measured to take ~10 us (k_sc_small's cluster messages)
                const int lane = threadIdx.x;
                const volatile uint32_t* rl = reinterpret_cast<const volatile uint32_t*>(a.relay->line);
                const volatile unsigned int* ab = reinterpret_cast<const volatile unsigned int*>(&a.relay->abort);
                unsigned int aborted = 0;
                uint32_t word;
                for (;;) {
                    word = lane < 16 ? rl[lane] : 0u;
                    const uint32_t seq = __shfl_sync(0xffffffffu, word, 8);
                    const uint32_t chk = __shfl_sync(0xffffffffu, word, 10);
                    const uint32_t x = __reduce_xor_sync(0xffffffffu, lane < 8 ? word : 0u);
                    if (seq == want && (x ^ (want * 0x9E3779B9u)) == chk) break;
                    if ((aborted = *ab)) break;
                    __nanosleep(32);
                }
                if (lane == 0) s_abort = aborted;
                __threadfence();  // (tables written by other CTAs in the previous round are read after this)
                if (!aborted) {
                    Fe r;
#pragma unroll
                    for (int k = 0; k < 8; ++k) r.l[k] = __shfl_sync(0xffffffffu, word, k);
                    if (TC) {  // lane i: column i of both byte matrices, T1_i = (1 - r) 2^(8 i + 32), T2_i = r 2^(8 i + 32) mod p
                        const Fe cp = a.cpow8[lane];
                        const Fe t1 = Fd::mul(Fd::sub(a.cpow8[32], r), cp), t2 = Fd::mul(r, cp);
                        uint8_t* m0 = reinterpret_cast<uint8_t*>(stage) + TcRoundSmem<NPTS>::mats_off + (lane / 16) * 512 + lane % 16;
#pragma unroll
                        for (int n = 0; n < 32; ++n) {
                            m0[n * 16] = (uint8_t)(t1.l[n / 4] >> (8 * (n % 4)));
                            m0[1024 + n * 16] = (uint8_t)(t2.l[n / 4] >> (8 * (n % 4)));
                        }
                    } else if (lane < 8) {  // row i of the fixed-multiplicand table: r 2^(32 i + 64)
                        const Fe t = Fd::mul(r, a.cpow[lane]);
#pragma unroll
                        for (int k = 0; k < 8; ++k) s_rt.t[lane][k] = t.l[k];
                    }
                }
            }
            __syncthreads();
            if (s_abort) return;
        }
        if (TC) fence_proxy_async();  // the matrices: generic-proxy writes, read by the tensor core through the async proxy
        __syncthreads();
        const uint64_t n_out = n_in >> 1;
        const TabRef* src = it == it0 ? a.in : a.out;
        if (n_out == 1) {  // last bind: publish the bound values and leave
            if (blockIdx.x == 0) {
                const int t = threadIdx.x;
                if (t < a.n_tables) {
                    Fe v = Fd::fold_fixed(ld_fe(src[t], 0), ld_fe(src[t], 1), s_rt);
                    st_fe(a.out[t], 0, v);
                    a.mb->finals[t] = v;
                    __threadfence_system();
                }
                __syncthreads();
                if (t == 0) a.mb->dev_seq = a.base_seq + it + 1;
            }
            return;
        }
        // only the CTAs that own quads this round compute and take a ticket
        const uint64_t ctas = ((n_out >> 1) + BLOCK - 1) / BLOCK;
        const unsigned int n_active = ctas < 1 ? 1u : (ctas < gridDim.x ? (unsigned int)ctas : gridDim.x);
        if (blockIdx.x < n_active) {
            Fe out[NPTS - 1];
            if (blockIdx.x == 0 && threadIdx.x == 0) a.mb->ts[2] = gtime();
            if (a.dbg && it == 1 && threadIdx.x == 0) a.dbg[2 * blockIdx.x] = gtime();
            if constexpr (TC) round_pass_tc<F, D, NPTS>(src, a.out, a.n_products, n_out, reinterpret_cast<uint8_t*>(stage), tmem, out);
            else round_pass<F, KIND, D, NPTS>(src, a.out, a.n_products, n_out, s_rt, stage, out);
            if (blockIdx.x == 0 && threadIdx.x == 0) a.mb->ts[3] = gtime();
            if (a.dbg && it == 1) {
                __syncthreads();
                if (threadIdx.x == 0) a.dbg[2 * blockIdx.x + 1] = gtime();
            }
            FinishArgs fin;
            fin.partials = a.partials;
            fin.ticket = a.ticket;
            fin.result = a.mb->evals;
            fin.result_wide = nullptr;
            fin.flag = &a.mb->dev_seq;
            fin.seq = a.base_seq + it + 1;
            fin.stamp = &a.mb->ts[4];
            fin.chk = a.mb->dev_chk;
            finish_round<F, NPTS - 1>(out, fin, n_active);
        }
        n_in = n_out;
        if (n_in <= a.stop_n) return;  // the shared-memory kernel k_sc_small takes over
    }
}

// ------------------------------------------------ small tables: k_sc_small
// Every remaining round of a sumcheck whose tables fit in shared memory, in ONE single-CTA launch.
// The tables are read from HBM once; after that a round is: fold in shared memory (one output entry
// per thread), evaluate (one pair per thread, plain Montgomery products and modular adds -- latency,
// not throughput, matters here), CTA reduction, mailbox exchange with the host transcript.
// With first_eval the kernel also produces round 0 (all NPTS points), so a whole GKR phase on a
// small layer is one launch.
// Tables of up to 16 x that size run the same way in ONE THREAD-BLOCK CLUSTER of nc = 2..16 CTAs: CTA c keeps the
// entries j = c (mod nc) of every table in its shared memory (the multi-GPU layout of zkb200.cu: the two entries of a
// pair differ in the top index bit, so they live in the same CTA), every CTA polls the mailbox for the challenge, and
// the CTAs' partial sums meet in CTA 0's shared memory (distributed shared memory stores + one arrival counter, no
// cluster barrier per round).  Once a CTA holds a single entry per table, CTA 0 collects them and finishes alone.
constexpr int SMALL_BLOCK = 512;
constexpr int SMALL_MAX_CLUSTER = 16;
constexpr int SMALL_SMEM_MAX = 200 * 1024;

// ---- device-side transcript.  With `enabled`, k_sc_small derives every challenge itself: warp 0
// interpolates the round polynomial (evaluations at 0, 1, 2 -> trimmed ascending coefficients,
// univariate_polynomial_dense.rs:14-18,48-74), serialises it as canonical little-endian bytes
// (fiat_shamir_transcript.rs:32-37), absorbs it into the running Keccak-256 sponge handed over by the host,
// squeezes the digest, re-seeds the sponge with it and reduces it mod p (fiat_shamir_transcript.rs:23-29).
// No round waits for PCIe.  Every round's sums AND the challenge drawn after them are written to a record in
// mapped host memory; the host replays its own transcript on those sums behind the kernel and refuses the
// proof if any challenge differs, so the host transcript stays the checked source of truth.
constexpr int DT_MAX_ROUNDS = 32;
struct DtRound {
    Fe evals[MAXPTS];
    Fe chal;
    long long t[6];  // clock64 stamps of the transcript step (diagnostics)
};
struct DtArgs {
    int enabled;
    uint32_t fill_words;  // 64-bit words absorbed since the last permutation (already XORed into st), < 17
    uint64_t st[25];      // Keccak-f[1600] state
    Fe claim0;            // s_prev(r0): the running claim when the launch starts by binding r0 (first_eval == 0)
    DtRound* rounds;      // [DT_MAX_ROUNDS] mapped host memory
};
__constant__ uint64_t KECCAK_RC[24] = {
    0x0000000000000001ull, 0x0000000000008082ull, 0x800000000000808aull, 0x8000000080008000ull, 0x000000000000808bull,
    0x0000000080000001ull, 0x8000000080008081ull, 0x8000000000008009ull, 0x000000000000008aull, 0x0000000000000088ull,
    0x0000000080008009ull, 0x000000008000000aull, 0x000000008000808bull, 0x800000000000008bull, 0x8000000000008089ull,
    0x8000000000008003ull, 0x8000000000008002ull, 0x8000000000000080ull, 0x000000000000800aull, 0x800000008000000aull,
    0x8000000080008081ull, 0x8000000000008080ull, 0x0000000080000001ull, 0x8000000080008008ull};
__device__ const uint8_t KECCAK_RHO[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};
// One warp, lane x + 5y holds lane A[x][y] of the state as (lo, hi); lanes 25..31 follow along with lane-0 sources.
struct KeccakWarp {
    int c1, c2, c3, c4, dm, dp, pi, x1, x2;
    uint32_t rho;
    bool l0;
    __device__ __forceinline__ void init(int lane) {
        const int l = lane < 25 ? lane : 0, x = l % 5, y = l / 5;
        c1 = (l + 5) % 25; c2 = (l + 10) % 25; c3 = (l + 15) % 25; c4 = (l + 20) % 25;
        dm = 5 * y + (x + 4) % 5;
        dp = 5 * y + (x + 1) % 5;
        pi = (x + 3 * y) % 5 + 5 * x;  // B[x][y] = A[(x + 3y) % 5][x] rotated
        x1 = 5 * y + (x + 1) % 5;
        x2 = 5 * y + (x + 2) % 5;
        rho = KECCAK_RHO[pi];          // the rotation belongs to the SOURCE lane of the pi step
        l0 = lane == 0;
    }
    __device__ __forceinline__ void permute(uint32_t& lo, uint32_t& hi) const {
        const unsigned int FULL = 0xffffffffu;
#pragma unroll 1
        for (int rnd = 0; rnd < 24; ++rnd) {
            // theta
            uint32_t cl = lo ^ __shfl_sync(FULL, lo, c1) ^ __shfl_sync(FULL, lo, c2) ^ __shfl_sync(FULL, lo, c3) ^ __shfl_sync(FULL, lo, c4);
            uint32_t ch = hi ^ __shfl_sync(FULL, hi, c1) ^ __shfl_sync(FULL, hi, c2) ^ __shfl_sync(FULL, hi, c3) ^ __shfl_sync(FULL, hi, c4);
            const uint32_t ml = __shfl_sync(FULL, cl, dm), mh = __shfl_sync(FULL, ch, dm);
            const uint32_t pl = __shfl_sync(FULL, cl, dp), ph = __shfl_sync(FULL, ch, dp);
            lo ^= ml ^ __funnelshift_l(ph, pl, 1);  // rotl64(C[x+1], 1)
            hi ^= mh ^ __funnelshift_l(pl, ph, 1);
            // pi (gather) then rho with the source lane's offset
            uint32_t bl = __shfl_sync(FULL, lo, pi), bh = __shfl_sync(FULL, hi, pi);
            if (rho & 32) {
                const uint32_t t = bl;
                bl = bh;
                bh = t;
            }
            const uint32_t rl = __funnelshift_l(bh, bl, rho & 31), rh = __funnelshift_l(bl, bh, rho & 31);
            // chi
            const uint32_t l1 = __shfl_sync(FULL, rl, x1), h1 = __shfl_sync(FULL, rh, x1);
            const uint32_t l2 = __shfl_sync(FULL, rl, x2), h2 = __shfl_sync(FULL, rh, x2);
            lo = rl ^ (~l1 & l2);
            hi = rh ^ (~h1 & h2);
            if (l0) {  // iota
                const uint64_t rc = KECCAK_RC[rnd];
                lo ^= (uint32_t)rc;
                hi ^= (uint32_t)(rc >> 32);
            }
        }
    }
};
struct SmallArgs {
    TabRef in[MAXT];
    TabRef out[MAXT];   // receives the bound value at entry 0
    int n_tables;
    int n_products;
    uint32_t n_in;      // entries per table at entry
    int first_eval;
    Fe r0;              // first challenge to bind (ignored with first_eval)
    TailMailbox* mb;
    unsigned int base_seq;
    long long timeout_clocks;
    DtArgs dt;          // device-side Fiat-Shamir transcript (NPTS == 3 shapes)
    int nc;             // CTAs of the launch (one thread-block cluster; 1 = a single CTA): see k_sc_small
};

// ---- thread-block cluster helpers (k_sc_small with nc > 1)
__device__ __forceinline__ uint32_t cluster_cta_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// generic address of the same shared-memory object in CTA `rank` of the cluster (distributed shared memory)
template <class T>
__device__ __forceinline__ T* cluster_map(T* p, uint32_t rank) {
    uint64_t out;
    asm volatile("mapa.u64 %0, %1, %2;" : "=l"(out) : "l"((uint64_t)(uintptr_t)p), "r"(rank));
    return (T*)(uintptr_t)out;
}

// integrand at t = 0, (1), 2, .. for one pair position, accumulated with modular adds
template <class F, int KIND, int D, int NPTS, bool SKIP1>
__device__ __forceinline__ void eval_direct(const Fe* lo, const Fe* hi, int n_products_unused, Fe* acc) {
    typedef Field<F> Fd;
    typedef Slots<NPTS, SKIP1> S;
    (void)n_products_unused;
    if (KIND == KIND_XYZ) {
        acc[0] = Fd::add(acc[0], Fd::add(Fd::mul(lo[0], lo[1]), lo[2]));
        if (!SKIP1) acc[S::of(1)] = Fd::add(acc[S::of(1)], Fd::add(Fd::mul(hi[0], hi[1]), hi[2]));
        const Fe x2 = Fd::sub(Fd::dbl(hi[0]), lo[0]), y2 = Fd::sub(Fd::dbl(hi[1]), lo[1]), z2 = Fd::sub(Fd::dbl(hi[2]), lo[2]);
        acc[S::of(2)] = Fd::add(acc[S::of(2)], Fd::add(Fd::mul(x2, y2), z2));
        return;
    }
    {
        Fe m = lo[0];
#pragma unroll
        for (int f = 1; f < D; ++f) m = Fd::mul(m, lo[f]);
        acc[0] = Fd::add(acc[0], m);
    }
    if (!SKIP1) {
        Fe m = hi[0];
#pragma unroll
        for (int f = 1; f < D; ++f) m = Fd::mul(m, hi[f]);
        acc[S::of(1)] = Fd::add(acc[S::of(1)], m);
    }
    Fe cur[D], dl[D];
#pragma unroll
    for (int f = 0; f < D; ++f) {
        dl[f] = Fd::sub(hi[f], lo[f]);
        cur[f] = hi[f];
    }
#pragma unroll
    for (int t = 2; t < NPTS; ++t) {
#pragma unroll
        for (int f = 0; f < D; ++f) cur[f] = Fd::add(cur[f], dl[f]);
        Fe m = cur[0];
#pragma unroll
        for (int f = 1; f < D; ++f) m = Fd::mul(m, cur[f]);
        acc[S::of(t)] = Fd::add(acc[S::of(t)], m);
    }
}

template <class F, int KIND, int D, int NPTS>
__global__ void __launch_bounds__(SMALL_BLOCK) k_sc_small(const __grid_constant__ SmallArgs a) {
    typedef Field<F> Fd;
    extern __shared__ uint4 tab[];  // table t: plane 0 at tab + 2*t*cap, plane 1 at tab + (2*t+1)*cap
    __shared__ Fe s_r;
    __shared__ unsigned int s_status;
    __shared__ Fe s_red[SMALL_BLOCK / 32][MAXPTS];
    // Cluster messages carry a checksum over payload and message number instead of being fenced (a cluster-scope fence is
    // MEMBAR.ALL.GPU + ERRBAR + CGAERRBAR + CCTL.IVALL, ~0.3 us per hop: tools/clustertest.cu): the reader re-reads until the
    // checksum matches.
    __shared__ Fe s_part[SMALL_MAX_CLUSTER][MAXPTS];       // CTA 0: the other CTAs' partial sums of the current round
    __shared__ unsigned int s_pchk[SMALL_MAX_CLUSTER][2];  // CTA 0: their checksums (msg_checksum)
    __shared__ __align__(64) uint32_t s_cmsg[16];          // CTAs 1..: the mailbox line as CTA 0 read it (same layout as the host's)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int T = a.n_tables;
    const uint32_t NC0 = a.nc > 1 ? (uint32_t)a.nc : 1u;
    const uint32_t crank = NC0 > 1 ? cluster_cta_rank() : 0u;
    uint32_t nc = NC0;                 // CTAs still working (1 after CTA 0 collected the tables)
    const uint32_t cap = a.n_in / NC0;  // entries per table in this CTA
    const int KD = KIND == KIND_XYZ ? 3 : D;
    for (uint32_t idx = tid; idx < (uint32_t)T * cap; idx += SMALL_BLOCK) {
        const uint32_t t = idx / cap, l = idx - t * cap;
        const uint32_t j = l * NC0 + crank;
        tab[(size_t)(2 * t) * cap + l] = a.in[t].base[j];
        tab[(size_t)(2 * t + 1) * cap + l] = a.in[t].base[a.in[t].stride + j];
    }
    if (tid == 0) {
        s_r = a.r0;
        s_status = 1;
    }
    if (tid < 16) s_cmsg[tid] = tid == 8 ? a.base_seq : 0u;  // (no challenge has this number: they start at base_seq + 1)
    if (tid < 2 * SMALL_MAX_CLUSTER) s_pchk[tid >> 1][tid & 1] = 0u;
    __syncthreads();
    if (NC0 > 1) cluster_sync_all();  // every CTA of the cluster runs (and CTA 0's counter is set) before any remote access
    auto ld = [&](int t, uint32_t j) { return fe_from_smem(tab + (size_t)(2 * t) * cap + j, tab + (size_t)(2 * t + 1) * cap + j); };
    auto st = [&](int t, uint32_t j, const Fe& v) {
        tab[(size_t)(2 * t) * cap + j] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
        tab[(size_t)(2 * t + 1) * cap + j] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
    };
    unsigned int pubs = 0;  // messages published so far
    // device transcript state (warp 0 only): the sponge lives in registers, one 64-bit lane per thread
    const bool dt = NPTS == 3 && a.dt.enabled;
    __shared__ Fe s_ev[MAXPTS], s_co[3], s_claim;
    __shared__ uint64_t s_words[12];
    __shared__ long long s_t[4];
    __shared__ int s_len, s_all;
    const int nthr = dt ? SMALL_BLOCK - 32 : SMALL_BLOCK;  // device transcript: the last warp does not fold or evaluate
    KeccakWarp kw;
    uint32_t k_lo = 0, k_hi = 0;
    int k_pos = 0;
    if (dt && warp == 0) {
        kw.init(lane);
        if (lane < 25) {
            k_lo = (uint32_t)a.dt.st[lane];
            k_hi = (uint32_t)(a.dt.st[lane] >> 32);
        }
        k_pos = (int)a.dt.fill_words;
        if (lane == 0) s_claim = a.dt.claim0;
    }
    // One transcript step, critical part, by warp 0: the round's sums are in s_ev (`all` = every point was computed;
    // otherwise {s(0), s(2)} and s(1) = claim - s(0)); leaves the next challenge in s_r and the coefficients in s_co.
    auto dt_step = [&](bool all) {
        int len = 0;
        const long long t0 = clock64();
        if (lane == 0) {
            const Fe e0 = s_ev[0];
            const Fe e1 = all ? s_ev[1] : Fd::sub(s_claim, e0);
            const Fe e2 = all ? s_ev[2] : s_ev[1];
            const Fe c2 = Fd::half(Fd::add(Fd::sub(e2, Fd::dbl(e1)), e0));
            const Fe c1 = Fd::sub(Fd::sub(e1, e0), c2);
            s_co[0] = e0;
            s_co[1] = c1;
            s_co[2] = c2;
            len = !Fd::is_zero(c2) ? 3 : (!Fd::is_zero(c1) ? 2 : (!Fd::is_zero(e0) ? 1 : 0));  // trim (:14-18)
            s_len = len;
            s_all = all ? 1 : 0;
        }
        len = __shfl_sync(0xffffffffu, len, 0);
        __syncwarp();
        if (lane < 3) {  // canonical little-endian limbs of each coefficient (fq_vec_to_bytes): c * R^-1
            const Fe cv = Fd::redc256(s_co[lane]);
#pragma unroll
            for (int w = 0; w < 4; ++w) s_words[lane * 4 + w] = (uint64_t)cv.l[2 * w] | ((uint64_t)cv.l[2 * w + 1] << 32);
        }
        __syncwarp();
        const long long t1 = clock64();
        for (int i = 0; i < len * 4; ++i) {  // absorb; the rate is 17 words
            const uint64_t w = s_words[i];
            if (lane == k_pos) {
                k_lo ^= (uint32_t)w;
                k_hi ^= (uint32_t)(w >> 32);
            }
            if (++k_pos == 17) {
                kw.permute(k_lo, k_hi);
                k_pos = 0;
            }
        }
        if (lane == k_pos) k_lo ^= 0x01u;       // Keccak (not SHA-3) padding: 0x01 .. 0x80
        if (lane == 16) k_hi ^= 0x80000000u;
        kw.permute(k_lo, k_hi);
        if (lane < 4) s_words[lane] = (uint64_t)k_lo | ((uint64_t)k_hi << 32);  // the digest
        if (lane >= 4) k_lo = k_hi = 0;         // fresh sponge seeded with the digest (:24-26)
        k_pos = 4;
        __syncwarp();
        const long long t2 = clock64();
        if (lane == 0) {
            Fe d;
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                d.l[2 * w] = (uint32_t)s_words[w];
                d.l[2 * w + 1] = (uint32_t)(s_words[w] >> 32);
            }
            s_r = Fd::to_mont(Fd::mod_p(d));  // from_le_bytes_mod_order, then into Montgomery form
            s_t[0] = t0;
            s_t[1] = t1;
            s_t[2] = t2;
            s_t[3] = clock64();
        }
        __syncwarp();
    };
    // ... and the part nobody waits for, by the last warp while the others fold: the next claim s(r) (Horner) and the
    // round's record for the host (sums, challenge), then the sequence number.
    auto dt_post = [&]() {
        if (tid != SMALL_BLOCK - 32) return;
        const Fe r = s_r;
        const int len = s_len;
        Fe cl = Fd::zero();
        if (len > 0) cl = s_co[len - 1];
        for (int k = len - 2; k >= 0; --k) cl = Fd::add(Fd::mul(cl, r), s_co[k]);
        s_claim = cl;
        DtRound* rec = a.dt.rounds + pubs;
        for (int p = 0; p < (s_all ? NPTS : NPTS - 1); ++p) rec->evals[p] = s_ev[p];
        rec->chal = r;
        for (int k = 0; k < 4; ++k) rec->t[k] = s_t[k];
        const long long f0 = clock64();
        __threadfence_system();
        rec->t[4] = f0;
        rec->t[5] = clock64();
        a.mb->dev_seq = a.base_seq + pubs + 1;
    };
    // CTA-wide modular sum of `n` per-thread values and publication as message pubs+1
    auto publish = [&](Fe* acc, int n) {
        for (int p = 0; p < n; ++p) {
            Fe v = warp_sum<F>(acc[p]);
            if (lane == 0) s_red[warp][p] = v;
        }
        __syncthreads();
        if (warp == 0) {
            for (int p = 0; p < n; ++p) {
                Fe v = lane < SMALL_BLOCK / 32 ? s_red[lane][p] : Fd::zero();
                v = warp_sum<F>(v);
                if (lane == 0) s_ev[p] = v;
            }
            if (nc > 1) {  // the cluster's sums meet in CTA 0
                __syncwarp();
                const unsigned int seq = a.base_seq + pubs + 1;
                if (crank != 0) {
                    if (lane == 0) {
                        unsigned int c0, c1;
                        msg_checksum(s_ev, n, seq, &c0, &c1);
                        Fe* rp = cluster_map(&s_part[crank][0], 0);
                        for (int p = 0; p < n; ++p) rp[p] = s_ev[p];
                        volatile unsigned int* rc = cluster_map(&s_pchk[crank][0], 0);
                        rc[0] = c0;
                        rc[1] = c1;
                    }
                } else {
                    {  // lane c waits for CTA c's message; the loop is warp-uniform (a divergent spin before the
                       // shuffles of warp_sum measured 10 us: the warp stayed split)
                        const bool mine = lane >= 1 && lane < (int)NC0;
                        const volatile uint32_t* w = reinterpret_cast<const volatile uint32_t*>(&s_part[mine ? lane : 0][0]);
                        const volatile unsigned int* ck = &s_pchk[mine ? lane : 0][0];
                        const long long t0 = clock64();
                        bool ok = !mine;
                        while (!__all_sync(0xffffffffu, ok)) {
                            unsigned int x = seq * 0x9E3779B9u, y = seq ^ 0x85EBCA6Bu;
                            for (int k = 0; k < n * 8; ++k) {
                                const unsigned int wv = w[k];
                                x ^= wv;
                                y = ((y << 5) | (y >> 27)) + wv;
                            }
                            ok = ok || (x == ck[0] && y == ck[1]);
                            if (clock64() - t0 > a.timeout_clocks) {
                                if (lane == 0) {
                                    s_status = 3;
                                    a.mb->dev_error = 1;
                                    __threadfence_system();
                                }
                                break;
                            }
                        }
                    }
                    __syncwarp();
                    for (int p = 0; p < n; ++p) {
                        Fe v = lane == 0 ? s_ev[p] : (lane < (int)NC0 ? s_part[lane][p] : Fd::zero());
                        v = warp_sum<F>(v);
                        if (lane == 0) s_ev[p] = v;
                    }
                }
                __syncwarp();
            }
            if (!dt && crank == 0 && lane == 0)
                for (int p = 0; p < n; ++p) a.mb->evals[p] = s_ev[p];
            if (crank != 0) {
            } else if (dt) {
                __syncwarp();
                dt_step(n == NPTS);
            } else if (lane == 0) {  // no system fence: the host validates the checksum (FinishArgs::chk)
                unsigned int c0, c1;
                msg_checksum(s_ev, n, a.base_seq + pubs + 1, &c0, &c1);
                a.mb->dev_chk[0] = c0;
                a.mb->dev_chk[1] = c1;
                a.mb->dev_seq = a.base_seq + pubs + 1;
            }
        }
        __syncthreads();
        if (dt) dt_post();
        ++pubs;
    };
    uint32_t m = cap;  // entries per table in this CTA (all of them without a cluster)
    if (a.first_eval) {  // round 0: all NPTS points of the unbound tables
        Fe acc[NPTS];
#pragma unroll
        for (int p = 0; p < NPTS; ++p) acc[p] = Fd::zero();
        const uint32_t half = m >> 1;
        for (uint32_t j = tid; tid < nthr && j < half; j += nthr) {
            for (int p = 0; p < (KIND == KIND_XYZ ? 1 : a.n_products); ++p) {
                Fe lo[KD], hi[KD];
#pragma unroll
                for (int f = 0; f < KD; ++f) {
                    lo[f] = ld(p * KD + f, j);
                    hi[f] = ld(p * KD + f, j + half);
                }
                eval_direct<F, KIND, D, NPTS, false>(lo, hi, 0, acc);
            }
        }
        publish(acc, NPTS);
        if (s_status != 1) return;
    }
    for (unsigned int chal = a.first_eval ? 1u : 0u;; ++chal) {
        if (chal > 0 && !dt) {  // challenge number `chal` from the host mailbox
            if (warp == 0) {
                // cluster: only CTA 0 reads the host's mailbox (reads of one host line by several CTAs serialise on PCIe:
                // measured 3 us per polling CTA and round) and passes the line on to the other CTAs' shared memories
                const bool from_host = crank == 0;
                const unsigned int want = a.base_seq + chal;
                const volatile uint32_t* line = from_host ? reinterpret_cast<const volatile uint32_t*>(a.mb) : reinterpret_cast<const volatile uint32_t*>(s_cmsg);
                const long long t0 = clock64();
                uint32_t word = 0;
                unsigned int status = 0;
                while (status == 0) {
                    word = lane < 16 ? line[lane] : 0u;
                    const uint32_t seq = __shfl_sync(0xffffffffu, word, 8);
                    const uint32_t ab = __shfl_sync(0xffffffffu, word, 9);
                    const uint32_t chk = __shfl_sync(0xffffffffu, word, 10);
                    const uint32_t x = __reduce_xor_sync(0xffffffffu, lane < 8 ? word : 0u);
                    if (ab) status = 2;
                    else if (seq == want && (x ^ (want * 0x9E3779B9u)) == chk) status = 1;
                    else if (clock64() - t0 > a.timeout_clocks) status = 3;
                }
                if (lane < 8) s_r.l[lane] = word;
                if (lane == 0) {
                    s_status = status;
                    if (status == 3 && from_host) {
                        a.mb->dev_error = 1;
                        __threadfence_system();
                    }
                }
                if (nc > 1 && from_host) {  // as read (number and checksum included); anything but a challenge becomes an abort
                    const uint32_t fw = lane == 9 && status != 1 ? 1u : word;
                    for (uint32_t cc = 1; cc < NC0; ++cc)
                        if (lane < 16) *reinterpret_cast<volatile uint32_t*>(cluster_map(&s_cmsg[lane], cc)) = fw;
                }
            }
            __syncthreads();
            if (s_status != 1) return;
        }
        const Fe r = s_r;
        // fold: one output entry per thread, in place (entry j reads j and j + n_out, writes j)
        const uint32_t n_out = m >> 1;
        for (uint32_t idx = tid; tid < nthr && idx < (uint32_t)T * n_out; idx += nthr) {
            const uint32_t t = idx / n_out, j = idx - t * n_out;
            st(t, j, Fd::fold(ld(t, j), ld(t, j + n_out), r));
        }
        __syncthreads();
        m = n_out;
        if (nc > 1 && m == 1) {
            // one entry per table and CTA left: CTA 0 collects them (entry c of a table comes from CTA c; staged behind
            // the entries CTA 0 itself still reads, cap >= 2 NC0) and finishes the remaining log2(NC0) rounds alone
            const unsigned int gseq = ~(a.base_seq + pubs + 1);  // (not the number of a round message)
            if (crank != 0) {
                if (tid == 0) {
                    uint4* rt = cluster_map(tab, 0);
                    unsigned int x = gseq * 0x9E3779B9u, y = gseq ^ 0x85EBCA6Bu;
                    for (int t = 0; t < T; ++t)
                        for (int h = 0; h < 2; ++h) {
                            const uint4 v = tab[(size_t)(2 * t + h) * cap];
                            rt[(size_t)(2 * t + h) * cap + NC0 + crank] = v;
                            const unsigned int wv[4] = {v.x, v.y, v.z, v.w};
                            for (int k = 0; k < 4; ++k) {
                                x ^= wv[k];
                                y = ((y << 5) | (y >> 27)) + wv[k];
                            }
                        }
                    volatile unsigned int* rc = cluster_map(&s_pchk[crank][0], 0);
                    rc[0] = x;
                    rc[1] = y;
                }
                return;
            }
            if (warp == 0) {  // lane c waits for CTA c's entries (warp-uniform loop, see publish)
                const bool mine = lane >= 1 && lane < (int)NC0;
                const volatile unsigned int* ck = &s_pchk[mine ? lane : 0][0];
                const long long t0 = clock64();
                bool ok = !mine;
                while (!__all_sync(0xffffffffu, ok)) {
                    unsigned int x = gseq * 0x9E3779B9u, y = gseq ^ 0x85EBCA6Bu;
                    for (int t = 0; t < T; ++t)
                        for (int h = 0; h < 2; ++h) {
                            const volatile uint32_t* w = reinterpret_cast<const volatile uint32_t*>(tab + (size_t)(2 * t + h) * cap + NC0 + (mine ? lane : 0));
                            for (int k = 0; k < 4; ++k) {
                                const unsigned int wv = w[k];
                                x ^= wv;
                                y = ((y << 5) | (y >> 27)) + wv;
                            }
                        }
                    ok = ok || (x == ck[0] && y == ck[1]);
                    if (clock64() - t0 > a.timeout_clocks) {
                        if (lane == 0) {
                            s_status = 3;
                            a.mb->dev_error = 1;
                            __threadfence_system();
                        }
                        break;
                    }
                }
            }
            __syncthreads();
            if (s_status != 1) return;
            for (uint32_t idx = tid; idx < (uint32_t)T * (NC0 - 1); idx += SMALL_BLOCK) {
                const uint32_t t = idx / (NC0 - 1), cc = 1 + idx % (NC0 - 1);
                tab[(size_t)(2 * t) * cap + cc] = tab[(size_t)(2 * t) * cap + NC0 + cc];
                tab[(size_t)(2 * t + 1) * cap + cc] = tab[(size_t)(2 * t + 1) * cap + NC0 + cc];
            }
            __syncthreads();
            nc = 1;
            m = NC0;
        }
        if (m == 1) {  // bound values
            if (tid < T) {
                const Fe v = ld(tid, 0);
                st_fe(a.out[tid], 0, v);
                a.mb->finals[tid] = v;
                __threadfence_system();
            }
            __syncthreads();
            if (tid == 0) a.mb->dev_seq = a.base_seq + pubs + 1;
            return;
        }
        Fe acc[NPTS - 1];
#pragma unroll
        for (int p = 0; p < NPTS - 1; ++p) acc[p] = Fd::zero();
        const uint32_t half = m >> 1;
        for (uint32_t j = tid; tid < nthr && j < half; j += nthr) {
            for (int p = 0; p < (KIND == KIND_XYZ ? 1 : a.n_products); ++p) {
                Fe lo[KD], hi[KD];
#pragma unroll
                for (int f = 0; f < KD; ++f) {
                    lo[f] = ld(p * KD + f, j);
                    hi[f] = ld(p * KD + f, j + half);
                }
                eval_direct<F, KIND, D, NPTS, true>(lo, hi, 0, acc);
            }
        }
        publish(acc, NPTS - 1);
        if (s_status != 1) return;
    }
}

// K3/K1 for lists: fold variable 0 of n_tables tables of 2*n_out entries.
struct FoldTablesArgs {
    TabRef in[MAXT];
    TabRef out[MAXT];
    int n_tables;
    uint64_t n_out;
    FixedMul rt;  // multiplication table of the challenge (fr.cuh FixedMul)
};
template <class F>
__global__ void __launch_bounds__(BLOCK) k_fold_tables(const FoldTablesArgs a) {
    const uint64_t step = (uint64_t)gridDim.x * BLOCK;
    const uint64_t total = a.n_out * (uint64_t)a.n_tables;
    for (uint64_t w = (uint64_t)blockIdx.x * BLOCK + threadIdx.x; w < total; w += step) {
        const int t = (int)(w / a.n_out);
        const uint64_t j = w - (uint64_t)t * a.n_out;
        Fe x0 = ld_fe(a.in[t], j), x1 = ld_fe(a.in[t], j + a.n_out);
        st_fe(a.out[t], j, Field<F>::fold_fixed(x0, x1, a.rt));
    }
}

// Last bind of a sumcheck: every table has 2 entries left; fold them, keep the
// value in the table (entry 0) and publish all n_tables values to the mailbox.
template <class F>
__global__ void k_final_bind(const FoldTablesArgs a, Fe* out, volatile unsigned int* flag, unsigned int seq) {
    const int t = threadIdx.x;
    if (t < a.n_tables) {
        Fe v = Field<F>::fold_fixed(ld_fe(a.in[t], 0), ld_fe(a.in[t], 1), a.rt);
        st_fe(a.out[t], 0, v);
        out[t] = v;
        __threadfence_system();
    }
    __syncthreads();
    if (t == 0 && flag) *flag = seq;
}

// K3: bind the K leading variables in ONE pass (evaluate / multi_partial_evaluate,
// multilinear_polynomial_evaluation.rs:65-91, where every challenge is known up front): thread v
// reads the 2^K entries v + i*n_out, folds variable 0 (the top bit of i) first, writes one entry.
// A chain of single folds moves 96 B per input entry; K = 3 moves 36 B.  In place is safe.
struct MultiFoldArgs {
    TabRef in, out;
    uint64_t n_out;
    FixedMul rt[3];
};
template <class F, int K>
__global__ void __launch_bounds__(BLOCK) k_multifold(const __grid_constant__ MultiFoldArgs a) {
    const uint64_t step = (uint64_t)gridDim.x * BLOCK;
    for (uint64_t v = (uint64_t)blockIdx.x * BLOCK + threadIdx.x; v < a.n_out; v += step) {
        Fe x[1 << K];
#pragma unroll
        for (int i = 0; i < (1 << K); ++i) x[i] = ld_fe(a.in, v + (uint64_t)i * a.n_out);
#pragma unroll
        for (int l = 0; l < K; ++l) {
            const int h = 1 << (K - 1 - l);
#pragma unroll
            for (int i = 0; i < h; ++i) x[i] = Field<F>::fold_fixed(x[i], x[i + h], a.rt[l]);
        }
        st_fe(a.out, v, x[0]);
    }
}

// K1: partial_evaluate(bit, v) (multilinear_polynomial_evaluation.rs:52-63).
// `shift` = n_vars - 1 - bit: output index v pairs inputs insert_bit(v, shift)
// and that | (1 << shift).
template <class F>
__global__ void __launch_bounds__(BLOCK) k_fold(TabRef in, TabRef out, uint64_t n_out, uint32_t shift, const FixedMul rt) {
    const uint64_t step = (uint64_t)gridDim.x * BLOCK;
    const uint64_t lowmask = (1ull << shift) - 1;
    for (uint64_t v = (uint64_t)blockIdx.x * BLOCK + threadIdx.x; v < n_out; v += step) {
        const uint64_t i0 = ((v >> shift) << (shift + 1)) | (v & lowmask);
        const uint64_t i1 = i0 | (1ull << shift);
        st_fe(out, v, Field<F>::fold_fixed(ld_fe(in, i0), ld_fe(in, i1), rt));
    }
}

// K6: layout conversion.  AoS = ark-ff Vec<Fp>: 32 bytes per element.
// conv: 0 = none, 1 = to Montgomery (input canonical), 2 = from Montgomery.
template <class F>
__device__ __forceinline__ Fe apply_conv(const Fe& v, int conv) {
    if (conv == 1) return Field<F>::to_mont(v);
    if (conv == 2) return Field<F>::from_mont(v);
    return v;
}
// Element i of the planar table <- AoS element (first + i*stride_elems).
template <class F>
__global__ void __launch_bounds__(BLOCK) k_aos_to_planar(const uint4* __restrict__ aos, TabRef out, uint64_t n,
                                                        uint64_t first, uint64_t stride_elems, int conv) {
    const uint64_t step = (uint64_t)gridDim.x * BLOCK;
    for (uint64_t i = (uint64_t)blockIdx.x * BLOCK + threadIdx.x; i < n; i += step) {
        const uint64_t s = first + i * stride_elems;
        uint4 a = aos[2 * s], b = aos[2 * s + 1];
        Fe v;
        v.l[0] = a.x; v.l[1] = a.y; v.l[2] = a.z; v.l[3] = a.w;
        v.l[4] = b.x; v.l[5] = b.y; v.l[6] = b.z; v.l[7] = b.w;
        st_fe(out, i, apply_conv<F>(v, conv));
    }
}
template <class F>
__global__ void __launch_bounds__(BLOCK) k_planar_to_aos(TabRef in, uint4* __restrict__ aos, uint64_t n, int conv) {
    const uint64_t step = (uint64_t)gridDim.x * BLOCK;
    for (uint64_t i = (uint64_t)blockIdx.x * BLOCK + threadIdx.x; i < n; i += step) {
        Fe v = apply_conv<F>(ld_fe(in, i), conv);
        aos[2 * i] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
        aos[2 * i + 1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
    }
}
// Interleave G per-rank shards (each n_local entries, planar, concatenated in
// `gathered` rank-major) into the global order i = local*G + rank (C2).
template <class F>
__global__ void __launch_bounds__(BLOCK) k_interleave_shards(TabRef gathered, uint64_t shard_pitch, TabRef out,
                                                            uint64_t n_local, uint32_t log2g) {
    const uint64_t step = (uint64_t)gridDim.x * BLOCK;
    const uint64_t n = n_local << log2g;
    for (uint64_t i = (uint64_t)blockIdx.x * BLOCK + threadIdx.x; i < n; i += step) {
        const uint64_t rank = i & ((1ull << log2g) - 1), loc = i >> log2g;
        // each shard is its own planar block: plane 0 at rank*pitch, plane 1 at rank*pitch + n_local
        TabRef src;
        src.base = gathered.base + rank * shard_pitch;
        src.stride = n_local;
        st_fe(out, i, ld_fe(src, loc));
    }
}

// Synthetic table (same rule as oracle synth_entry): canonical limbs
// from SplitMix64, top limb masked, converted to Montgomery form on the fly.
__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
template <class F>
__global__ void __launch_bounds__(BLOCK) k_generate(TabRef out, uint64_t n, uint64_t seed, uint64_t table,
                                                   uint64_t first, uint64_t stride_elems) {
    const uint64_t step = (uint64_t)gridDim.x * BLOCK;
    const uint64_t base0 = splitmix64(seed ^ (table << 48));
    for (uint64_t i = (uint64_t)blockIdx.x * BLOCK + threadIdx.x; i < n; i += step) {
        const uint64_t gi = first + i * stride_elems;
        const uint64_t base = base0 + 4 * gi;
        Fe v;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            uint64_t w = splitmix64(base + (uint64_t)k);
            if (k == 3) w &= (1ull << (F::ID == 2 ? 62 : 61)) - 1;
            v.l[2 * k] = (uint32_t)w;
            v.l[2 * k + 1] = (uint32_t)(w >> 32);
        }
        st_fe(out, i, Field<F>::to_mont(v));
    }
}

// K4: elementwise.  op: 0 add, 1 sub, 2 mul (Add/Sub/Mul impls :113-156).
template <class F>
__global__ void __launch_bounds__(BLOCK) k_vec_op(TabRef x, TabRef y, TabRef out, uint64_t n, int op) {
    const uint64_t step = (uint64_t)gridDim.x * BLOCK;
    for (uint64_t i = (uint64_t)blockIdx.x * BLOCK + threadIdx.x; i < n; i += step) {
        Fe a = ld_fe(x, i), b = ld_fe(y, i), r;
        if (op == 0) r = Field<F>::add(a, b);
        else if (op == 1) r = Field<F>::sub(a, b);
        else r = Field<F>::mul(a, b);
        st_fe(out, i, r);
    }
}
// out = alpha*x (+ beta*y when y.base != nullptr): scale (:93-97) and the
// alpha/beta merge of gkr_protocol.rs:277-281.
template <class F>
__global__ void __launch_bounds__(BLOCK) k_axpby(TabRef x, TabRef y, TabRef out, uint64_t n, Fe alpha, Fe beta) {
    const uint64_t step = (uint64_t)gridDim.x * BLOCK;
    for (uint64_t i = (uint64_t)blockIdx.x * BLOCK + threadIdx.x; i < n; i += step) {
        Fe r = Field<F>::mul(alpha, ld_fe(x, i));
        if (y.base) r = Field<F>::add(r, Field<F>::mul(beta, ld_fe(y, i)));
        st_fe(out, i, r);
    }
}
// K5: out[i*nb + j] = a[i] (op) b[j]  (tensor_add_mul_polynomials :99-111)
template <class F>
__global__ void __launch_bounds__(BLOCK) k_tensor(TabRef x, TabRef y, TabRef out, uint64_t na, uint64_t nb, int op) {
    const uint64_t step = (uint64_t)gridDim.x * BLOCK;
    for (uint64_t w = (uint64_t)blockIdx.x * BLOCK + threadIdx.x; w < na * nb; w += step) {
        Fe a = ld_fe(x, w / nb), b = ld_fe(y, w % nb);
        st_fe(out, w, op == 0 ? Field<F>::add(a, b) : Field<F>::mul(a, b));
    }
}

// K10: one circuit layer (Circuit::evaluate, gkr_circuit.rs:127-143):
// out[g] = in[2g] (op[g]) in[2g+1].
template <class F>
__global__ void __launch_bounds__(BLOCK) k_layer_eval(TabRef in, TabRef out, const uint8_t* __restrict__ ops,
                                                     uint64_t n_gates) {
    const uint64_t step = (uint64_t)gridDim.x * BLOCK;
    for (uint64_t g = (uint64_t)blockIdx.x * BLOCK + threadIdx.x; g < n_gates; g += step) {
        Fe a = ld_fe(in, 2 * g), b = ld_fe(in, 2 * g + 1);
        st_fe(out, g, ops[g] ? Field<F>::mul(a, b) : Field<F>::add(a, b));
    }
}

// Dense add_i / mul_i indicator (Layer::get_add_mul_i, gkr_circuit.rs:39-104) over index a||b||c with widths
// (w, w+1, w+1), w = log2(G) (or (1,1,1) for a single gate): a one at (g, 2g, 2g+1) for every gate g of
// the requested operation.  `out` must be zero-filled.  Only feasible for small layers (2^(3w+2) entries);
// the prover itself uses the sparse two-phase form (K12-K14).
template <class F>
__global__ void __launch_bounds__(BLOCK) k_add_mul_i(const uint8_t* __restrict__ ops, uint32_t n_gates, int op, int w, TabRef out) {
    const uint32_t g = blockIdx.x * BLOCK + threadIdx.x;
    if (g >= n_gates || ops[g] != op) return;
    const int wb = n_gates == 1 ? 1 : w + 1;
    const uint64_t idx = ((((uint64_t)g << wb) | (uint64_t)(2 * g)) << wb) | (uint64_t)(2 * g + 1);
    st_fe(out, idx, Field<F>::one());
}

// K11: eq(r, .) in split form.  r has n challenges (variable 0 = MSB).  The
// table over the first n_hi variables goes to `hi` (2^n_hi entries), the one
// over the remaining n - n_hi to `lo`; eq(r, x) = hi[x >> n_lo] * lo[x & mask].
// One thread per entry, <= 20 multiplications each; both tables are tiny.
template <class F>
__global__ void __launch_bounds__(BLOCK) k_eq_split(const ChalList r, int n, int n_hi, TabRef hi, TabRef lo) {
    typedef Field<F> Fd;
    const int n_lo = n - n_hi;
    const uint64_t nh = 1ull << n_hi, nl = 1ull << n_lo;
    const uint64_t step = (uint64_t)gridDim.x * BLOCK;
    for (uint64_t w = (uint64_t)blockIdx.x * BLOCK + threadIdx.x; w < nh + nl; w += step) {
        const bool is_hi = w < nh;
        const uint64_t x = is_hi ? w : w - nh;
        const int cnt = is_hi ? n_hi : n_lo, off = is_hi ? 0 : n_hi;
        Fe v = Fd::one();
        for (int k = 0; k < cnt; ++k) {
            const Fe rk = r.r[off + k];
            const bool bit = (x >> (cnt - 1 - k)) & 1;
            v = Fd::mul(v, bit ? rk : Fd::sub(Fd::one(), rk));
        }
        st_fe(is_hi ? hi : lo, x, v);
    }
}
template <class F>
__device__ __forceinline__ Fe eq_lookup(const TabRef& hi, const TabRef& lo, int n_lo, uint64_t x) {
    return Field<F>::mul(ld_fe(hi, x >> n_lo), ld_fe(lo, x & ((1ull << n_lo) - 1)));
}

// K12: phase-1 tables of the two-phase GKR layer sumcheck (sparse restatement
// of get_fbc_poly / get_folded_fbc_poly, gkr_protocol.rs:243-292).
//   coef[g] = first layer : eq((r0), g)                       (one `a` variable)
//             otherwise   : alpha*eq(r_b, g) + beta*eq(r_c, g)
//   gate g reads wires b = 2g, c = 2g+1 (gkr_circuit.rs:76-78):
//   Add: H1[b] = coef, HA2[b] = coef*W[c];   Mul: H1[b] = coef*W[c], HA2[b] = 0;
//   odd b: zero.  Round polynomial of phase 1: sum_b W(b)*H1(b) + HA2(b).
struct GkrP1Args {
    TabRef W, H1, HA2, coef;
    TabRef eb_hi, eb_lo, ec_hi, ec_lo;  // split eq tables of r_b, r_c (unused for the first layer)
    int n_lo;
    const uint8_t* ops;
    uint64_t n_gates;
    int first_layer;
    Fe r0, alpha, beta;
};
template <class F>
__global__ void __launch_bounds__(BLOCK) k_gkr_phase1(const GkrP1Args a) {
    typedef Field<F> Fd;
    const uint64_t step = (uint64_t)gridDim.x * BLOCK;
    for (uint64_t g = (uint64_t)blockIdx.x * BLOCK + threadIdx.x; g < a.n_gates; g += step) {
        Fe c;
        if (a.first_layer) {
            c = g ? a.r0 : Fd::sub(Fd::one(), a.r0);
        } else {
            Fe eb = eq_lookup<F>(a.eb_hi, a.eb_lo, a.n_lo, g);
            Fe ec = eq_lookup<F>(a.ec_hi, a.ec_lo, a.n_lo, g);
            c = Fd::add(Fd::mul(a.alpha, eb), Fd::mul(a.beta, ec));
        }
        st_fe(a.coef, g, c);
        Fe wc = ld_fe(a.W, 2 * g + 1);
        Fe cw = Fd::mul(c, wc);
        const bool is_mul = a.ops[g] != 0;
        st_fe(a.H1, 2 * g, is_mul ? cw : c);
        st_fe(a.HA2, 2 * g, is_mul ? Fd::zero() : cw);
        st_fe(a.H1, 2 * g + 1, Fd::zero());
        st_fe(a.HA2, 2 * g + 1, Fd::zero());
    }
}
// K13: phase-2 tables, b bound to u: with v = coef[g]*eq(u, 2g), Wu = W(u):
//   Add: C[c] = v, D[c] = Wu*v;   Mul: C[c] = Wu*v, D[c] = 0;   even c: zero.
// Round polynomial of phase 2: sum_c W(c)*C(c) + D(c).
struct GkrP2Args {
    TabRef C, D, coef;
    TabRef eu_hi, eu_lo;
    int n_lo;
    const uint8_t* ops;
    uint64_t n_gates;
    Fe Wu;
};
template <class F>
__global__ void __launch_bounds__(BLOCK) k_gkr_phase2(const GkrP2Args a) {
    typedef Field<F> Fd;
    const uint64_t step = (uint64_t)gridDim.x * BLOCK;
    for (uint64_t g = (uint64_t)blockIdx.x * BLOCK + threadIdx.x; g < a.n_gates; g += step) {
        Fe v = Fd::mul(ld_fe(a.coef, g), eq_lookup<F>(a.eu_hi, a.eu_lo, a.n_lo, 2 * g));
        Fe wv = Fd::mul(a.Wu, v);
        const bool is_mul = a.ops[g] != 0;
        st_fe(a.C, 2 * g + 1, is_mul ? wv : v);
        st_fe(a.D, 2 * g + 1, is_mul ? Fd::zero() : wv);
        st_fe(a.C, 2 * g, Fd::zero());
        st_fe(a.D, 2 * g, Fd::zero());
    }
}

// ----------------------------------------------------------------- general wiring (extension, SURVEY F8 ii)
// Gate g of a layer reads wires in1[g], in2[g] of the layer below (width W, any power of two).  The phase tables
// are gathered through CSR lists of the gates by first / second input, made on the host at circuit creation:
//   phase 1: H1[b]  = sum_{g: in1[g]=b} coef[g] * (Add ? 1 : W[in2[g]]),  HA2[b] = sum_{g Add, in1[g]=b} coef[g]*W[in2[g]]
//   phase 2: C[c]   = sum_{g: in2[g]=c} coef[g]*eq(u,in1[g]) * (Add ? 1 : W(u)),  D[c] = sum_{g Add, in2[g]=c} coef[g]*eq(u,in1[g])*W(u)
// with coef[g] = alpha*eq(ra1, g) + beta*eq(ra2, g) (output layer: eq(r0, g)).  With in1 = 2g, in2 = 2g+1 these are
// exactly the tables of k_gkr_phase1 / k_gkr_phase2.
struct WiredCoef {
    TabRef a1_hi, a1_lo, a2_hi, a2_lo;
    int n_lo;
    int two;  // 0: coef = eq(ra1, g) (output layer); 1: alpha*eq(ra1,g) + beta*eq(ra2,g)
    Fe alpha, beta;
};
template <class F>
__device__ __forceinline__ Fe wired_coef(const WiredCoef& w, uint64_t g) {
    typedef Field<F> Fd;
    Fe e1 = eq_lookup<F>(w.a1_hi, w.a1_lo, w.n_lo, g);
    if (!w.two) return e1;
    Fe e2 = eq_lookup<F>(w.a2_hi, w.a2_lo, w.n_lo, g);
    return Fd::add(Fd::mul(w.alpha, e1), Fd::mul(w.beta, e2));
}
// Each phase is two launches so that the multiplications run one gate per thread (no divergence) and the
// per-wire gather -- whose trip count varies with the wire's fan-out -- only adds:
//   k_gkr_w_gates1: coef[g], tmp[g] = coef[g]*W[in2[g]]          k_gkr_w_phase1: H1, HA2 from coef/tmp via lst1
//   k_gkr_w_gates2: tmp[g] = coef[g]*eq(u, in1[g])                k_gkr_w_phase2: sA, sM from tmp via lst2, then
//                                                                  C = sA + W(u)*sM, D = W(u)*sA
struct GkrW1Args {
    TabRef W, H1, HA2, coef, tmp;
    WiredCoef wc;
    const uint8_t* ops;
    const uint32_t *in2, *off1, *lst1;
    uint64_t width, n_gates;
};
template <class F>
__global__ void __launch_bounds__(BLOCK) k_gkr_w_gates1(const GkrW1Args a) {
    typedef Field<F> Fd;
    const uint64_t step = (uint64_t)gridDim.x * BLOCK;
    for (uint64_t g = (uint64_t)blockIdx.x * BLOCK + threadIdx.x; g < a.n_gates; g += step) {
        const Fe c = wired_coef<F>(a.wc, g);
        st_fe(a.coef, g, c);
        st_fe(a.tmp, g, Fd::mul(c, ld_fe(a.W, a.in2[g])));
    }
}
template <class F>
__global__ void __launch_bounds__(BLOCK) k_gkr_w_phase1(const GkrW1Args a) {
    typedef Field<F> Fd;
    const uint64_t step = (uint64_t)gridDim.x * BLOCK;
    for (uint64_t b = (uint64_t)blockIdx.x * BLOCK + threadIdx.x; b < a.width; b += step) {
        Fe h1 = Fd::zero(), ha2 = Fd::zero();
        for (uint32_t e = a.off1[b]; e < a.off1[b + 1]; ++e) {
            const uint32_t g = a.lst1[e];
            const Fe cw = ld_fe(a.tmp, g);
            if (a.ops[g]) h1 = Fd::add(h1, cw);
            else {
                h1 = Fd::add(h1, ld_fe(a.coef, g));
                ha2 = Fd::add(ha2, cw);
            }
        }
        st_fe(a.H1, b, h1);
        st_fe(a.HA2, b, ha2);
    }
}
struct GkrW2Args {
    TabRef C, D, coef, tmp;
    TabRef eu_hi, eu_lo;
    int n_lo;
    const uint8_t* ops;
    const uint32_t *in1, *off2, *lst2;
    uint64_t width, n_gates;
    Fe Wu;
};
template <class F>
__global__ void __launch_bounds__(BLOCK) k_gkr_w_gates2(const GkrW2Args a) {
    const uint64_t step = (uint64_t)gridDim.x * BLOCK;
    for (uint64_t g = (uint64_t)blockIdx.x * BLOCK + threadIdx.x; g < a.n_gates; g += step)
        st_fe(a.tmp, g, Field<F>::mul(ld_fe(a.coef, g), eq_lookup<F>(a.eu_hi, a.eu_lo, a.n_lo, a.in1[g])));
}
template <class F>
__global__ void __launch_bounds__(BLOCK) k_gkr_w_phase2(const GkrW2Args a) {
    typedef Field<F> Fd;
    const uint64_t step = (uint64_t)gridDim.x * BLOCK;
    for (uint64_t cc = (uint64_t)blockIdx.x * BLOCK + threadIdx.x; cc < a.width; cc += step) {
        Fe sA = Fd::zero(), sM = Fd::zero();
        for (uint32_t e = a.off2[cc]; e < a.off2[cc + 1]; ++e) {
            const uint32_t g = a.lst2[e];
            const Fe v = ld_fe(a.tmp, g);
            if (a.ops[g]) sM = Fd::add(sM, v);
            else sA = Fd::add(sA, v);
        }
        st_fe(a.C, cc, Fd::add(sA, Fd::mul(a.Wu, sM)));
        st_fe(a.D, cc, Fd::mul(a.Wu, sA));
    }
}
struct GkrWWiringArgs {
    WiredCoef wc;
    TabRef eu_hi, eu_lo, ew_hi, ew_lo;
    int n_lo_w;
    const uint8_t* ops;
    const uint32_t *in1, *in2;
    uint64_t n_gates;
    FinishArgs fin;
};
template <class F>
__global__ void __launch_bounds__(BLOCK) k_gkr_w_wiring(const GkrWWiringArgs a) {
    typedef Field<F> Fd;
    Fe acc[2] = {Fd::zero(), Fd::zero()};
    const uint64_t step = (uint64_t)gridDim.x * BLOCK;
    for (uint64_t g = (uint64_t)blockIdx.x * BLOCK + threadIdx.x; g < a.n_gates; g += step) {
        const Fe c = wired_coef<F>(a.wc, g);
        const Fe t = Fd::mul(c, Fd::mul(eq_lookup<F>(a.eu_hi, a.eu_lo, a.n_lo_w, a.in1[g]), eq_lookup<F>(a.ew_hi, a.ew_lo, a.n_lo_w, a.in2[g])));
        if (a.ops[g]) acc[1] = Fd::add(acc[1], t);
        else acc[0] = Fd::add(acc[0], t);
    }
    finish_round<F, 2>(acc, a.fin);
}
template <class F>
__global__ void __launch_bounds__(BLOCK) k_layer_eval_w(TabRef in, TabRef out, const uint8_t* __restrict__ ops, const uint32_t* __restrict__ in1,
                                                       const uint32_t* __restrict__ in2, uint64_t n_gates) {
    const uint64_t step = (uint64_t)gridDim.x * BLOCK;
    for (uint64_t g = (uint64_t)blockIdx.x * BLOCK + threadIdx.x; g < n_gates; g += step) {
        Fe x = ld_fe(in, in1[g]), y = ld_fe(in, in2[g]);
        st_fe(out, g, ops[g] ? Field<F>::mul(x, y) : Field<F>::add(x, y));
    }
}

// Element `idx` of each listed table -> out[t] (final bound values, openings).
struct GatherArgs {
    TabRef t[MAXT];
    int n;
    uint64_t idx;
    Fe* out;
};
static __global__ void k_gather_elems(const GatherArgs a) {
    int t = threadIdx.x;
    if (t < a.n) a.out[t] = ld_fe(a.t[t], a.idx);
}

// Verifier side of a GKR layer (get_verifier_claim / get_folded_verifier_claim,
// gkr_protocol.rs:294-341) in O(G): the wiring predicates at the random point,
//   a_r = sum_{g: Add} coef[g]*eq(u,2g)*eq(w,2g+1),   m_r likewise for Mul,
// with coef as in k_gkr_phase1.  result[0] = a_r, result[1] = m_r.
struct GkrWiringArgs {
    TabRef eb_hi, eb_lo, ec_hi, ec_lo;  // split eq of the previous r_b, r_c
    int n_lo_a;
    TabRef eu_hi, eu_lo, ew_hi, ew_lo;  // split eq of this layer's r_b (u) and r_c (w)
    int n_lo_w;
    const uint8_t* ops;
    uint64_t n_gates;
    int first_layer;
    Fe r0, alpha, beta;
    FinishArgs fin;
};
template <class F>
__global__ void __launch_bounds__(BLOCK) k_gkr_wiring(const GkrWiringArgs a) {
    typedef Field<F> Fd;
    Fe acc[2] = {Fd::zero(), Fd::zero()};
    const uint64_t step = (uint64_t)gridDim.x * BLOCK;
    for (uint64_t g = (uint64_t)blockIdx.x * BLOCK + threadIdx.x; g < a.n_gates; g += step) {
        Fe c;
        if (a.first_layer) {
            c = g ? a.r0 : Fd::sub(Fd::one(), a.r0);
        } else {
            Fe eb = eq_lookup<F>(a.eb_hi, a.eb_lo, a.n_lo_a, g);
            Fe ec = eq_lookup<F>(a.ec_hi, a.ec_lo, a.n_lo_a, g);
            c = Fd::add(Fd::mul(a.alpha, eb), Fd::mul(a.beta, ec));
        }
        Fe t = Fd::mul(c, Fd::mul(eq_lookup<F>(a.eu_hi, a.eu_lo, a.n_lo_w, 2 * g),
                                  eq_lookup<F>(a.ew_hi, a.ew_lo, a.n_lo_w, 2 * g + 1)));
        if (a.ops[g]) acc[1] = Fd::add(acc[1], t);
        else acc[0] = Fd::add(acc[0], t);
    }
    finish_round<F, 2>(acc, a.fin);
}

// ------------------------------------------------------- microbenchmarks
// Register-resident multiplier throughput (fills the IMAD-roofline number the
// driver does not measure): each thread runs `iters` dependent products on
// ILP independent chains.
#include "ntt_merkle.cuh"

template <class F, int ILP, bool SPLIT>
__global__ void __launch_bounds__(BLOCK) k_bench_mul(Fe* out, uint32_t iters, Fe seed) {
    Fe x[ILP], y = seed;
#pragma unroll
    for (int k = 0; k < ILP; ++k) {
        x[k] = seed;
        x[k].l[0] += threadIdx.x + k;
    }
    for (uint32_t it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < ILP; ++k) x[k] = SPLIT ? Field<F>::mul_split(x[k], y) : Field<F>::mul(x[k], y);
    }
    Fe s = x[0];
#pragma unroll
    for (int k = 1; k < ILP; ++k) s = Field<F>::add(s, x[k]);
    if (s.l[0] == 0x12345678u && s.l[7] == 0x9abcdef0u) out[blockIdx.x * BLOCK + threadIdx.x] = s;  // keep the loop alive
}
// Raw multiply-pipe rate: MODE 0 = IMAD (32-bit mad.lo), 1 = IMAD.HI, 2 = mad.wide (64-bit accumulate,
// no carry), 3 = the multiplier's own chain: mul.wide + addc.cc.u64 -> IMAD.WIDE.U32.X with predicate carry
template <int MODE>
__global__ void __launch_bounds__(BLOCK) k_bench_imad(uint64_t* out, uint32_t iters, uint32_t a, uint32_t b) {
    uint64_t acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = threadIdx.x + k;
    uint32_t x = a + threadIdx.x, y = b;
    for (uint32_t it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (MODE == 0) {
                    uint32_t lo = (uint32_t)acc[k];
                    asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(lo) : "r"(x), "r"(y));
                    acc[k] = lo;
                } else if (MODE == 1) {
                    uint32_t lo = (uint32_t)acc[k];
                    asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(lo) : "r"(x), "r"(y));
                    acc[k] = lo;
                } else if (MODE == 2) {
                    asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[k]) : "r"(x), "r"(y));
                } else {
                    acc[k] = (k & 3) == 0 ? madw_cc(x, y, acc[k]) : madwc_cc(x, y, acc[k]);
                }
            }
        }
    }
    uint64_t s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += acc[k];
    if (s == 0x123456789abcdef0ull) out[blockIdx.x * BLOCK + threadIdx.x] = s;
}

}  // namespace zkb
