// tcfold.cuh -- binding a challenge on the 5th-generation tensor cores (tcgen05.mma kind::i8, sm_100a).
//
// A fold with a challenge r that is fixed for a whole launch,
//     new = a + r (b - a) = a (1 - r) + b r            (multilinear_polynomial_evaluation.rs:52-63),
// is linear in the BYTES of a and b: with T1_i = (1 - r) 2^(8 i + 32) mod p and T2_i = r 2^(8 i + 32) mod p,
//     S = sum_i a_i T1_i + b_i T2_i  ==  (a (1 - r) + b r) 2^32   (mod p),      S < 64 * 255 * p < 2^14 p,
// and S is two u8 x u8 -> s32 matrix products [128 entries x 32 bytes] . [32 x 32] accumulated in TMEM.  The rows of
// the A operand are table entries exactly as they lie in the planar layout (plane 0 = bytes 0..15, plane 1 = bytes
// 16..31 of every entry), which IS the K-major no-swizzle operand layout of the instruction (8-row x 16-byte core
// matrices: SBO = 128 bytes, LBO = the distance of the two planes), so a tile goes from HBM to shared memory by TMA bulk
// copies and from there into the tensor core without passing through a register.  The CUDA cores only carry the 32
// column sums (< 2^22 each) into 9 limbs, run ONE 32-bit Montgomery row (division by 2^32: 8 wide multiplies instead
// of the 82 of Field::fold_fixed) and subtract p once: the result is the canonical residue, bit-identical to the
// CUDA-core fold.
//
// Contents: tcgen05 / TMEM / UMMA-descriptor primitives; tc_fold_finish (column sums -> canonical residue); k_tc_fold (the
// stand-alone check, tools/tcfold.cu); round_pass_tc (fused round pass, 2 factors; used by k_sc_fold_eval_tc and k_sc_tail<TC>);
// k_sc_eval_tc (round 0 of 2 factors as one Gram matrix); GramAcc2 + round_pass_tc_gram + k_sc_eval_gram (3 factors: raw
// partial products through Gram tiles); k_multifold_tc (evaluate); k_tc_probe (co-residency of the plain-launched persistent
// kernel).  DESIGN.md section 6a has the derivations and the measurements.
#pragma once
// (included by kernels.cuh after the round-pass building blocks and before the persistent kernel)

namespace zkb {

template <class F>
ZK_HD void tc_fold_mats(const Fe& r_mont, TcFoldMats* out) {
    typedef Field<F> Fd;
    const Fe one_minus_r = Fd::sub(Fd::one(), r_mont);
    Fe v = Fd::zero();
    v.l[1] = 1;  // 2^32 as a plain integer: mul(x R, v) = x v mod p
    for (int i = 0; i < 32; ++i) {
        const Fe t1 = Fd::mul(one_minus_r, v), t2 = Fd::mul(r_mont, v);
        for (int n = 0; n < 32; ++n) {
            const int off = (i / 16) * 512 + n * 16 + i % 16;
            out->b[0][off] = (uint8_t)(t1.l[n / 4] >> (8 * (n % 4)));
            out->b[1][off] = (uint8_t)(t2.l[n / 4] >> (8 * (n % 4)));
        }
        for (int k = 0; k < 8; ++k) v = Fd::add(v, v);
    }
}

#if defined(__CUDACC__)
// ---- tcgen05 / TMEM primitives (PTX; the encodings follow cute/arch/mma_sm100_desc.hpp) ----
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// shared-memory matrix descriptor, K-major, no swizzle: start address, LBO (distance of the two 16-byte K chunks),
// SBO (distance of 8-row groups), version 1 (Blackwell)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) |
           (1ull << 46);
}
// instruction descriptor: D = s32 (2 << 4), A = B = u8 (0), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
constexpr uint32_t TC_IDESC_U8_M128_N32 = (2u << 4) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// this thread's row (TMEM lane) of 32 consecutive 32-bit columns
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* c) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
        "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(c[0]), "=r"(c[1]), "=r"(c[2]), "=r"(c[3]), "=r"(c[4]), "=r"(c[5]), "=r"(c[6]), "=r"(c[7]), "=r"(c[8]), "=r"(c[9]), "=r"(c[10]),
          "=r"(c[11]), "=r"(c[12]), "=r"(c[13]), "=r"(c[14]), "=r"(c[15]), "=r"(c[16]), "=r"(c[17]), "=r"(c[18]), "=r"(c[19]), "=r"(c[20]),
          "=r"(c[21]), "=r"(c[22]), "=r"(c[23]), "=r"(c[24]), "=r"(c[25]), "=r"(c[26]), "=r"(c[27]), "=r"(c[28]), "=r"(c[29]), "=r"(c[30]),
          "=r"(c[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

template <class F>
__device__ __forceinline__ Fe tc_mont_row(uint32_t* s);
__device__ __forceinline__ void tcg_assemble(const uint32_t* c, uint32_t* v);
// 32 column sums (byte weights 2^(8 n), each < 2^22) -> the canonical residue S 2^-32 mod p
template <class F>
__device__ __forceinline__ Fe tc_fold_finish(const uint32_t* c) {
    uint32_t lo[8], hi[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const uint32_t t01 = c[4 * k] + (c[4 * k + 1] << 8), t23 = c[4 * k + 2] + (c[4 * k + 3] << 8);  // < 2^31
        const uint64_t w = (uint64_t)t01 + ((uint64_t)t23 << 16);                                       // limb k + its overflow (< 2^15)
        lo[k] = lo32(w);
        hi[k] = hi32(w);
    }
    uint32_t s[9];
    s[0] = lo[0];
    s[1] = add_cc(lo[1], hi[0]);
#pragma unroll
    for (int k = 2; k < 8; ++k) s[k] = addc_cc(lo[k], hi[k - 1]);
    s[8] = addc(hi[7], 0u);
    return tc_mont_row<F>(s);
}
// (S + m p) >> 32 for a 9-limb S < 2^31 p, then one conditional subtraction: the canonical residue S 2^-32 mod p
template <class F>
__device__ __forceinline__ Fe tc_mont_row(uint32_t* s) {
    // one Montgomery row: (S + m p) >> 32 < p (1 + 2^-18).  m p is taken as even-limb products chained onto S and
    // odd-limb products computed stand-alone (no carry, full-rate), merged by the final addition.
    const uint32_t m = mul_lo(s[0], F::INV);
    uint64_t e0 = madw_cc(F::P(0), m, pack64(s[0], s[1]));
    uint64_t e1 = madwc_cc(F::P(2), m, pack64(s[2], s[3]));
    uint64_t e2 = madwc_cc(F::P(4), m, pack64(s[4], s[5]));
    uint64_t e3 = madwc_cc(F::P(6), m, pack64(s[6], s[7]));
    s[8] = addc(s[8], 0u);
    const uint64_t o0 = mul_wide(F::P(1), m), o1 = mul_wide(F::P(3), m), o2 = mul_wide(F::P(5), m), o3 = mul_wide(F::P(7), m);
    Fe r;
    r.l[0] = add_cc(hi32(e0), lo32(o0));
    r.l[1] = addc_cc(lo32(e1), hi32(o0));
    r.l[2] = addc_cc(hi32(e1), lo32(o1));
    r.l[3] = addc_cc(lo32(e2), hi32(o1));
    r.l[4] = addc_cc(hi32(e2), lo32(o2));
    r.l[5] = addc_cc(lo32(e3), hi32(o2));
    r.l[6] = addc_cc(hi32(e3), lo32(o3));
    r.l[7] = addc(s[8], hi32(o3));
    return Field<F>::reduce_once(r);
}

// ---- stand-alone fold kernel (partial_evaluate of variable 0): one tile of 128 output entries per step ----
constexpr int TC_FOLD_THREADS = 128;
constexpr int TC_TILE_BYTES = 2 * 128 * 16;                        // one operand tile: two planes of 128 x 16 bytes
constexpr int TC_FOLD_SMEM = 2 * TC_TILE_BYTES + 2048 + 64;       // a, b, the two B matrices, barriers + TMEM address
template <class F>
__global__ void __launch_bounds__(TC_FOLD_THREADS) k_tc_fold(TabRef in, TabRef out, uint64_t n_out, const TcFoldMats* __restrict__ mats) {
    extern __shared__ __align__(128) uint8_t tc_sm[];
    uint8_t* sm_a = tc_sm;
    uint8_t* sm_b = tc_sm + TC_TILE_BYTES;
    uint8_t* sm_m = tc_sm + 2 * TC_TILE_BYTES;
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(tc_sm + 2 * TC_TILE_BYTES + 2048);
    uint64_t* bar_mma = bar_full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_full + 2);
    const uint32_t tid = threadIdx.x, warp = tid >> 5;
    for (uint32_t i = tid; i < 2048 / 16; i += TC_FOLD_THREADS) reinterpret_cast<uint4*>(sm_m)[i] = reinterpret_cast<const uint4*>(mats)[i];
    if (tid == 0) {
        mbar_init(bar_full, 1);
        mbar_init(bar_mma, 1);
        fence_barrier_init();
    }
    fence_proxy_async();  // the matrices were written through the generic proxy, the tensor core reads through the async one
    if (warp == 0) tmem_alloc(tmem_slot, 32);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint64_t tiles = (n_out + 127) / 128;
    uint32_t phase = 0;
    for (uint64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const uint64_t j0 = tile * 128;
        const uint32_t rows = (uint32_t)(n_out - j0 < 128 ? n_out - j0 : 128);
        if (tid == 0) {
            mbar_expect_tx(bar_full, rows * 64u);
            bulk_g2s(sm_a, in.base + j0, rows * 16u, bar_full);
            bulk_g2s(sm_a + 2048, in.base + in.stride + j0, rows * 16u, bar_full);
            bulk_g2s(sm_b, in.base + j0 + n_out, rows * 16u, bar_full);
            bulk_g2s(sm_b + 2048, in.base + in.stride + j0 + n_out, rows * 16u, bar_full);
        }
#ifdef TC_TIMING
        const long long t0 = clock64();
#endif
        mbar_wait(bar_full, phase);
        tc_fence_after();
#ifdef TC_TIMING
        const long long t1 = clock64();
#endif
        if (tid == 0) {
            umma_i8(tmem, umma_desc(smem_u32(sm_a), 2048, 128), umma_desc(smem_u32(sm_m), 512, 128), TC_IDESC_U8_M128_N32, 0u);
            umma_i8(tmem, umma_desc(smem_u32(sm_b), 2048, 128), umma_desc(smem_u32(sm_m + 1024), 512, 128), TC_IDESC_U8_M128_N32, 1u);
            umma_commit(bar_mma);
        }
        mbar_wait(bar_mma, phase);
        tc_fence_after();
#ifdef TC_TIMING
        const long long t2 = clock64();
#endif
        uint32_t c[32];
        tmem_ld32(tmem + ((warp * 32u) << 16), c);
        tmem_wait_ld();
#ifdef TC_TIMING
        const long long t3 = clock64();
        if (tid == 0 && blockIdx.x == 0 && tile < 8 * (uint64_t)gridDim.x)
            printf("tile %llu: TMA issue->landed %lld, MMA issue->barrier %lld, LDTM %lld cycles\n", (unsigned long long)tile, t1 - t0, t2 - t1, t3 - t2);
#endif
        const Fe v = tc_fold_finish<F>(c);
        if (tid < rows) st_fe(out, j0 + tid, v);
        tc_fence_before();
        __syncthreads();  // every row has been read: the accumulator and the operand tiles may be overwritten
        phase ^= 1u;
    }
    if (warp == 0) tmem_dealloc(tmem, 32);
}

// ---- the fused round pass (fold all tables with the challenge, write them once, accumulate the next round's sums
// from the folded values) with the folds on the tensor cores.  Same ownership as round_pass (kernels.cuh): row r of
// a tile is the quad position j = tile * 128 + r, its "lo" fold is entries (j, j + n_out) -> new[j], its "hi" fold
// (j + half, j + half + n_out) -> new[j + half]; in-place operation stays safe.
// A CTA is two independent warpgroups of 128 threads; a warpgroup walks its tiles unit by unit (unit = one fold of one
// table for the 128 positions of the tile = 8 KiB of operands + one 32-column accumulator):
//   TMA (leader thread):   unit q + NS  -> operand stage q % NS          (after the MMAs of unit q have completed)
//   MMA (leader thread):   unit q + 1   -> TMEM stage (q + 1) % 2         (after every warp has read unit q - 1)
//   all 128 threads:       unit q: TMEM -> registers, carry + Montgomery row, store, product accumulation
// Needs n_out / 2 to be a multiple of 128 and BLOCK == 256.
constexpr int TC_UNIT_BYTES = 2 * TC_TILE_BYTES;  // a tile + b tile
#ifndef ZKB_TC_NT
#define ZKB_TC_NT 2
#endif
// Pipeline depth: NS operand stages (shared memory, 8 KiB each) and NT accumulator stages (TMEM, 32 columns each) per
// warpgroup, chosen so that two CTAs fit one SM next to the parked accumulators of the round sums.
template <int NPTS>
struct TcCfg {
#ifdef ZKB_TC_NS
    static constexpr int NS = ZKB_TC_NS;
#else
    static constexpr int NS = NPTS <= 3 ? 4 : 3;
#endif
    static constexpr int NT = ZKB_TC_NT;      // power of two, <= NS
    static constexpr int tmem_cols = 2 * NT * 32;  // per CTA (a power of two >= 32)
    static_assert(NT <= NS && (NT & (NT - 1)) == 0, "accumulator stages");
};
template <int NPTS>
struct TcRoundSmem {
    typedef TcCfg<NPTS> C;
    static constexpr int stage_bytes = 2 * C::NS * TC_UNIT_BYTES;
    static constexpr int mats_off = stage_bytes;
    static constexpr int bars_off = mats_off + 2048;
    static constexpr int bars_per_wg = (C::NS + 2 * C::NT) * 8;  // full[NS], mma[NT], empty[NT]
    static constexpr int tmem_off = bars_off + 2 * bars_per_wg;
    static constexpr int accs_off = bars_off + 256;
    static constexpr int bytes = accs_off + (NPTS - 1) * ACC_VECS * BLOCK * 16;
    static_assert(2 * bars_per_wg + 4 <= 256, "barrier block");
};
__device__ __forceinline__ void mbar_init_u32(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_u32(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait_u32(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "ZKB_TCW_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra ZKB_TCW_%=;\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void wg_sync(uint32_t wg) {  // named barrier 1 / 2 (immediate ids: a register id makes ptxas reserve all 16)
    if (wg == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
    else asm volatile("bar.sync 2, 128;" ::: "memory");
}

template <class F, int D, int NPTS>
__device__ __forceinline__ void round_pass_tc(const TabRef* __restrict__ in, const TabRef* __restrict__ outp, int n_products, uint64_t n_out,
                                              uint8_t* smb, uint32_t tmem_cta, Fe* out) {
    static_assert(BLOCK == 256, "two warpgroups per CTA");
    typedef TcRoundSmem<NPTS> L;
    constexpr int NS = TcCfg<NPTS>::NS, NT = TcCfg<NPTS>::NT, LOG_NT = NT == 1 ? 0 : (NT == 2 ? 1 : 2);
    const uint32_t wg = threadIdx.x >> 7, r = threadIdx.x & 127u, wq = (threadIdx.x >> 5) & 3u, lane = threadIdx.x & 31u;
    const uint64_t half = n_out >> 1;
    const uint64_t tiles = half >> 7;
    const uint64_t first = (uint64_t)blockIdx.x * 2 + wg, tstride = (uint64_t)gridDim.x * 2;
    const uint32_t upt = 2u * (uint32_t)n_products * D;  // units per tile
    const uint32_t my_tiles = first < tiles ? (uint32_t)((tiles - first + tstride - 1) / tstride) : 0u;
    const uint32_t U = my_tiles * upt;
    const uint32_t st0 = smem_u32(smb) + wg * (NS * TC_UNIT_BYTES);
    const uint32_t mats = smem_u32(smb + L::mats_off);
    const uint32_t b_full = smem_u32(smb + L::bars_off) + wg * L::bars_per_wg, b_mma = b_full + NS * 8, b_empty = b_mma + NT * 8;
    const uint32_t tmem = tmem_cta + wg * (NT * 32u);
    if (r == 0) {  // the barriers live for one pass (phase counting restarts with every pass)
#pragma unroll
        for (int b = 0; b < NS + 2 * NT; ++b) mbar_init_u32(b_full + b * 8, b < NS + NT ? 1u : 4u);
        fence_barrier_init();
    }
    wg_sync(wg);
    // leader state: next unit to load (tile, unit in tile, stage) and next unit to multiply (index, stage, parity)
    uint64_t t_tile = first;
    uint32_t t_u = 0, t_stage = 0, m_q = 0, m_stage = 0, m_par = 0;
    auto issue_tma = [&]() {
        const TabRef& tb = in[t_u >> 1];
        const uint4* g0 = tb.base + t_tile * 128 + ((t_u & 1u) ? half : 0);
        const uint4* g1 = g0 + tb.stride;
        const uint32_t dst = st0 + t_stage * TC_UNIT_BYTES, bar = b_full + t_stage * 8;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)TC_UNIT_BYTES) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(g0), "r"(2048u), "r"(bar) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst + 2048u), "l"(g1), "r"(2048u), "r"(bar) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst + 4096u), "l"(g0 + n_out), "r"(2048u), "r"(bar) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst + 6144u), "l"(g1 + n_out), "r"(2048u), "r"(bar) : "memory");
        if (++t_u == upt) {
            t_u = 0;
            t_tile += tstride;
        }
        t_stage = t_stage + 1 == NS ? 0u : t_stage + 1;
    };
    auto issue_mma = [&]() {
        const uint32_t ts = m_q & (NT - 1);
        mbar_wait_u32(b_full + m_stage * 8, m_par);
        if (m_q >= NT) mbar_wait_u32(b_empty + ts * 8, ((m_q >> LOG_NT) & 1u) ^ 1u);  // every warp has read unit m_q - NT
        tc_fence_after();
        const uint32_t d = tmem + ts * 32u, a0 = st0 + m_stage * TC_UNIT_BYTES;
        umma_i8(d, umma_desc(a0, 2048, 128), umma_desc(mats, 512, 128), TC_IDESC_U8_M128_N32, 0u);
        umma_i8(d, umma_desc(a0 + TC_TILE_BYTES, 2048, 128), umma_desc(mats + 1024, 512, 128), TC_IDESC_U8_M128_N32, 1u);
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(b_mma + ts * 8) : "memory");
        ++m_q;
        if (++m_stage == NS) {
            m_stage = 0;
            m_par ^= 1u;
        }
    };
    // The two producer roles sit on different warps, and on different ones in the two warpgroups of a CTA: warp w of
    // every CTA issues from scheduler w % 4, so with two CTAs per SM each scheduler carries two roles (one thread
    // doing both for all four warpgroups of an SM loaded one scheduler 20-40 % above the other three).
    const bool tma_role = lane == 0 && wq == ((2u * wg) & 3u), mma_role = lane == 0 && wq == ((2u * wg + 1u) & 3u);
    if (tma_role && U) {
        for (uint32_t k = 0; k < NS && k < U; ++k) issue_tma();
    }
    if (mma_role && U) {
        for (uint32_t k = 0; k + 1 < NT && k < U; ++k) issue_mma();
    }
    uint4* accs = reinterpret_cast<uint4*>(smb + L::accs_off) + threadIdx.x;
    RoundAcc<F, D, NPTS, true, true> acc;
    acc.init(accs);
    uint32_t q = 0;
    for (uint32_t tk = 0; tk < my_tiles; ++tk) {
        const uint64_t j = (first + (uint64_t)tk * tstride) * 128 + r;
        for (int p = 0; p < n_products; ++p) {
            Fe m[NPTS - 1];
#pragma unroll 1
            for (int f = 0; f < D; ++f) {
                Fe lo, hi;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const uint32_t ts = q & (NT - 1);
                    if (mma_role && q + NT - 1 < U) issue_mma();
                    mbar_wait_u32(b_mma + ts * 8, (q >> LOG_NT) & 1u);
                    tc_fence_after();
                    uint32_t c[32];
                    tmem_ld32(tmem + ts * 32u + ((wq * 32u) << 16), c);
                    tmem_wait_ld();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_u32(b_empty + ts * 8);
                    if (tma_role && q + NS < U) issue_tma();  // stage q % NS: its MMAs have completed
                    ++q;
                    Fe& dst = h ? hi : lo;
                    dst = tc_fold_finish<F>(c);
                    st_fe(outp[p * D + f], h ? j + half : j, dst);
                }
                acc.factor(f, lo, hi, m);
            }
        }
    }
    // The folded entries were stored through the generic proxy; the next round of the persistent kernel reads them (in this
    // or another CTA, after the grid-wide hand-shake) with TMA, i.e. through the async proxy: order the two proxies.
    asm volatile("fence.proxy.async;" ::: "memory");
    acc.finish(out);
    wg_sync(wg);  // nobody of this warpgroup still waits on a barrier
    if (r == 0) {
#pragma unroll
        for (int b = 0; b < NS + 2 * NT; ++b) asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(b_full + b * 8) : "memory");
    }
}

template <int NPTS>
struct TailSmemTc {  // the persistent kernel may start with the evaluation pass, which stages through STAGE_BYTES
    static constexpr int bytes = TcRoundSmem<NPTS>::bytes > STAGE_BYTES ? TcRoundSmem<NPTS>::bytes : STAGE_BYTES;
};
// Shapes the tensor-core round pass is instantiated for (products of >= 2 factors, at most 4 points: the parked
// accumulators of more points leave no room for the operand stages of two CTAs per SM).
constexpr bool tc_shape(int kind, int D, int npts) { return kind == KIND_PROD && D >= 2 && npts <= 4; }
struct ScArgsTc {
    ScArgs s;
    TcFoldMats mats;
};

// ================================================================================================================
// Sums of products on the tensor cores (round 0 of a product of two factors).
//   s(0) = sum_j lo0_j lo1_j,   s(1) = sum_j hi0_j hi1_j,   s(2) = sum_j (2 hi0_j - lo0_j)(2 hi1_j - lo1_j) = 4 s(1) + s(0) - 2 X,
//   X = sum_j lo0_j hi1_j + hi0_j lo1_j
// are inner products over the table index j -- a contraction: with the elements written in bytes,
//   sum_j a_j b_j = sum_{i,k} 2^(8 (i + k)) sum_j a_{j,i} b_{j,k} = sum_{i,k} 2^(8 (i + k)) G[i][k],   G = A^T B  (u8 x u8 -> s32),
// so one tcgen05.mma stream over the rows j produces the 128 x 128 Gram matrix of the byte columns
// [lo0 | hi0 | lo1 | hi1] (A and B are the SAME shared-memory tile, both MN-major: 8 rows x 16 bytes of one plane are a
// core matrix exactly as they lie in HBM), and the CUDA cores do not multiply at all: every 2^15 rows (a column sum of
// 255^2 per row stays below 2^31) the four needed 32 x 32 blocks are read from TMEM and assembled into the 544-bit
// integers the lazy accumulators (fr.cuh Wide) already stand for.  Bit-identical to k_sc_eval.
constexpr int TCG_TILE_BYTES = 8 * 2048;  // 128 rows x 8 pieces (factor, half, plane) of 16 bytes
constexpr int TCG_NS = 4;                 // operand stages per warpgroup
constexpr uint32_t TC_IDESC_GRAM = (2u << 4) | (1u << 15) | (1u << 16) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
constexpr int TCG_SMEM = 2 * TCG_NS * TCG_TILE_BYTES + 3 * 32 * 12 * 4 * 2 + 256;  // stages, drain scratch, barriers
constexpr uint32_t TCG_DRAIN_ROWS = 1u << 15;

// V = sum_k c[k] 2^(8 k) for 32 column sums c[k] < 2^31: 10 limbs
__device__ __forceinline__ void tcg_assemble(const uint32_t* c, uint32_t* v) {
    uint64_t carry = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const uint64_t w = (uint64_t)c[4 * k] + ((uint64_t)c[4 * k + 1] << 8) + ((uint64_t)c[4 * k + 2] << 16) + ((uint64_t)c[4 * k + 3] << 24) + carry;
        v[k] = lo32(w);
        carry = w >> 32;
    }
    v[8] = lo32(carry);
    v[9] = hi32(carry);
}

template <class F, int NPTS>
__global__ void __launch_bounds__(BLOCK, 1) k_sc_eval_tc(const __grid_constant__ ScArgs a) {
    typedef Field<F> Fd;
    extern __shared__ __align__(128) uint8_t tc_sm[];
    __shared__ uint32_t s_tmem;
    __shared__ Fe s_out[2][2];
    const uint32_t wg = threadIdx.x >> 7, r = threadIdx.x & 127u, wq = (threadIdx.x >> 5) & 3u, lane = threadIdx.x & 31u;
    if (threadIdx.x < 32) tmem_alloc(&s_tmem, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem + wg * 128u;
    const uint64_t half = a.n_out >> 1;   // n_out = table size here (k_sc_eval convention)
    const uint64_t tiles = half >> 7;
    const uint64_t first = (uint64_t)blockIdx.x * 2 + wg, tstride = (uint64_t)gridDim.x * 2;
    const uint32_t P = (uint32_t)a.n_products;
    const uint32_t my_tiles = first < tiles ? (uint32_t)((tiles - first + tstride - 1) / tstride) : 0u;
    const uint32_t U = my_tiles * P;  // units: one product of one tile
    const uint32_t st0 = smem_u32(tc_sm) + wg * (TCG_NS * TCG_TILE_BYTES);
    uint32_t* scratch = reinterpret_cast<uint32_t*>(tc_sm + 2 * TCG_NS * TCG_TILE_BYTES) + wg * (3 * 32 * 12);
    const uint32_t b_full = smem_u32(tc_sm + 2 * TCG_NS * TCG_TILE_BYTES + 3 * 32 * 12 * 4 * 2) + wg * 128, b_free = b_full + TCG_NS * 8, b_acc = b_free + TCG_NS * 8;
    if (r == 0) {
        for (int b = 0; b < 2 * TCG_NS + 1; ++b) mbar_init_u32(b_full + b * 8, 1u);
        fence_barrier_init();
    }
    wg_sync(wg);
    // lazy sums: s(0) rows in warp 0, s(1) in warp 1, X shared; kept as Wide integers by lane 0 of warps 0 and 1
    Wide acc_d = Fd::wide_zero(), acc_x = Fd::wide_zero();  // "diagonal" (lo*lo resp. hi*hi) and cross sums of this warp
    const uint32_t units_per_drain = TCG_DRAIN_ROWS / 128u;
    uint32_t done = 0;
    while (done < U) {
        const uint32_t batch = U - done < units_per_drain ? U - done : units_per_drain;
        if (r == 0) {  // TMA producer
            for (uint32_t u = 0; u < batch; ++u) {
                const uint32_t q = done + u, stage = q % TCG_NS;
                if (q >= TCG_NS) mbar_wait_u32(b_free + stage * 8, ((q / TCG_NS) & 1u) ^ 1u);  // the MMAs of unit q - NS have read it
                const uint64_t tile = first + (uint64_t)(q / P) * tstride;
                const uint32_t p = q % P;
                const uint32_t dst = st0 + stage * TCG_TILE_BYTES, bar = b_full + stage * 8;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)TCG_TILE_BYTES) : "memory");
#pragma unroll
                for (int f = 0; f < 2; ++f) {
                    const TabRef& tb = a.in[p * 2 + f];
                    const uint4* g = tb.base + tile * 128;
#pragma unroll
                    for (int hp = 0; hp < 4; ++hp) {  // lo plane 0, lo plane 1, hi plane 0, hi plane 1
                        const uint4* src = g + ((hp & 2) ? half : 0) + ((hp & 1) ? tb.stride : 0);
                        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst + (f * 4 + hp) * 2048u), "l"(src), "r"(2048u), "r"(bar) : "memory");
                    }
                }
            }
        } else if (r == 32) {  // MMA issuer
            for (uint32_t u = 0; u < batch; ++u) {
                const uint32_t q = done + u, stage = q % TCG_NS;
                mbar_wait_u32(b_full + stage * 8, (q / TCG_NS) & 1u);
                tc_fence_after();
                const uint32_t t0 = st0 + stage * TCG_TILE_BYTES;
#pragma unroll
                for (uint32_t k4 = 0; k4 < 4; ++k4) {
                    const uint64_t d = umma_desc(t0 + k4 * 512u, 128, 2048);
                    umma_i8(tmem, d, d, TC_IDESC_GRAM, (u | k4) ? 1u : 0u);
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(b_free + stage * 8) : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(b_acc) : "memory");
        }
        __syncwarp();
        // drain: every thread waits for the batch, rows 0..63 (warps 0, 1 of the warpgroup) hold the needed blocks
        mbar_wait_u32(b_acc, (done / units_per_drain) & 1u);
        tc_fence_after();
        if (wq < 2) {
            uint32_t c[32], v[10];
            // columns 64..95 = bytes of lo1, 96..127 = bytes of hi1.  Warp 0 (rows = bytes of lo0): lo1 -> s(0), hi1 -> X;
            // warp 1 (rows = bytes of hi0): hi1 -> s(1), lo1 -> X.
#pragma unroll
            for (int part = 0; part < 2; ++part) {
                const uint32_t col = 64u + 32u * ((wq == 0) ? part : 1 - part);  // part 0 = diagonal, part 1 = cross
                tmem_ld32(tmem + col + ((wq * 32u) << 16), c);
                tmem_wait_ld();
                tcg_assemble(c, v);
                // this lane's term is V 2^(8 lane): shift by 8 (lane % 4) bits inside the limbs, by lane / 4 limbs through the scratch
                const uint32_t sh = 8u * (lane & 3u);
                uint32_t w[11];
                w[0] = v[0] << sh;
#pragma unroll
                for (int k = 1; k < 10; ++k) w[k] = sh ? (v[k] << sh) | (v[k - 1] >> (32u - sh)) : v[k];
                w[10] = sh ? v[9] >> (32u - sh) : 0u;
                uint32_t* row = scratch + (wq * 32u + lane) * 12u;
#pragma unroll
                for (int k = 0; k < 11; ++k) row[k] = w[k];
                __syncwarp();
                // column sums: limb L of the total = sum_i w_i[L - i / 4]; lane L < 18 takes limb L, then a serial carry
                uint64_t colsum = 0;
                if (lane < 18) {
                    for (uint32_t i = 0; i < 32; ++i) {
                        const int k = (int)lane - (int)(i >> 2);
                        if (k >= 0 && k < 11) colsum += scratch[(wq * 32u + i) * 12u + k];
                    }
                }
                __syncwarp();
                // carry propagation by shuffles (18 steps), result limb in `lim`
                uint64_t carry = 0;
                uint32_t lim = 0;
                for (uint32_t L = 0; L < 18; ++L) {
                    const uint64_t cs = __shfl_sync(0xffffffffu, colsum, L) + carry;
                    if (lane == L) lim = lo32(cs);
                    carry = cs >> 32;
                }
                // lane 0 adds the 18-limb total (limb 17 is zero: < 2^544) into its Wide accumulator
                Wide& dst = part == 0 ? acc_d : acc_x;
                uint32_t cc = 0;
                for (uint32_t L = 0; L < 17; ++L) {
                    const uint32_t x = __shfl_sync(0xffffffffu, lim, L);
                    if (lane == 0) {
                        const uint64_t t = (uint64_t)dst.l[L] + x + cc;
                        dst.l[L] = lo32(t);
                        cc = hi32(t);
                    }
                }
                __syncwarp();
            }
        }
        tc_fence_before();
        wg_sync(wg);  // the accumulator has been read: the next batch may overwrite it
        done += batch;
    }
    // s(0), s(1), X of this warpgroup -> CTA totals -> the round's three evaluations
    Fe out[NPTS];
#pragma unroll
    for (int p3 = 0; p3 < NPTS; ++p3) out[p3] = Fd::zero();
    if (wq < 2 && lane == 0) {
        s_out[wg][wq == 0 ? 0 : 1] = Fd::reduce_wide(acc_d);
        // the two cross sums: warp 0's goes through shared memory to be added to warp 1's
    }
    __shared__ Fe s_cross[2][2];
    if (wq < 2 && lane == 0) s_cross[wg][wq] = Fd::reduce_wide(acc_x);
    __syncthreads();
    if (r == 0) {
        const Fe s0 = s_out[wg][0], s1 = s_out[wg][1], x = Fd::add(s_cross[wg][0], s_cross[wg][1]);
        // s(t) = c0 + c1 t + c2 t^2 with c0 = s(0), c2 = sum d0 d1 = s(0) + s(1) - X: forward differences
        // s(t + 1) - s(t) = (s(1) - s(0)) + 2 t c2
        const Fe two_c2 = Fd::dbl(Fd::sub(Fd::add(s0, s1), x));
        Fe delta = Fd::sub(s1, s0);
        out[0] = s0;
#pragma unroll
        for (int t = 1; t < NPTS; ++t) {
            out[t] = Fd::add(out[t - 1], delta);
            delta = Fd::add(delta, two_c2);
        }
    }
    finish_round<F, NPTS>(out, a.fin);
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(s_tmem, 256);
}

// ================================================================================================================
// Products of >= 3 factors: the sums of products of EVERY round go to the tensor cores the same way.  Per pair
// position the CUDA cores still multiply the first D - 1 factors (full Montgomery products: the only multiplier work
// left), m_k = prod_{f < D-1} (lo_f + t_k d_f) at the round's points t_k; the last factor enters linearly,
//   s(t_k) = sum_j m_k,j ((1 - t_k) lo_j + t_k hi_j) = (1 - t_k) G(m_k, lo) + t_k G(m_k, hi),
// and the inner products G are blocks of the Gram matrix [m_0 | .. | m_{K-1}]^T [lo | hi] of byte columns: the threads
// of a warpgroup write their K + 2 values into a shared-memory tile (row = pair position, MN-major pieces of 128 rows x
// 16 bytes), one thread issues four tcgen05.mma (128 rows) that accumulate into a 128 x 64 s32 tile in TMEM, and every
// 2^15 rows warp k turns rows 32 k .. 32 k + 31 of the accumulator into two 544-bit integers (fr.cuh Wide).
constexpr uint32_t TC_IDESC_GRAM64 = (2u << 4) | (1u << 15) | (1u << 16) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
// rows 32 wq .. 32 wq + 31 (this warp) x 32 columns starting at `col` of a Gram accumulator -> sum_i,k C[i][k] 2^(8 (i + k)),
// added by lane 0 into the 17-word integer at dst (shared memory).  scratch: 32 x 12 words of this warp.
__device__ __forceinline__ void tcg_block_add(uint32_t taddr, uint32_t* scratch, uint32_t lane, uint32_t* dst) {
    uint32_t c[32], v[10];
    tmem_ld32(taddr, c);
    tmem_wait_ld();
    tcg_assemble(c, v);
    // this lane's term is V 2^(8 lane): 8 (lane % 4) bits inside the limbs, lane / 4 limbs through the scratch rows
    const uint32_t sh = 8u * (lane & 3u);
    uint32_t* row = scratch + lane * 12u;
    row[0] = v[0] << sh;
#pragma unroll
    for (int k = 1; k < 10; ++k) row[k] = sh ? (v[k] << sh) | (v[k - 1] >> (32u - sh)) : v[k];
    row[10] = sh ? v[9] >> (32u - sh) : 0u;
    __syncwarp();
    uint64_t colsum = 0;  // limb L of the total = sum_i row_i[L - i / 4]
    if (lane < 18) {
        for (uint32_t i = 0; i < 32; ++i) {
            const int k = (int)lane - (int)(i >> 2);
            if (k >= 0 && k < 11) colsum += scratch[i * 12u + k];
        }
    }
    __syncwarp();
    uint64_t carry = 0;
    uint32_t lim = 0;
    for (uint32_t L = 0; L < 18; ++L) {
        const uint64_t cs = __shfl_sync(0xffffffffu, colsum, L) + carry;
        if (lane == L) lim = lo32(cs);
        carry = cs >> 32;
    }
    uint32_t cc = 0;
    for (uint32_t L = 0; L < 17; ++L) {
        const uint32_t x = __shfl_sync(0xffffffffu, lim, L);
        if (lane == 0) {
            const uint64_t t = (uint64_t)dst[L] + x + cc;
            dst[L] = lo32(t);
            cc = hi32(t);
        }
    }
    __syncwarp();
}

// GramAcc2: the accumulator, with UNREDUCED products among the A rows.  A raw 512-bit product T = x y of two residues
// costs 64 wide multiplies instead of the 136 of a Montgomery product; written as 64 byte rows (low half L, high half H,
// T = L + 2^256 H) it enters the contraction like any other value, and the two missing divisions by R = 2^256 are done once,
// at the very end: sum_j T_j c_j R^-2 = redc(redc(G(L, c))) + redc(G(H, c)).  The tile is [A: 4 slots of 32 bytes][B: lo, hi]
// = 24 KiB per warpgroup; GROUPS accumulator tiles (64 columns each) take turns on the same A buffer, the B rows of a
// push are shared by its groups.  Warp w of the warpgroup owns accumulator rows 32 w .. 32 w + 31 = A slot w.
template <class F, int GROUPS>
struct GramAcc2 {
    typedef Field<F> Fd;
    static constexpr int tile_bytes = 12 * 2048;
    static constexpr int sums_words = GROUPS * 4 * 2 * 17;
    static constexpr int bytes_per_wg = tile_bytes + sums_words * 4 + GROUPS * 4 * 2 * 32 + 64;
    uint32_t tile, b_full, b_free, tmem, r, wq, lane, rows, steps;
    uint32_t* sums;
    Fe* blk;  // [GROUPS][4][2]: the reduced block sums after finish()
    bool mma_role;
    __device__ __forceinline__ void init(uint8_t* base_wg, uint32_t tmem_cols, uint32_t wg_, uint32_t mma_warp) {
        r = threadIdx.x & 127u;
        wq = (threadIdx.x >> 5) & 3u;
        lane = threadIdx.x & 31u;
        tile = smem_u32(base_wg);
        sums = reinterpret_cast<uint32_t*>(base_wg + tile_bytes);
        blk = reinterpret_cast<Fe*>(base_wg + tile_bytes + sums_words * 4);
        b_full = smem_u32(base_wg + tile_bytes + sums_words * 4 + GROUPS * 4 * 2 * 32);
        b_free = b_full + 8;
        tmem = tmem_cols;
        mma_role = lane == 0 && wq == mma_warp;
        rows = steps = 0;
        for (uint32_t i = r; i < (uint32_t)sums_words; i += 128) sums[i] = 0;
        if (r == 0) {
            mbar_init_u32(b_full, 128u);
            mbar_init_u32(b_free, 1u);
            fence_barrier_init();
        }
        wg_sync(wg_);
    }
    __device__ __forceinline__ void put(uint32_t piece, const uint32_t* v) {
        const uint32_t a0 = tile + piece * 2048u + r * 16u;
        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a0), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]) : "memory");
        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a0 + 2048u), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
    }
    // the A buffer (and, for the first group of a push, the B rows) may be written again
    __device__ __forceinline__ void wait_free() {
        if (steps) mbar_wait_u32(b_free, (steps - 1u) & 1u);
    }
    __device__ __forceinline__ void put_a(uint32_t slot, const Fe& v) { put(2 * slot, v.l); }
    // the raw product x * y (x, y < 2^256) into slots `slot` (low half) and `slot + 1` (high half)
    __device__ __forceinline__ void put_raw(uint32_t slot, const Fe& x, const Fe& y) {
        Wide t = Fd::wide_zero();
        Fd::mac_wide(t, x, y);
        put(2 * slot, t.l);
        put(2 * slot + 2, t.l + 8);
    }
    __device__ __forceinline__ void put_b(const Fe& lo, const Fe& hi) {
        put(8, lo.l);
        put(10, hi.l);
    }
    // every thread of the warpgroup has written its row: accumulate the tile into group g
    __device__ __forceinline__ void commit(uint32_t g, uint32_t wg_) {
        fence_proxy_async();
        mbar_arrive_u32(b_full);
        if (mma_role) {
            mbar_wait_u32(b_full, steps & 1u);
            tc_fence_after();
#pragma unroll
            for (uint32_t k4 = 0; k4 < 4; ++k4)
                umma_i8(tmem + g * 64u, umma_desc(tile + k4 * 512u, 128, 2048), umma_desc(tile + 8u * 2048u + k4 * 512u, 128, 2048), TC_IDESC_GRAM64,
                        (rows | k4) ? 1u : 0u);
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(b_free) : "memory");
        }
        ++steps;
        if (g == GROUPS - 1) {
            rows += 128u;
            if (rows == TCG_DRAIN_ROWS) drain(wg_);
        }
    }
    __device__ __forceinline__ void drain(uint32_t wg_) {
        if (!rows) return;
        mbar_wait_u32(b_free, (steps - 1u) & 1u);  // the last commit covers every MMA before it; the tile is free: scratch
        tc_fence_after();
        uint32_t* scratch = reinterpret_cast<uint32_t*>(sums) - (tile_bytes / 4) + wq * (32 * 12);  // inside the A buffer
#pragma unroll
        for (uint32_t g = 0; g < (uint32_t)GROUPS; ++g) {
            tcg_block_add(tmem + g * 64u + ((wq * 32u) << 16), scratch, lane, sums + ((g * 4 + wq) * 2) * 17);
            tcg_block_add(tmem + g * 64u + 32u + ((wq * 32u) << 16), scratch, lane, sums + ((g * 4 + wq) * 2 + 1) * 17);
        }
        tc_fence_before();
        wg_sync(wg_);
        rows = 0;
    }
    // blk[g][w][part] = G(A slot w of group g, part of B) R^-1 mod p, canonical; valid for every thread of the warpgroup
    __device__ __forceinline__ void finish(uint32_t wg_) {
        drain(wg_);
        if (lane < 2u * GROUPS) {
            const uint32_t g = lane >> 1, part = lane & 1u;
            Wide w;
#pragma unroll
            for (int i = 0; i < 17; ++i) w.l[i] = sums[((g * 4 + wq) * 2 + part) * 17 + i];
            blk[(g * 4 + wq) * 2 + part] = Fd::reduce_wide(w);
        }
        wg_sync(wg_);
        if (r == 0) {
            asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(b_full) : "memory");
            asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(b_free) : "memory");
        }
    }
    // one more division by R (for the low half of a raw product)
    __device__ __forceinline__ static Fe redc(const Fe& x) {
        Fe one = Fd::zero();
        one.l[0] = 1;
        return Fd::mul(x, one);
    }
    // sum_j T_j c_j for the raw product in slots (s, s + 1) of group g against B part `part`, as a Montgomery residue
    __device__ __forceinline__ Fe raw_sum(uint32_t g, uint32_t s, uint32_t part) const {
        return Fd::add(redc(blk[(g * 4 + s) * 2 + part]), blk[(g * 4 + s + 1) * 2 + part]);
    }
};

// ---- fold rounds of products of >= 3 factors: round_pass_tc with the sums through GramAcc ----
template <int D, int NPTS>
struct TcGramRoundSmem {
    static constexpr int NS = 3, NT = 2;
    static constexpr int stage_bytes = 2 * NS * TC_UNIT_BYTES;
    static constexpr int mats_off = stage_bytes;
    static constexpr int bars_off = mats_off + 2048;
    static constexpr int bars_per_wg = (NS + 2 * NT) * 8;
    static constexpr int tmem_off = bars_off + 2 * bars_per_wg;
    static constexpr int gram_off = bars_off + 256;
    static constexpr int gram_per_wg = (GramAcc2<Bn254Fr, 1>::bytes_per_wg + 127) / 128 * 128;
    static constexpr int bytes = gram_off + 2 * gram_per_wg;
    static constexpr int tmem_cols = 256;  // per CTA: per warpgroup 2 x 32 fold columns + 64 Gram columns
};
template <class F, int D, int NPTS>
__device__ __forceinline__ void round_pass_tc_gram(const TabRef* __restrict__ in, const TabRef* __restrict__ outp, int n_products, uint64_t n_out,
                                                   uint8_t* smb, uint32_t tmem_cta, Fe* out) {
    static_assert(BLOCK == 256 && D == 3 && NPTS == 4, "two warpgroups per CTA; one raw and two reduced partial products fill the 128 accumulator rows");
    typedef Field<F> Fd;
    typedef TcGramRoundSmem<D, NPTS> L;
    typedef Slots<NPTS, true> S;
    constexpr int NS = L::NS, NT = L::NT, LOG_NT = 1, K = NPTS - 1;
    const uint32_t wg = threadIdx.x >> 7, r = threadIdx.x & 127u, wq = (threadIdx.x >> 5) & 3u, lane = threadIdx.x & 31u;
    const uint64_t half = n_out >> 1;
    const uint64_t tiles = half >> 7;
    const uint64_t first = (uint64_t)blockIdx.x * 2 + wg, tstride = (uint64_t)gridDim.x * 2;
    const uint32_t upt = 2u * (uint32_t)n_products * D;
    const uint32_t my_tiles = first < tiles ? (uint32_t)((tiles - first + tstride - 1) / tstride) : 0u;
    const uint32_t U = my_tiles * upt;
    const uint32_t st0 = smem_u32(smb) + wg * (NS * TC_UNIT_BYTES);
    const uint32_t mats = smem_u32(smb + L::mats_off);
    const uint32_t b_full = smem_u32(smb + L::bars_off) + wg * L::bars_per_wg, b_mma = b_full + NS * 8, b_empty = b_mma + NT * 8;
    const uint32_t tmem = tmem_cta + wg * 128u;
    if (r == 0) {
#pragma unroll
        for (int b = 0; b < NS + 2 * NT; ++b) mbar_init_u32(b_full + b * 8, b < NS + NT ? 1u : 4u);
        fence_barrier_init();
    }
    const bool tma_role = lane == 0 && wq == ((2u * wg) & 3u), mma_role = lane == 0 && wq == ((2u * wg + 1u) & 3u);
    GramAcc2<F, 1> gram;
    gram.init(smb + L::gram_off + wg * L::gram_per_wg, tmem + 64u, wg, (2u * wg + 2u) & 3u);  // (its wg_sync publishes the barriers too)
    uint64_t t_tile = first;
    uint32_t t_u = 0, t_stage = 0, m_q = 0, m_stage = 0, m_par = 0;
    auto issue_tma = [&]() {
        const TabRef& tb = in[t_u >> 1];
        const uint4* g0 = tb.base + t_tile * 128 + ((t_u & 1u) ? half : 0);
        const uint4* g1 = g0 + tb.stride;
        const uint32_t dst = st0 + t_stage * TC_UNIT_BYTES, bar = b_full + t_stage * 8;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)TC_UNIT_BYTES) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(g0), "r"(2048u), "r"(bar) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst + 2048u), "l"(g1), "r"(2048u), "r"(bar) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst + 4096u), "l"(g0 + n_out), "r"(2048u), "r"(bar) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst + 6144u), "l"(g1 + n_out), "r"(2048u), "r"(bar) : "memory");
        if (++t_u == upt) {
            t_u = 0;
            t_tile += tstride;
        }
        t_stage = t_stage + 1 == NS ? 0u : t_stage + 1;
    };
    auto issue_mma = [&]() {
        const uint32_t ts = m_q & (NT - 1);
        mbar_wait_u32(b_full + m_stage * 8, m_par);
        if (m_q >= NT) mbar_wait_u32(b_empty + ts * 8, ((m_q >> LOG_NT) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d = tmem + ts * 32u, a0 = st0 + m_stage * TC_UNIT_BYTES;
        umma_i8(d, umma_desc(a0, 2048, 128), umma_desc(mats, 512, 128), TC_IDESC_U8_M128_N32, 0u);
        umma_i8(d, umma_desc(a0 + TC_TILE_BYTES, 2048, 128), umma_desc(mats + 1024, 512, 128), TC_IDESC_U8_M128_N32, 1u);
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(b_mma + ts * 8) : "memory");
        ++m_q;
        if (++m_stage == NS) {
            m_stage = 0;
            m_par ^= 1u;
        }
    };
    if (tma_role && U) {
        for (uint32_t k = 0; k < NS && k < U; ++k) issue_tma();
    }
    if (mma_role && U) {
        for (uint32_t k = 0; k + 1 < NT && k < U; ++k) issue_mma();
    }
    uint32_t q = 0;
    for (uint32_t tk = 0; tk < my_tiles; ++tk) {
        const uint64_t j = (first + (uint64_t)tk * tstride) * 128 + r;
        for (int p = 0; p < n_products; ++p) {
            // A rows of this position: the RAW product lo0 * lo1 (point 0, slots 0-1) and the reduced products at the points
            // 2 and 3 (slots 2, 3); B rows: the folded pair of the last factor
            Fe m[K], y0, lo, hi;
#pragma unroll 1
            for (int f = 0; f < D; ++f) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const uint32_t ts = q & (NT - 1);
                    if (mma_role && q + NT - 1 < U) issue_mma();
                    mbar_wait_u32(b_mma + ts * 8, (q >> LOG_NT) & 1u);
                    tc_fence_after();
                    uint32_t c[32];
                    tmem_ld32(tmem + ts * 32u + ((wq * 32u) << 16), c);
                    tmem_wait_ld();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_u32(b_empty + ts * 8);
                    if (tma_role && q + NS < U) issue_tma();
                    ++q;
                    Fe& dst = h ? hi : lo;
                    dst = tc_fold_finish<F>(c);
                    st_fe(outp[p * D + f], h ? j + half : j, dst);
                }
                if (f < D - 1) {  // the first two factors at the points 0, 2, 3
                    Fe cur = lo;
                    const Fe d = Fd::sub(hi, lo);
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        if (k == 1) cur = Fd::add(hi, d);
                        else if (k > 1) cur = Fd::add(cur, d);
                        if (f == 0) m[k] = cur;
                        else if (k == 0) y0 = cur;
                        else m[k] = Fd::mul(m[k], cur);
                    }
                }
            }
            gram.wait_free();
            gram.put_raw(0, m[0], y0);
            gram.put_a(2, m[1]);
            gram.put_a(3, m[2]);
            gram.put_b(lo, hi);
            gram.commit(0, wg);
        }
    }
    asm volatile("fence.proxy.async;" ::: "memory");  // generic-proxy stores of the folded tables before later TMA reads (see round_pass_tc)
    gram.finish(wg);
#pragma unroll
    for (int k = 0; k < K; ++k) out[k] = Fd::zero();
    if (r == 0) {  // s(t) = (1 - t) G(., lo) + t G(., hi)
        out[0] = gram.raw_sum(0, 0, 0);
#pragma unroll
        for (int k = 1; k < K; ++k) {
            const Fe gl = gram.blk[(k + 1) * 2], d = Fd::sub(gram.blk[(k + 1) * 2 + 1], gl);
            Fe v = gl;
            for (int i = 0; i < k + 1; ++i) v = Fd::add(v, d);
            out[k] = v;
        }
    }
    if (r == 0) {
#pragma unroll
        for (int b = 0; b < NS + 2 * NT; ++b) asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(b_full + b * 8) : "memory");
    }
}

// ---- round 0 of a product of 3 factors: all four points.  q(t) = (lo0 + t d0)(lo1 + t d1) is a quadratic in t, so three
// full products (t = 0, 1, infinity) give it at every point by second differences; the third factor enters through the
// Gram accumulator.  Reads as in eval_pass (a private cp.async slot per thread), tiles of 128 pair positions per warpgroup.
template <int NPTS>
struct TcGramEvalSmem {
    static constexpr int NBUF = 2;
    static constexpr int stage_bytes = NBUF * 4 * BLOCK * 16;
    static constexpr int gram_off = stage_bytes;
    static constexpr int gram_per_wg = (GramAcc2<Bn254Fr, 2>::bytes_per_wg + 127) / 128 * 128;
    static constexpr int bytes = gram_off + 2 * gram_per_wg;
    static constexpr int tmem_cols = 256;  // per CTA: two 64-column Gram accumulators per warpgroup
};
template <class F, int D, int NPTS>
__global__ void __launch_bounds__(BLOCK, 2) k_sc_eval_gram(const __grid_constant__ ScArgs a) {
    static_assert(D == 3 && NPTS == 4 && BLOCK == 256, "product of three factors");
    typedef Field<F> Fd;
    typedef TcGramEvalSmem<NPTS> L;
    extern __shared__ __align__(128) uint8_t tc_sm[];
    __shared__ uint32_t s_tmem;
    const uint32_t wg = threadIdx.x >> 7, r = threadIdx.x & 127u;
    if (threadIdx.x < 32) tmem_alloc(&s_tmem, L::tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint64_t half = a.n_out >> 1;  // n_out = table size (k_sc_eval convention)
    const uint64_t tiles = half >> 7;
    const uint64_t first = (uint64_t)blockIdx.x * 2 + wg, tstride = (uint64_t)gridDim.x * 2;
    const uint32_t my_tiles = first < tiles ? (uint32_t)((tiles - first + tstride - 1) / tstride) : 0u;
    const int T = a.n_products * D;
    GramAcc2<F, 2> gram;
    gram.init(tc_sm + L::gram_off + wg * L::gram_per_wg, s_tmem + wg * 128u, wg, wg);
    uint4* my = reinterpret_cast<uint4*>(tc_sm) + threadIdx.x;
    // flattened prefetch sequence: (tile, table)
    uint32_t p_tile = 0;
    int pt = 0, pbuf = 0, cbuf = 0;
    auto issue = [&]() {
        if (p_tile < my_tiles) {
            const TabRef& t = a.in[pt];
            const uint64_t pj = (first + (uint64_t)p_tile * tstride) * 128 + r;
            uint4* dst = my + (size_t)pbuf * 4 * BLOCK;
            cp_async16(dst, t.base + pj);
            cp_async16(dst + BLOCK, t.base + t.stride + pj);
            cp_async16(dst + 2 * BLOCK, t.base + pj + half);
            cp_async16(dst + 3 * BLOCK, t.base + t.stride + pj + half);
        }
        cp_async_commit();
        pbuf = (pbuf + 1) & (L::NBUF - 1);
        if (++pt == T) {
            pt = 0;
            ++p_tile;
        }
    };
#pragma unroll
    for (int k = 0; k < L::NBUF; ++k) issue();
    auto take = [&](Fe& lo, Fe& hi) {
        cp_async_wait<L::NBUF - 1>();
        const uint4* src = my + (size_t)cbuf * 4 * BLOCK;
        lo = fe_from_smem(src, src + BLOCK);
        hi = fe_from_smem(src + 2 * BLOCK, src + 3 * BLOCK);
        cbuf = (cbuf + 1) & (L::NBUF - 1);
        issue();
    };
    for (uint32_t tk = 0; tk < my_tiles; ++tk) {
        for (int p = 0; p < a.n_products; ++p) {
            // q(t) = (lo0 + t d0)(lo1 + t d1) at t = 0..3 as RAW products (64 wide multiplies each instead of three Montgomery
            // products and second differences): points 0, 1 go to accumulator group 0, points 2, 3 to group 1
            Fe lo0, hi0, lo1, hi1, lo2, hi2;
            take(lo0, hi0);
            take(lo1, hi1);
            gram.wait_free();
            gram.put_raw(0, lo0, lo1);
            gram.put_raw(2, hi0, hi1);
            take(lo2, hi2);
            gram.put_b(lo2, hi2);
            gram.commit(0, wg);
            const Fe d0 = Fd::sub(hi0, lo0), d1 = Fd::sub(hi1, lo1);
            const Fe a2 = Fd::add(hi0, d0), b2 = Fd::add(hi1, d1);
            const Fe a3 = Fd::add(a2, d0), b3 = Fd::add(b2, d1);
            gram.wait_free();
            gram.put_raw(0, a2, b2);
            gram.put_raw(2, a3, b3);
            gram.commit(1, wg);
        }
    }
    Fe out[NPTS];
    gram.finish(wg);
#pragma unroll
    for (int k = 0; k < NPTS; ++k) out[k] = Fd::zero();
    if (r == 0) {  // s(t) = (1 - t) G(q_t, lo2) + t G(q_t, hi2)
#pragma unroll
        for (int k = 0; k < NPTS; ++k) {
            const Fe gl = gram.raw_sum(k >> 1, 2 * (k & 1), 0), d = Fd::sub(gram.raw_sum(k >> 1, 2 * (k & 1), 1), gl);
            Fe v = gl;
            for (int i = 0; i < k; ++i) v = Fd::add(v, d);
            out[k] = v;
        }
    }
    cp_async_wait<0>();
    finish_round<F, NPTS>(out, a.fin);
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(s_tmem, L::tmem_cols);
}

// ---- evaluate / multi_partial_evaluate: three variables per pass on the tensor cores.  The bound entry is a LINEAR
// combination of 8 input entries, out[v] = sum_i w_i in[v + i n_out] with w_i = prod_l (bit_l(i) ? r_l : 1 - r_l), i.e.
// eight u8 x u8 -> s32 products accumulated in one 128 x 32 tile (column sums < 2^24, S < 2^16 p: still one Montgomery
// row).  The CUDA-core version does 7 fixed-multiplicand folds per output and is multiplier-bound at 0.40 of HBM.
struct MultiFoldTcArgs {
    TabRef in, out;
    uint64_t n_out;
    uint8_t mats[8][1024];  // B operand of weight w_i: byte n of w_i 2^(8 k + 32) mod p at (k / 16) * 512 + n * 16 + k % 16
};
constexpr int TCM_NS = 3, TCM_UNIT = 16 * 2048, TCM_THREADS = 128;
constexpr int TCM_SMEM = TCM_NS * TCM_UNIT + 8192 + 128;
template <class F>
__global__ void __launch_bounds__(TCM_THREADS, 2) k_multifold_tc(const __grid_constant__ MultiFoldTcArgs a) {
    extern __shared__ __align__(128) uint8_t tc_sm[];
    const uint32_t r = threadIdx.x, wq = r >> 5, lane = r & 31u;
    const uint32_t st0 = smem_u32(tc_sm), mats = st0 + TCM_NS * TCM_UNIT;
    const uint32_t b_full = mats + 8192, b_mma = b_full + TCM_NS * 8, b_empty = b_mma + 16;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tc_sm + TCM_NS * TCM_UNIT + 8192 + (TCM_NS + 4) * 8);
    for (uint32_t i = r; i < 8192 / 16; i += TCM_THREADS) reinterpret_cast<uint4*>(tc_sm + TCM_NS * TCM_UNIT)[i] = reinterpret_cast<const uint4*>(a.mats)[i];
    if (r == 0) {
        for (int b = 0; b < TCM_NS + 4; ++b) mbar_init_u32(b_full + b * 8, b < TCM_NS + 2 ? 1u : 4u);
        fence_barrier_init();
    }
    fence_proxy_async();
    if (r < 32) tmem_alloc(tmem_slot, 64);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint64_t tiles = a.n_out >> 7;
    const uint32_t U = blockIdx.x < tiles ? (uint32_t)((tiles - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0u;
    const bool tma_role = r == 0, mma_role = r == 32;
    uint32_t t_q = 0, t_stage = 0, m_q = 0, m_stage = 0, m_par = 0;
    auto issue_tma = [&]() {
        const uint64_t j0 = ((uint64_t)blockIdx.x + (uint64_t)t_q * gridDim.x) * 128;
        const uint32_t dst = st0 + t_stage * TCM_UNIT, bar = b_full + t_stage * 8;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)TCM_UNIT) : "memory");
#pragma unroll
        for (uint32_t i = 0; i < 8; ++i) {
            const uint4* g0 = a.in.base + (uint64_t)i * a.n_out + j0;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst + i * 4096u), "l"(g0), "r"(2048u), "r"(bar) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst + i * 4096u + 2048u), "l"(g0 + a.in.stride), "r"(2048u), "r"(bar) : "memory");
        }
        ++t_q;
        t_stage = t_stage + 1 == TCM_NS ? 0u : t_stage + 1;
    };
    auto issue_mma = [&]() {
        const uint32_t ts = m_q & 1u;
        mbar_wait_u32(b_full + m_stage * 8, m_par);
        if (m_q >= 2) mbar_wait_u32(b_empty + ts * 8, ((m_q >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d = tmem + ts * 32u, a0 = st0 + m_stage * TCM_UNIT;
#pragma unroll
        for (uint32_t i = 0; i < 8; ++i) umma_i8(d, umma_desc(a0 + i * 4096u, 2048, 128), umma_desc(mats + i * 1024u, 512, 128), TC_IDESC_U8_M128_N32, i ? 1u : 0u);
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(b_mma + ts * 8) : "memory");
        ++m_q;
        if (++m_stage == TCM_NS) {
            m_stage = 0;
            m_par ^= 1u;
        }
    };
    if (tma_role) {
        for (uint32_t k = 0; k < TCM_NS && k < U; ++k) issue_tma();
    }
    if (mma_role && U) issue_mma();
    for (uint32_t q = 0; q < U; ++q) {
        const uint32_t ts = q & 1u;
        if (mma_role && q + 1 < U) issue_mma();
        mbar_wait_u32(b_mma + ts * 8, (q >> 1) & 1u);
        tc_fence_after();
        uint32_t c[32], s[10];
        tmem_ld32(tmem + ts * 32u + ((wq * 32u) << 16), c);
        tmem_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_u32(b_empty + ts * 8);
        if (tma_role && q + TCM_NS < U) issue_tma();
        tcg_assemble(c, s);
        st_fe(a.out, ((uint64_t)blockIdx.x + (uint64_t)q * gridDim.x) * 128 + r, tc_mont_row<F>(s));
    }
    tc_fence_before();
    __syncthreads();
    if (r < 32) tmem_dealloc(tmem, 64);
}

// ---- one round as one launch: folds on the tensor cores; for >= 3 factors the sums of products too ----
template <int D, int NPTS>
struct TcFoldEvalCfg {
    static constexpr bool gram = D >= 3;
    static constexpr int smem = gram ? TcGramRoundSmem<D, NPTS>::bytes : TcRoundSmem<NPTS>::bytes;
    static constexpr int tmem_cols = gram ? TcGramRoundSmem<D, NPTS>::tmem_cols : TcCfg<NPTS>::tmem_cols;
    static constexpr int tmem_off = gram ? TcGramRoundSmem<D, NPTS>::tmem_off : TcRoundSmem<NPTS>::tmem_off;
    static constexpr int mats_off = gram ? TcGramRoundSmem<D, NPTS>::mats_off : TcRoundSmem<NPTS>::mats_off;
};
template <class F, int KIND, int D, int NPTS>
__global__ void __launch_bounds__(BLOCK, 2) k_sc_fold_eval_tc(const __grid_constant__ ScArgsTc a) {
    typedef TcFoldEvalCfg<D, NPTS> L;
    extern __shared__ __align__(128) uint8_t tc_sm[];
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tc_sm + L::tmem_off);
    for (uint32_t i = threadIdx.x; i < 2048 / 16; i += BLOCK)
        reinterpret_cast<uint4*>(tc_sm + L::mats_off)[i] = reinterpret_cast<const uint4*>(&a.mats)[i];
    fence_proxy_async();
    if (threadIdx.x < 32) tmem_alloc(tmem_slot, L::tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    Fe out[NPTS - 1];
    if constexpr (L::gram) round_pass_tc_gram<F, D, NPTS>(a.s.in, a.s.out, a.s.n_products, a.s.n_out, tc_sm, tmem, out);
    else round_pass_tc<F, D, NPTS>(a.s.in, a.s.out, a.s.n_products, a.s.n_out, tc_sm, tmem, out);
    finish_round<F, NPTS - 1>(out, a.s.fin);
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tmem, L::tmem_cols);
}

// ---- co-residency probe.  The tensor-core variant of the persistent kernel is a plain launch of 2 CTAs per SM that wait for
// each other (cudaLaunchCooperativeKernel refuses it: the occupancy calculator answers 1 for kernels that allocate tensor
// memory).  This kernel takes the same resources (shared memory, 128 TMEM columns, 256 threads) and reports whether all
// CTAs of the grid were resident at the same time; the host runs it once per context and keeps the persistent kernel on the
// CUDA cores if they were not.
static __global__ void __launch_bounds__(BLOCK, 2) k_tc_probe(unsigned int* counter, unsigned int* failed, long long timeout_clocks) {
    extern __shared__ __align__(128) uint8_t tc_sm[];
    __shared__ uint32_t s_tmem;
    if (threadIdx.x < 32) tmem_alloc(&s_tmem, 128);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (threadIdx.x == 0) {
        tc_sm[0] = 1;  // (the dynamic allocation is what matters)
        atomicAdd(counter, 1u);
        const long long t0 = clock64();
        while (*reinterpret_cast<volatile unsigned int*>(counter) < gridDim.x) {
            if (clock64() - t0 > timeout_clocks) {
                *failed = 1u;
                break;
            }
            __nanosleep(200);
        }
    }
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(s_tmem, 128);
}
#endif  // __CUDACC__

}  // namespace zkb
