// fr.cuh -- 256-bit prime-field arithmetic for sm_100a: 8 x 32-bit limbs,
// Montgomery form with R = 2^256 (bit-compatible with ark-ff's
// Fp<MontBackend<_,4>>: the same residues, the same little-endian limb order,
// so a host `Vec<F>` uploads without conversion).
//
// Replaces (as the arithmetic under every table kernel) the ark-ff operators
// the reference uses through `F: PrimeField`
// (multilinear_polynomial/src/multilinear_polynomial_evaluation.rs:13-14,59,94).
//
// The multiplier is a row-wise (CIOS) Montgomery product with the partial
// products kept in two interleaved accumulators ("even"/"odd" limb columns) so
// that every 32x32->64 product lands on a 64-bit-aligned limb pair and ptxas
// can fuse each `mad.lo.cc / madc.hi.cc` pair into ONE `IMAD.WIDE.U32.X` with a
// predicate carry: 136 wide multiply-accumulates per product, no carry-save
// shuffling.  The same source compiles for the host with the carry flag
// emulated, which is how the limb algorithm is unit-tested without a GPU
// (tests/test_fr_host.py).
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define ZK_HD __host__ __device__ __forceinline__
#define ZK_D __device__ __forceinline__
#else
#define ZK_HD inline
#define ZK_D inline
#endif

namespace zkb {

// ---------------------------------------------------------------- carry ops
// Device: PTX carry-flag instructions (asm volatile keeps their order).
// Host: the flag is a thread-local variable.
#if defined(__CUDA_ARCH__)
#define ZK_ASM asm volatile
ZK_D uint32_t add_cc(uint32_t a, uint32_t b) { uint32_t r; ZK_ASM("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ZK_D uint32_t addc_cc(uint32_t a, uint32_t b) { uint32_t r; ZK_ASM("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ZK_D uint32_t addc(uint32_t a, uint32_t b) { uint32_t r; ZK_ASM("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ZK_D uint32_t sub_cc(uint32_t a, uint32_t b) { uint32_t r; ZK_ASM("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ZK_D uint32_t subc_cc(uint32_t a, uint32_t b) { uint32_t r; ZK_ASM("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ZK_D uint32_t subc(uint32_t a, uint32_t b) { uint32_t r; ZK_ASM("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ZK_D uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; ZK_ASM("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
ZK_D uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; ZK_ASM("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
ZK_D uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; ZK_ASM("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
ZK_D uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; ZK_ASM("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
ZK_D uint32_t mul_lo(uint32_t a, uint32_t b) { return a * b; }
ZK_D uint32_t mul_hi(uint32_t a, uint32_t b) { return __umulhi(a, b); }
// 32x32->64 multiply + 64-bit add with carry: each block is ONE IMAD.WIDE.U32(.X)
// with a predicate carry after ptxas (checked with cuobjdump -sass).
ZK_D uint64_t mul_wide(uint32_t a, uint32_t b) { uint64_t r; ZK_ASM("mul.wide.u32 %0, %1, %2;" : "=l"(r) : "r"(a), "r"(b)); return r; }
ZK_D uint64_t madw_cc(uint32_t a, uint32_t b, uint64_t c) { uint64_t r; ZK_ASM("{\n\t.reg .u64 p;\n\tmul.wide.u32 p, %1, %2;\n\tadd.cc.u64 %0, %3, p;\n\t}" : "=l"(r) : "r"(a), "r"(b), "l"(c)); return r; }
ZK_D uint64_t madwc_cc(uint32_t a, uint32_t b, uint64_t c) { uint64_t r; ZK_ASM("{\n\t.reg .u64 p;\n\tmul.wide.u32 p, %1, %2;\n\taddc.cc.u64 %0, %3, p;\n\t}" : "=l"(r) : "r"(a), "r"(b), "l"(c)); return r; }
ZK_D uint64_t madwc(uint32_t a, uint32_t b, uint64_t c) { uint64_t r; ZK_ASM("{\n\t.reg .u64 p;\n\tmul.wide.u32 p, %1, %2;\n\taddc.u64 %0, %3, p;\n\t}" : "=l"(r) : "r"(a), "r"(b), "l"(c)); return r; }
#else
namespace detail { inline uint32_t& cf() { static thread_local uint32_t f = 0; return f; } }
inline uint32_t add_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a + b; detail::cf() = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t addc_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a + b + detail::cf(); detail::cf() = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t addc(uint32_t a, uint32_t b) { return a + b + detail::cf(); }
inline uint32_t sub_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a - b; detail::cf() = (uint32_t)((t >> 32) & 1); return (uint32_t)t; }
inline uint32_t subc_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a - b - detail::cf(); detail::cf() = (uint32_t)((t >> 32) & 1); return (uint32_t)t; }
inline uint32_t subc(uint32_t a, uint32_t b) { return a - b - detail::cf(); }
inline uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return add_cc((uint32_t)((uint64_t)a * b), c); }
inline uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return addc_cc((uint32_t)((uint64_t)a * b), c); }
inline uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { return addc_cc((uint32_t)(((uint64_t)a * b) >> 32), c); }
inline uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { return addc((uint32_t)(((uint64_t)a * b) >> 32), c); }
inline uint32_t mul_lo(uint32_t a, uint32_t b) { return a * b; }
inline uint32_t mul_hi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
inline uint64_t mul_wide(uint32_t a, uint32_t b) { return (uint64_t)a * b; }
inline uint64_t madw_cc(uint32_t a, uint32_t b, uint64_t c) { unsigned __int128 t = (unsigned __int128)c + (uint64_t)a * b; detail::cf() = (uint32_t)(t >> 64); return (uint64_t)t; }
inline uint64_t madwc_cc(uint32_t a, uint32_t b, uint64_t c) { unsigned __int128 t = (unsigned __int128)c + (uint64_t)a * b + detail::cf(); detail::cf() = (uint32_t)(t >> 64); return (uint64_t)t; }
inline uint64_t madwc(uint32_t a, uint32_t b, uint64_t c) { return c + (uint64_t)a * b + detail::cf(); }
#endif
ZK_HD uint32_t lo32(uint64_t x) { return (uint32_t)x; }
ZK_HD uint32_t hi32(uint64_t x) { return (uint32_t)(x >> 32); }
ZK_HD uint64_t pack64(uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; }

// ------------------------------------------------------------ field params
// SURVEY.md App. B.  P = modulus, R2 = R^2 mod p, ONE = R mod p,
// INV = -p^{-1} mod 2^32.  All as 8 x u32, little-endian.
struct Bn254Fr {
    static constexpr int ID = 0;
    static constexpr bool SLACK3P = true;  // 3p < 2^256: values up to 3p fit the 8 limbs (unreduced operands)
    static constexpr uint32_t INV = 0xefffffffu;
    ZK_HD static constexpr uint32_t P(int i) {
        constexpr uint32_t v[8] = {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
        return v[i];
    }
    ZK_HD static constexpr uint32_t R2(int i) {
        constexpr uint32_t v[8] = {0xae216da7u, 0x1bb8e645u, 0xe35c59e3u, 0x53fe3ab1u, 0x53bb8085u, 0x8c49833du, 0x7f4e44a5u, 0x0216d0b1u};
        return v[i];
    }
    ZK_HD static constexpr uint32_t ONE(int i) {
        constexpr uint32_t v[8] = {0x4ffffffbu, 0xac96341cu, 0x9f60cd29u, 0x36fc7695u, 0x7879462eu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
        return v[i];
    }
};
struct Bn254Fq {
    static constexpr int ID = 1;
    static constexpr bool SLACK3P = true;
    static constexpr uint32_t INV = 0xe4866389u;
    ZK_HD static constexpr uint32_t P(int i) {
        constexpr uint32_t v[8] = {0xd87cfd47u, 0x3c208c16u, 0x6871ca8du, 0x97816a91u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
        return v[i];
    }
    ZK_HD static constexpr uint32_t R2(int i) {
        constexpr uint32_t v[8] = {0x538afa89u, 0xf32cfc5bu, 0xd44501fbu, 0xb5e71911u, 0x0a417ff6u, 0x47ab1effu, 0xcab8351fu, 0x06d89f71u};
        return v[i];
    }
    ZK_HD static constexpr uint32_t ONE(int i) {
        constexpr uint32_t v[8] = {0xc58f0d9du, 0xd35d438du, 0xf5c70b3du, 0x0a78eb28u, 0x7879462cu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
        return v[i];
    }
};
struct Bls12381Fr {
    static constexpr int ID = 2;
    static constexpr bool SLACK3P = false;  // 255-bit modulus: only 2p < 2^256
    static constexpr uint32_t INV = 0xffffffffu;
    ZK_HD static constexpr uint32_t P(int i) {
        constexpr uint32_t v[8] = {0x00000001u, 0xffffffffu, 0xfffe5bfeu, 0x53bda402u, 0x09a1d805u, 0x3339d808u, 0x299d7d48u, 0x73eda753u};
        return v[i];
    }
    ZK_HD static constexpr uint32_t R2(int i) {
        constexpr uint32_t v[8] = {0xf3f29c6du, 0xc999e990u, 0x87925c23u, 0x2b6cedcbu, 0x7254398fu, 0x05d31496u, 0x9f59ff11u, 0x0748d9d9u};
        return v[i];
    }
    ZK_HD static constexpr uint32_t ONE(int i) {
        constexpr uint32_t v[8] = {0xfffffffeu, 0x00000001u, 0x00034802u, 0x5884b7fau, 0xecbc4ff5u, 0x998c4fefu, 0xacc5056fu, 0x1824b159u};
        return v[i];
    }
};

// ------------------------------------------------------------ the element
struct alignas(16) Fe {
    uint32_t l[8];
};

// ------------------------------------------------ 64-bit carry helpers
#if defined(__CUDA_ARCH__)
ZK_D uint64_t addc64(uint64_t a, uint64_t b) { uint64_t r; ZK_ASM("addc.u64 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
#else
inline uint64_t addc64(uint64_t a, uint64_t b) { return a + b + detail::cf(); }
#endif

// Multiplication by a value fixed for a whole launch (the round challenge r):
// t[i] = r * 2^(32 i + 64) * R^-1 mod p as plain 8 x u32 integers, made on the
// host once per round.  x*r*R^-1 = (sum_i x_i * t[i]) * 2^-64, i.e. 64 wide
// multiply-accumulates WITHOUT interleaved reduction plus two reduction rows
// (16 more) instead of the 136 of a general Montgomery product.
struct FixedMul {
    uint32_t t[8][8];
};

// The same challenge for the tensor-core fold (tcfold.cuh): the two u8 B operands of tcgen05.mma kind::i8, [0] for the
// "a" rows (1 - r), [1] for the "b" rows (r).  B[n][k] = byte n of T_k with T_k = (1 - r) 2^(8 k + 32) mod p resp.
// r 2^(8 k + 32) mod p, stored K-major without swizzle: byte offset (k / 16) * 512 + n * 16 + k % 16.
struct alignas(16) TcFoldMats {
    uint8_t b[2][1024];
};

// 512-bit (+32 guard bits) integer accumulator for sums of raw products
// a*b of Montgomery residues; one Montgomery reduction at the very end.
struct Wide {
    uint32_t l[17];
};

template <class F>
struct Field {
    ZK_HD static Fe zero() {
        Fe r;
#pragma unroll
        for (int i = 0; i < 8; ++i) r.l[i] = 0;
        return r;
    }
    ZK_HD static Fe one() {
        Fe r;
#pragma unroll
        for (int i = 0; i < 8; ++i) r.l[i] = F::ONE(i);
        return r;
    }
    ZK_HD static Fe r2() {
        Fe r;
#pragma unroll
        for (int i = 0; i < 8; ++i) r.l[i] = F::R2(i);
        return r;
    }
    ZK_HD static bool is_zero(const Fe& a) {
        uint32_t o = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) o |= a.l[i];
        return o == 0;
    }

    // r = a - p if a >= p else a   (a < 2p)
    ZK_HD static Fe reduce_once(const Fe& a) {
        Fe t;
        t.l[0] = sub_cc(a.l[0], F::P(0));
#pragma unroll
        for (int i = 1; i < 8; ++i) t.l[i] = subc_cc(a.l[i], F::P(i));
        uint32_t borrow = subc(0u, 0u);  // 0xffffffff if a < p
        Fe r;
#pragma unroll
        for (int i = 0; i < 8; ++i) r.l[i] = borrow ? a.l[i] : t.l[i];
        return r;
    }
    // a + b mod p (a, b < p; 2p < 2^256 for all three moduli so no carry-out)
    ZK_HD static Fe add(const Fe& a, const Fe& b) {
        Fe s;
        s.l[0] = add_cc(a.l[0], b.l[0]);
#pragma unroll
        for (int i = 1; i < 7; ++i) s.l[i] = addc_cc(a.l[i], b.l[i]);
        s.l[7] = addc(a.l[7], b.l[7]);
        return reduce_once(s);
    }
    // a - b mod p
    ZK_HD static Fe sub(const Fe& a, const Fe& b) {
        Fe d;
        d.l[0] = sub_cc(a.l[0], b.l[0]);
#pragma unroll
        for (int i = 1; i < 8; ++i) d.l[i] = subc_cc(a.l[i], b.l[i]);
        uint32_t m = subc(0u, 0u);  // all ones if a < b
        Fe r;
        r.l[0] = add_cc(d.l[0], F::P(0) & m);
#pragma unroll
        for (int i = 1; i < 7; ++i) r.l[i] = addc_cc(d.l[i], F::P(i) & m);
        r.l[7] = addc(d.l[7], F::P(7) & m);
        return r;
    }
    ZK_HD static Fe neg(const Fe& a) { return sub(zero(), a); }
    // 2a mod p
    ZK_HD static Fe dbl(const Fe& a) { return add(a, a); }

    // ---- Montgomery product, a*b*R^-1 mod p, fully reduced ----------------
    // acc[0..7] += x[0,2,4,6] * y at 64-bit aligned limb pairs, one carry chain;
    // leaves the carry flag set for the caller.
    ZK_HD static void cmad(uint32_t* acc, const uint32_t* x, uint32_t y) {
        acc[0] = mad_lo_cc(x[0], y, acc[0]);
        acc[1] = madc_hi_cc(x[0], y, acc[1]);
#pragma unroll
        for (int j = 2; j < 8; j += 2) {
            acc[j] = madc_lo_cc(x[j], y, acc[j]);
            acc[j + 1] = madc_hi_cc(x[j], y, acc[j + 1]);
        }
    }
    // same with the modulus' limbs (compile-time constants -> immediates)
    template <int OFF>
    ZK_HD static void cmad_p(uint32_t* acc, uint32_t y) {
        acc[0] = mad_lo_cc(F::P(OFF), y, acc[0]);
        acc[1] = madc_hi_cc(F::P(OFF), y, acc[1]);
#pragma unroll
        for (int j = 2; j < 8; j += 2) {
            acc[j] = madc_lo_cc(F::P(OFF + j), y, acc[j]);
            acc[j + 1] = madc_hi_cc(F::P(OFF + j), y, acc[j + 1]);
        }
    }
    // One CIOS row.  On entry T = ev + (od << 32) where od[] is the PREVIOUS
    // row's "even" accumulator whose limb 0 is zero, limb 1 is a stray limb to
    // be folded into ev[0], and limbs 2..7 are the shifted-down content.
    ZK_HD static void row(uint32_t* ev, uint32_t* od, const uint32_t* a, uint32_t bi, bool first) {
        if (first) {
#pragma unroll
            for (int j = 0; j < 8; j += 2) {
                od[j] = mul_lo(a[j + 1], bi);
                od[j + 1] = mul_hi(a[j + 1], bi);
                ev[j] = mul_lo(a[j], bi);
                ev[j + 1] = mul_hi(a[j], bi);
            }
        } else {
            ev[0] = add_cc(ev[0], od[1]);
            // od = (od >> 64) + a_odd * bi (+ carry of the stray-limb add)
#pragma unroll
            for (int j = 0; j < 6; j += 2) {
                od[j] = madc_lo_cc(a[j + 1], bi, od[j + 2]);
                od[j + 1] = madc_hi_cc(a[j + 1], bi, od[j + 3]);
            }
            od[6] = madc_lo_cc(a[7], bi, 0u);
            od[7] = madc_hi(a[7], bi, 0u);
            cmad(ev, a, bi);
            od[7] = addc(od[7], 0u);
        }
        uint32_t m = mul_lo(ev[0], F::INV);
        cmad_p<1>(od, m);
        cmad_p<0>(ev, m);
        od[7] = addc(od[7], 0u);
    }
    // 32-bit-limb variant (kept for the IMAD vs IMAD.WIDE microbenchmark)
    ZK_HD static Fe mul_split(const Fe& a, const Fe& b) {
        uint32_t ev[8], od[8];
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
            row(ev, od, a.l, b.l[i], i == 0);
            row(od, ev, a.l, b.l[i + 1], false);
        }
        // result = ev + (od >> 32)
        Fe r;
        r.l[0] = add_cc(ev[0], od[1]);
#pragma unroll
        for (int i = 1; i < 7; ++i) r.l[i] = addc_cc(ev[i], od[i + 1]);
        r.l[7] = addc(ev[7], 0u);
        return reduce_once(r);
    }
    // ---- the same rows on 4 x u64 accumulators (default multiplier) ---------
    // ev/od hold limb PAIRS: ev[k] = limbs (2k, 2k+1) of the even column sum,
    // od[k] = limbs (2k+1, 2k+2) of the odd one (T = ev + (od << 32)).
    template <int OFF>
    ZK_HD static void cmadw_p(uint64_t* acc, uint32_t m) {
        acc[0] = madw_cc(F::P(OFF), m, acc[0]);
        acc[1] = madwc_cc(F::P(OFF + 2), m, acc[1]);
        acc[2] = madwc_cc(F::P(OFF + 4), m, acc[2]);
        acc[3] = madwc_cc(F::P(OFF + 6), m, acc[3]);
    }
    ZK_HD static void roww(uint64_t* ev, uint64_t* od, const uint32_t* a, uint32_t bi, bool first) {
        if (first) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                od[k] = mul_wide(a[2 * k + 1], bi);
                ev[k] = mul_wide(a[2 * k], bi);
            }
        } else {
            // od is the previous row's even accumulator: limb 0 is zero, limb 1
            // (hi32(od[0])) is the stray limb; its carry feeds the od chain.
            uint32_t e0 = add_cc(lo32(ev[0]), hi32(od[0]));
            od[0] = madwc_cc(a[1], bi, od[1]);
            od[1] = madwc_cc(a[3], bi, od[2]);
            od[2] = madwc_cc(a[5], bi, od[3]);
            od[3] = madwc(a[7], bi, 0ull);
            ev[0] = pack64(e0, hi32(ev[0]));
            ev[0] = madw_cc(a[0], bi, ev[0]);
            ev[1] = madwc_cc(a[2], bi, ev[1]);
            ev[2] = madwc_cc(a[4], bi, ev[2]);
            ev[3] = madwc_cc(a[6], bi, ev[3]);
            od[3] = pack64(lo32(od[3]), addc(hi32(od[3]), 0u));
        }
        uint32_t m = mul_lo(lo32(ev[0]), F::INV);
        cmadw_p<1>(od, m);  // carry-out provably zero: od << 32 <= T < 2^288
        cmadw_p<0>(ev, m);
        od[3] = pack64(lo32(od[3]), addc(hi32(od[3]), 0u));
    }
    ZK_HD static Fe mul(const Fe& a, const Fe& b) {
        uint64_t ev[4], od[4];
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
            roww(ev, od, a.l, b.l[i], i == 0);
            roww(od, ev, a.l, b.l[i + 1], false);
        }
        Fe r;  // ev + (od >> 32); od's limb 0 is zero
        r.l[0] = add_cc(lo32(ev[0]), hi32(od[0]));
        r.l[1] = addc_cc(hi32(ev[0]), lo32(od[1]));
        r.l[2] = addc_cc(lo32(ev[1]), hi32(od[1]));
        r.l[3] = addc_cc(hi32(ev[1]), lo32(od[2]));
        r.l[4] = addc_cc(lo32(ev[2]), hi32(od[2]));
        r.l[5] = addc_cc(hi32(ev[2]), lo32(od[3]));
        r.l[6] = addc_cc(lo32(ev[3]), hi32(od[3]));
        r.l[7] = addc(hi32(ev[3]), 0u);
        return reduce_once(r);
    }
    ZK_HD static Fe sqr(const Fe& a) { return mul(a, a); }

    // ---- fixed-multiplicand product: x * r * R^-1 mod p (see FixedMul) ------
    ZK_HD static Fe mul_fixed(const Fe& x, const FixedMul& T) {
        // S = sum_i x_i * T_i < 2^289 in two interleaved accumulators:
        // ev[k] = limbs (2k, 2k+1), od[k] = limbs (2k+1, 2k+2); word 4 collects carries.
        uint64_t ev[4], od[4];
        uint32_t ce = 0, co = 0;  // carries out of word 3 of either accumulator (limb 8 resp. limb 9), at most 7 each
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            ev[k] = mul_wide(T.t[0][2 * k], x.l[0]);
            od[k] = mul_wide(T.t[0][2 * k + 1], x.l[0]);
        }
#pragma unroll
        for (int i = 1; i < 8; ++i) {
            ev[0] = madw_cc(T.t[i][0], x.l[i], ev[0]);
            ev[1] = madwc_cc(T.t[i][2], x.l[i], ev[1]);
            ev[2] = madwc_cc(T.t[i][4], x.l[i], ev[2]);
            ev[3] = madwc_cc(T.t[i][6], x.l[i], ev[3]);
            ce = addc(ce, 0u);
            od[0] = madw_cc(T.t[i][1], x.l[i], od[0]);
            od[1] = madwc_cc(T.t[i][3], x.l[i], od[1]);
            od[2] = madwc_cc(T.t[i][5], x.l[i], od[2]);
            od[3] = madwc_cc(T.t[i][7], x.l[i], od[3]);
            co = addc(co, 0u);
        }
        // Two Montgomery rows (divide by 2^64) on the split accumulators, so every product stays on a 64-bit aligned
        // limb pair (no re-pairing moves).  Row 1 clears limb 0 = lo32(ev[0]); row 2 clears limb 1 =
        // hi32(ev[0]) + lo32(od[0]) (mod 2^32): m1 * p sits one limb higher, so its even-limb products go to the
        // odd accumulator and its odd-limb products to the even one, one word up.
        const uint32_t m0 = mul_lo(lo32(ev[0]), F::INV);
        ev[0] = madw_cc(F::P(0), m0, ev[0]);
        ev[1] = madwc_cc(F::P(2), m0, ev[1]);
        ev[2] = madwc_cc(F::P(4), m0, ev[2]);
        ev[3] = madwc_cc(F::P(6), m0, ev[3]);
        ce = addc(ce, 0u);
        od[0] = madw_cc(F::P(1), m0, od[0]);
        od[1] = madwc_cc(F::P(3), m0, od[1]);
        od[2] = madwc_cc(F::P(5), m0, od[2]);
        od[3] = madwc_cc(F::P(7), m0, od[3]);
        co = addc(co, 0u);
        const uint32_t m1 = mul_lo(hi32(ev[0]) + lo32(od[0]), F::INV);
        od[0] = madw_cc(F::P(0), m1, od[0]);
        od[1] = madwc_cc(F::P(2), m1, od[1]);
        od[2] = madwc_cc(F::P(4), m1, od[2]);
        od[3] = madwc_cc(F::P(6), m1, od[3]);
        co = addc(co, 0u);
        uint64_t e4 = ce;  // limbs (8, 9) of the even accumulator
        ev[1] = madw_cc(F::P(1), m1, ev[1]);
        ev[2] = madwc_cc(F::P(3), m1, ev[2]);
        ev[3] = madwc_cc(F::P(5), m1, ev[3]);
        e4 = madwc(F::P(7), m1, e4);
        // result = (ev + (od << 32)) >> 64 < p (1 + 2^-29): S < 2^35 p and m0 + m1 2^32 < 2^64.  Limb 1 is zero mod 2^32;
        // its carry (1 unless both halves are zero) enters limb 2.
        Fe r;
        (void)add_cc(hi32(ev[0]), lo32(od[0]));
        r.l[0] = addc_cc(lo32(ev[1]), hi32(od[0]));
        r.l[1] = addc_cc(hi32(ev[1]), lo32(od[1]));
        r.l[2] = addc_cc(lo32(ev[2]), hi32(od[1]));
        r.l[3] = addc_cc(hi32(ev[2]), lo32(od[2]));
        r.l[4] = addc_cc(lo32(ev[3]), hi32(od[2]));
        r.l[5] = addc_cc(hi32(ev[3]), lo32(od[3]));
        r.l[6] = addc_cc(lo32(e4), hi32(od[3]));
        r.l[7] = addc(hi32(e4), co);
        return reduce_once(r);
    }
    // b - a + p in (0, 2p): an UNREDUCED difference (no compare / select), fine as the x of mul_fixed and as an
    // operand of mac_wide, which only need x < 2^256
    ZK_HD static Fe sub_lazy(const Fe& b, const Fe& a) {
        Fe d;
        d.l[0] = sub_cc(b.l[0], a.l[0]);
#pragma unroll
        for (int i = 1; i < 8; ++i) d.l[i] = subc_cc(b.l[i], a.l[i]);
        d.l[0] = add_cc(d.l[0], F::P(0));  // the borrow out of limb 7 cancels against the carry out of this chain
#pragma unroll
        for (int i = 1; i < 7; ++i) d.l[i] = addc_cc(d.l[i], F::P(i));
        d.l[7] = addc(d.l[7], F::P(7));
        return d;
    }
    // 2*hi - lo + p in (0, 3p): the unreduced value of the pair's line at t = 2 (needs 3p < 2^256)
    ZK_HD static Fe line2_lazy(const Fe& lo, const Fe& hi) {
        Fe d;
        d.l[0] = add_cc(hi.l[0], hi.l[0]);
#pragma unroll
        for (int i = 1; i < 7; ++i) d.l[i] = addc_cc(hi.l[i], hi.l[i]);
        d.l[7] = addc(hi.l[7], hi.l[7]);
        d.l[0] = sub_cc(d.l[0], lo.l[0]);
#pragma unroll
        for (int i = 1; i < 7; ++i) d.l[i] = subc_cc(d.l[i], lo.l[i]);
        d.l[7] = subc(d.l[7], lo.l[7]);
        d.l[0] = add_cc(d.l[0], F::P(0));
#pragma unroll
        for (int i = 1; i < 7; ++i) d.l[i] = addc_cc(d.l[i], F::P(i));
        d.l[7] = addc(d.l[7], F::P(7));
        return d;
    }
    // a + r*(b - a) with the fixed-multiplicand product
    ZK_HD static Fe fold_fixed(const Fe& a, const Fe& b, const FixedMul& T) { return add(a, mul_fixed(sub_lazy(b, a), T)); }

    // ---- lazy sums of products ---------------------------------------------
    ZK_HD static Wide wide_zero() {
        Wide w;
#pragma unroll
        for (int i = 0; i < 17; ++i) w.l[i] = 0;
        return w;
    }
    // The raw 512-bit product a * b as two interleaved column sums:
    // ev[k] = limbs (2k, 2k+1), od[k] = limbs (2k+1, 2k+2); a*b = ev + (od << 32).
    ZK_HD static void mul_wide16(const Fe& a, const Fe& b, uint64_t* ev, uint64_t* od) {
        // schoolbook rows in order
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            ev[k] = mul_wide(a.l[2 * k], b.l[0]);
            od[k] = mul_wide(a.l[2 * k + 1], b.l[0]);
            ev[k + 4] = 0;
            od[k + 4] = 0;
        }
#pragma unroll
        for (int j = 1; j < 8; ++j) {
            if (j & 1) {  // a_even * b_j -> od[(j-1)/2 ..], a_odd * b_j -> ev[(j+1)/2 ..]
                const int o = (j - 1) / 2, e = (j + 1) / 2;
                od[o] = madw_cc(a.l[0], b.l[j], od[o]);
                od[o + 1] = madwc_cc(a.l[2], b.l[j], od[o + 1]);
                od[o + 2] = madwc_cc(a.l[4], b.l[j], od[o + 2]);
                od[o + 3] = madwc_cc(a.l[6], b.l[j], od[o + 3]);
                if (o + 4 < 8) od[o + 4] = addc64(od[o + 4], 0ull);
                ev[e] = madw_cc(a.l[1], b.l[j], ev[e]);
                ev[e + 1] = madwc_cc(a.l[3], b.l[j], ev[e + 1]);
                ev[e + 2] = madwc_cc(a.l[5], b.l[j], ev[e + 2]);
                ev[e + 3] = madwc_cc(a.l[7], b.l[j], ev[e + 3]);
                if (e + 4 < 8) ev[e + 4] = addc64(ev[e + 4], 0ull);
            } else {
                const int e = j / 2;
                ev[e] = madw_cc(a.l[0], b.l[j], ev[e]);
                ev[e + 1] = madwc_cc(a.l[2], b.l[j], ev[e + 1]);
                ev[e + 2] = madwc_cc(a.l[4], b.l[j], ev[e + 2]);
                ev[e + 3] = madwc_cc(a.l[6], b.l[j], ev[e + 3]);
                if (e + 4 < 8) ev[e + 4] = addc64(ev[e + 4], 0ull);
                od[e] = madw_cc(a.l[1], b.l[j], od[e]);
                od[e + 1] = madwc_cc(a.l[3], b.l[j], od[e + 1]);
                od[e + 2] = madwc_cc(a.l[5], b.l[j], od[e + 2]);
                od[e + 3] = madwc_cc(a.l[7], b.l[j], od[e + 3]);
                if (e + 4 < 8) od[e + 4] = addc64(od[e + 4], 0ull);
            }
        }
    }
    // acc += ev + (od << 32)
    ZK_HD static void wide_add(Wide& acc, const uint64_t* ev, const uint64_t* od) {
        acc.l[0] = add_cc(acc.l[0], lo32(ev[0]));
        acc.l[1] = addc_cc(acc.l[1], hi32(ev[0]));
#pragma unroll
        for (int k = 1; k < 8; ++k) {
            acc.l[2 * k] = addc_cc(acc.l[2 * k], lo32(ev[k]));
            acc.l[2 * k + 1] = addc_cc(acc.l[2 * k + 1], hi32(ev[k]));
        }
        acc.l[16] = addc(acc.l[16], 0u);
        acc.l[1] = add_cc(acc.l[1], lo32(od[0]));
        acc.l[2] = addc_cc(acc.l[2], hi32(od[0]));
#pragma unroll
        for (int k = 1; k < 8; ++k) {
            acc.l[2 * k + 1] = addc_cc(acc.l[2 * k + 1], lo32(od[k]));
            if (k < 7) acc.l[2 * k + 2] = addc_cc(acc.l[2 * k + 2], hi32(od[k]));
        }
        acc.l[16] = addc(acc.l[16], hi32(od[7]));
    }
    // acc += a * b (plain 512-bit integer product; up to 2^36 products fit)
    ZK_HD static void mac_wide(Wide& acc, const Fe& a, const Fe& b) {
        uint64_t ev[8], od[8];
        mul_wide16(a, b, ev, od);
        wide_add(acc, ev, od);
    }
    // x * R^-1 mod p for any 256-bit x (Montgomery reduction without a multiplication): eight rows of
    // "m = t0 * inv; t = (t + m*p) >> 32".  Result < p + 1, fully reduced on return.
    ZK_HD static Fe redc256(const Fe& x) {
        uint32_t s[9];
#pragma unroll
        for (int j = 0; j < 8; ++j) s[j] = x.l[j];
        s[8] = 0;
#pragma unroll
        for (int row = 0; row < 8; ++row) {
            const uint32_t m = mul_lo(s[0], F::INV);
            s[0] = mad_lo_cc(F::P(0), m, s[0]);
            s[1] = madc_hi_cc(F::P(0), m, s[1]);
#pragma unroll
            for (int j = 2; j < 8; j += 2) {
                s[j] = madc_lo_cc(F::P(j), m, s[j]);
                s[j + 1] = madc_hi_cc(F::P(j), m, s[j + 1]);
            }
            s[8] = addc(s[8], 0u);
            s[1] = mad_lo_cc(F::P(1), m, s[1]);
            s[2] = madc_hi_cc(F::P(1), m, s[2]);
#pragma unroll
            for (int j = 3; j < 8; j += 2) {
                s[j] = madc_lo_cc(F::P(j), m, s[j]);
                s[j + 1] = madc_hi_cc(F::P(j), m, s[j + 1]);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) s[j] = s[j + 1];  // s[0] is zero now: divide by 2^32
            s[8] = 0;
        }
        Fe r;
#pragma unroll
        for (int j = 0; j < 8; ++j) r.l[j] = s[j];
        return reduce_once(r);
    }
    // any 256-bit x -> x mod p by repeated conditional subtraction (2^256 < 6p for BN254, < 3p for BLS12-381)
    ZK_HD static Fe mod_p(const Fe& x) {
        Fe r = x;
#pragma unroll
        for (int k = 0; k < (F::SLACK3P ? 5 : 2); ++k) r = reduce_once(r);
        return r;
    }
    // acc * R^-1 mod p, fully reduced: acc = A0 + A1*2^256 + A2*2^512 ->
    // A0*R^-1 + A1 + A2*R  =  redc256(A0) + (A1 mod p) + mul(R^2, A2)   (mul = Montgomery product; the chunk
    // A2 goes in as the limb-iterated operand, for which mul() only needs b < 2^256).  A2 is zero until about
    // 27 products have been accumulated, so short sums skip the multiplication.
    ZK_HD static Fe reduce_wide(const Wide& acc) {
        Fe a0, a1;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            a0.l[i] = acc.l[i];
            a1.l[i] = acc.l[8 + i];
        }
        Fe r = add(redc256(a0), mod_p(a1));
        if (acc.l[16] != 0) {
            Fe a2 = zero();
            a2.l[0] = acc.l[16];
            r = add(r, mul(r2(), a2));
        }
        return r;
    }


    ZK_HD static Fe to_mont(const Fe& canon) { return mul(canon, r2()); }
    ZK_HD static Fe from_mont(const Fe& m) {
        Fe o = zero();
        o.l[0] = 1;
        return mul(m, o);
    }
    // small unsigned integer -> Montgomery form
    ZK_HD static Fe from_u32(uint32_t x) {
        Fe c = zero();
        c.l[0] = x;
        return to_mont(c);
    }
    // a / 2 mod p (also correct on Montgomery residues: halving commutes with the factor R)
    ZK_HD static Fe half(const Fe& a) {
        const uint32_t m = 0u - (a.l[0] & 1u);  // odd: add p first (a + p < 2^256 for every modulus here)
        Fe t;
        t.l[0] = add_cc(a.l[0], F::P(0) & m);
#pragma unroll
        for (int i = 1; i < 7; ++i) t.l[i] = addc_cc(a.l[i], F::P(i) & m);
        t.l[7] = addc(a.l[7], F::P(7) & m);
        Fe r;
#pragma unroll
        for (int i = 0; i < 7; ++i) r.l[i] = (t.l[i] >> 1) | (t.l[i + 1] << 31);
        r.l[7] = t.l[7] >> 1;
        return r;
    }
    // a + r*(b - a): the fold of multilinear_polynomial_evaluation.rs:59
    ZK_HD static Fe fold(const Fe& a, const Fe& b, const Fe& r) { return add(a, mul(r, sub(b, a))); }
};

}  // namespace zkb
