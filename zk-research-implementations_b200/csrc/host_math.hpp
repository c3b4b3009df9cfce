// host_math.hpp -- the host-resident part of the path: the Fiat-Shamir
// transcript (Keccak-256), canonical serialisation, and the <=5-point
// univariate interpolation of round messages.  These stay on the host by
// design (BASELINE.json north_star): per round only (d+1) field elements come
// back from the GPU and one challenge goes out as a kernel argument.
//
// Mirrors fiat_shamir/src/fiat_shamir_transcript.rs:5-37 and
// univariate_polynomial/src/univariate_polynomial_dense.rs:14-26,48-74.
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

#include "launch.h"

namespace zkb {

// Keccak-f[1600], fully unrolled with the 25 lanes in locals (two rounds per loop iteration so that no lane moves).
// Built twice: for the baseline x86-64 ISA and for BMI1/BMI2 (andn, rorx); picked once at run time.  The transcript
// hashes at most a few hundred bytes per round, but whole tables when the reference absorbs them
// (sum_check_protocol.rs:27, the output layer in gkr_protocol.rs:233), where this permutation is the entire cost.
namespace keccak_detail {
static inline uint64_t rotl(uint64_t x, int n) { return (x << n) | (x >> (64 - n)); }
#define KECCAK_ROUND(A,E,rc) do { \
  uint64_t c0=A##0^A##5^A##10^A##15^A##20, c1=A##1^A##6^A##11^A##16^A##21, c2=A##2^A##7^A##12^A##17^A##22, c3=A##3^A##8^A##13^A##18^A##23, c4=A##4^A##9^A##14^A##19^A##24; \
  uint64_t d0=c4^rotl(c1,1), d1=c0^rotl(c2,1), d2=c1^rotl(c3,1), d3=c2^rotl(c4,1), d4=c3^rotl(c0,1); \
  { uint64_t b0=(A##0^d0), b1=rotl((A##6^d1),44), b2=rotl((A##12^d2),43), b3=rotl((A##18^d3),21), b4=rotl((A##24^d4),14); \
    E##0 = b0 ^ (~b1 & b2) ^ (rc); \
    E##1 = b1 ^ (~b2 & b3); \
    E##2 = b2 ^ (~b3 & b4); \
    E##3 = b3 ^ (~b4 & b0); \
    E##4 = b4 ^ (~b0 & b1); \
  } \
  { uint64_t b0=rotl((A##3^d3),28), b1=rotl((A##9^d4),20), b2=rotl((A##10^d0),3), b3=rotl((A##16^d1),45), b4=rotl((A##22^d2),61); \
    E##5 = b0 ^ (~b1 & b2); \
    E##6 = b1 ^ (~b2 & b3); \
    E##7 = b2 ^ (~b3 & b4); \
    E##8 = b3 ^ (~b4 & b0); \
    E##9 = b4 ^ (~b0 & b1); \
  } \
  { uint64_t b0=rotl((A##1^d1),1), b1=rotl((A##7^d2),6), b2=rotl((A##13^d3),25), b3=rotl((A##19^d4),8), b4=rotl((A##20^d0),18); \
    E##10 = b0 ^ (~b1 & b2); \
    E##11 = b1 ^ (~b2 & b3); \
    E##12 = b2 ^ (~b3 & b4); \
    E##13 = b3 ^ (~b4 & b0); \
    E##14 = b4 ^ (~b0 & b1); \
  } \
  { uint64_t b0=rotl((A##4^d4),27), b1=rotl((A##5^d0),36), b2=rotl((A##11^d1),10), b3=rotl((A##17^d2),15), b4=rotl((A##23^d3),56); \
    E##15 = b0 ^ (~b1 & b2); \
    E##16 = b1 ^ (~b2 & b3); \
    E##17 = b2 ^ (~b3 & b4); \
    E##18 = b3 ^ (~b4 & b0); \
    E##19 = b4 ^ (~b0 & b1); \
  } \
  { uint64_t b0=rotl((A##2^d2),62), b1=rotl((A##8^d3),55), b2=rotl((A##14^d4),39), b3=rotl((A##15^d0),41), b4=rotl((A##21^d1),2); \
    E##20 = b0 ^ (~b1 & b2); \
    E##21 = b1 ^ (~b2 & b3); \
    E##22 = b2 ^ (~b3 & b4); \
    E##23 = b3 ^ (~b4 & b0); \
    E##24 = b4 ^ (~b0 & b1); \
  } \
} while (0)

#define ZK_KECCAK_BODY \
    static const uint64_t RC[24] = { \
        0x0000000000000001ull, 0x0000000000008082ull, 0x800000000000808aull, 0x8000000080008000ull, 0x000000000000808bull, \
        0x0000000080000001ull, 0x8000000080008081ull, 0x8000000000008009ull, 0x000000000000008aull, 0x0000000000000088ull, \
        0x0000000080008009ull, 0x000000008000000aull, 0x000000008000808bull, 0x800000000000008bull, 0x8000000000008089ull, \
        0x8000000000008003ull, 0x8000000000008002ull, 0x8000000000000080ull, 0x000000000000800aull, 0x800000008000000aull, \
        0x8000000080008081ull, 0x8000000000008080ull, 0x0000000080000001ull, 0x8000000080008008ull}; \
    uint64_t a0 = st[0], a1 = st[1], a2 = st[2], a3 = st[3], a4 = st[4], a5 = st[5], a6 = st[6], a7 = st[7], a8 = st[8], a9 = st[9], \
             a10 = st[10], a11 = st[11], a12 = st[12], a13 = st[13], a14 = st[14], a15 = st[15], a16 = st[16], a17 = st[17], \
             a18 = st[18], a19 = st[19], a20 = st[20], a21 = st[21], a22 = st[22], a23 = st[23], a24 = st[24]; \
    uint64_t e0, e1, e2, e3, e4, e5, e6, e7, e8, e9, e10, e11, e12, e13, e14, e15, e16, e17, e18, e19, e20, e21, e22, e23, e24; \
    for (int r = 0; r < 24; r += 2) { \
        KECCAK_ROUND(a, e, RC[r]); \
        KECCAK_ROUND(e, a, RC[r + 1]); \
    } \
    st[0] = a0; st[1] = a1; st[2] = a2; st[3] = a3; st[4] = a4; st[5] = a5; st[6] = a6; st[7] = a7; st[8] = a8; st[9] = a9; \
    st[10] = a10; st[11] = a11; st[12] = a12; st[13] = a13; st[14] = a14; st[15] = a15; st[16] = a16; st[17] = a17; \
    st[18] = a18; st[19] = a19; st[20] = a20; st[21] = a21; st[22] = a22; st[23] = a23; st[24] = a24;
inline void permute_plain(uint64_t* st) { ZK_KECCAK_BODY }
#if defined(__x86_64__) && defined(__GNUC__)
__attribute__((target("bmi,bmi2"))) inline void permute_bmi(uint64_t* st) { ZK_KECCAK_BODY }
#endif
#undef ZK_KECCAK_BODY
#undef KECCAK_ROUND
}  // namespace keccak_detail
typedef void (*keccak_fn)(uint64_t*);
inline keccak_fn keccak_f1600() {
    static const keccak_fn fn = [] {
#if defined(__x86_64__) && defined(__GNUC__)
        if (__builtin_cpu_supports("bmi") && __builtin_cpu_supports("bmi2")) return (keccak_fn)keccak_detail::permute_bmi;
#endif
        return (keccak_fn)keccak_detail::permute_plain;
    }();
    return fn;
}

// ------------------------------------------------------------- Keccak-256
// sha3 0.10.8 `Keccak256` = Keccak[r=1088,c=512] with the ORIGINAL 0x01 padding
// (not SHA3's 0x06).  Sponge state is 25 lanes; absorption is sequential.
class Keccak256 {
  public:
    Keccak256() { reset(); }
    void reset() {
        std::memset(st_, 0, sizeof st_);
        fill_ = 0;
    }
    void update(const uint8_t* d, size_t n) {
        while (n) {
            if (fill_ == 0 && n >= RATE) {  // whole blocks straight from the input
                absorb(d);
                d += RATE;
                n -= RATE;
                continue;
            }
            size_t take = RATE - fill_;
            if (take > n) take = n;
            std::memcpy(buf_ + fill_, d, take);
            fill_ += take;
            d += take;
            n -= take;
            if (fill_ == RATE) {
                absorb(buf_);
                fill_ = 0;
            }
        }
    }
    // finalize_reset(): digest of everything absorbed, then a fresh sponge
    void finalize_reset(uint8_t out[32]) {
        std::memset(buf_ + fill_, 0, RATE - fill_);
        buf_[fill_] ^= 0x01;
        buf_[RATE - 1] ^= 0x80;
        absorb(buf_);
        std::memcpy(out, st_, 32);
        reset();
    }

    // The sponge as a device continuation needs it (kernels.cuh DtArgs): state with the pending bytes already
    // XORed in, and how many 64-bit words of the rate they occupy.  False if the pending bytes are not whole words.
    bool snapshot(uint64_t st[25], uint32_t* fill_words) const {
        if (fill_ % 8) return false;
        std::memcpy(st, st_, sizeof st_);
        for (size_t i = 0; i < fill_ / 8; ++i) {
            uint64_t lane;
            std::memcpy(&lane, buf_ + 8 * i, 8);
            st[i] ^= lane;
        }
        *fill_words = (uint32_t)(fill_ / 8);
        return true;
    }

  private:
    static constexpr size_t RATE = 136;
    uint64_t st_[25];
    uint8_t buf_[RATE];
    size_t fill_;

    static inline uint64_t rotl(uint64_t x, int n) { return (x << n) | (x >> (64 - n)); }
    void absorb(const uint8_t* blk) {
        for (size_t i = 0; i < RATE / 8; ++i) {
            uint64_t lane;
            std::memcpy(&lane, blk + 8 * i, 8);
            st_[i] ^= lane;
        }
        permute();
    }
    void permute() { keccak_f1600()(st_); }
};

// ------------------------------------------------------- host field helper
// Montgomery-residue arithmetic on the host, 4 x 64-bit limbs with unsigned
// __int128 (the same residues the kernels produce: R = 2^256).  Used for the
// few field operations per round that stay on the host: interpolation of the
// round message, the claim chain, challenge tables, serialisation.
struct HostField {
    const FieldKernels* K = nullptr;
    uint64_t p[4] = {0, 0, 0, 0};
    uint64_t ninv = 0;  // -p^-1 mod 2^64
    Fe one_, r2_;

    static HostField make(const FieldKernels* k) {
        HostField H;
        H.K = k;
        Fe m;
        k->h_modulus(m);
        std::memcpy(H.p, m.l, 32);
        uint64_t x = 1;  // Newton: x = p^-1 mod 2^64
        for (int i = 0; i < 6; ++i) x *= 2 - H.p[0] * x;
        H.ninv = ~x + 1;
        Fe v = H.zero();
        v.l[0] = 1;
        for (int i = 0; i < 256; ++i) v = H.add(v, v);
        H.one_ = v;
        for (int i = 0; i < 256; ++i) v = H.add(v, v);
        H.r2_ = v;
        return H;
    }
    Fe zero() const {
        Fe z;
        for (int i = 0; i < 8; ++i) z.l[i] = 0;
        return z;
    }
    static void ld(const Fe& a, uint64_t v[4]) { std::memcpy(v, a.l, 32); }
    static Fe st(const uint64_t v[4]) {
        Fe r;
        std::memcpy(r.l, v, 32);
        return r;
    }
    bool ge_p(const uint64_t a[4]) const {
        for (int i = 3; i >= 0; --i) {
            if (a[i] > p[i]) return true;
            if (a[i] < p[i]) return false;
        }
        return true;
    }
    void sub_p(uint64_t a[4]) const {
        unsigned __int128 bw = 0;
        for (int i = 0; i < 4; ++i) {
            unsigned __int128 d = (unsigned __int128)a[i] - p[i] - (uint64_t)bw;
            a[i] = (uint64_t)d;
            bw = (d >> 64) & 1;
        }
    }
    Fe add(const Fe& x, const Fe& y) const {
        uint64_t a[4], b[4], r[4];
        ld(x, a);
        ld(y, b);
        unsigned __int128 c = 0;
        for (int i = 0; i < 4; ++i) {
            c += (unsigned __int128)a[i] + b[i];
            r[i] = (uint64_t)c;
            c >>= 64;
        }
        if (c || ge_p(r)) sub_p(r);
        return st(r);
    }
    Fe sub(const Fe& x, const Fe& y) const {
        uint64_t a[4], b[4], r[4];
        ld(x, a);
        ld(y, b);
        unsigned __int128 bw = 0;
        for (int i = 0; i < 4; ++i) {
            unsigned __int128 d = (unsigned __int128)a[i] - b[i] - (uint64_t)bw;
            r[i] = (uint64_t)d;
            bw = (d >> 64) & 1;
        }
        if (bw) {
            unsigned __int128 c = 0;
            for (int i = 0; i < 4; ++i) {
                c += (unsigned __int128)r[i] + p[i];
                r[i] = (uint64_t)c;
                c >>= 64;
            }
        }
        return st(r);
    }
    // Montgomery product (CIOS); x < 2^256 arbitrary, y < p
    Fe mul(const Fe& x, const Fe& y) const {
        uint64_t a[4], b[4];
        ld(x, a);
        ld(y, b);
        uint64_t t[6] = {0, 0, 0, 0, 0, 0};
        for (int i = 0; i < 4; ++i) {
            unsigned __int128 c = 0;
            for (int j = 0; j < 4; ++j) {
                c += (unsigned __int128)b[j] * a[i] + t[j];
                t[j] = (uint64_t)c;
                c >>= 64;
            }
            c += t[4];
            t[4] = (uint64_t)c;
            t[5] = (uint64_t)(c >> 64);
            const uint64_t m = t[0] * ninv;
            c = (unsigned __int128)m * p[0] + t[0];
            c >>= 64;
            for (int j = 1; j < 4; ++j) {
                c += (unsigned __int128)m * p[j] + t[j];
                t[j - 1] = (uint64_t)c;
                c >>= 64;
            }
            c += t[4];
            t[3] = (uint64_t)c;
            t[4] = t[5] + (uint64_t)(c >> 64);
        }
        if (t[4] || ge_p(t)) sub_p(t);
        return st(t);
    }
    Fe to_mont(const Fe& a) const { return mul(a, r2_); }
    Fe from_mont(const Fe& a) const {
        Fe o = zero();
        o.l[0] = 1;
        return mul(a, o);
    }
    Fe from_u64(uint64_t x) const {
        Fe c = zero();
        c.l[0] = (uint32_t)x;
        c.l[1] = (uint32_t)(x >> 32);
        return to_mont(c);
    }
    Fe one() const { return one_; }
    bool is_zero(const Fe& a) const {
        uint32_t o = 0;
        for (int i = 0; i < 8; ++i) o |= a.l[i];
        return o == 0;
    }
    bool eq(const Fe& a, const Fe& b) const { return std::memcmp(a.l, b.l, 32) == 0; }
    Fe modulus() const { return st(p); }
    // a^(p-2)
    Fe inv(const Fe& a) const {
        Fe e = modulus();
        uint64_t bw = 2;  // e = p - 2 with borrow propagation (BLS12-381 Fr's low limb is 1)
        for (int i = 0; i < 8 && bw; ++i) {
            uint64_t d = (uint64_t)e.l[i] - bw;
            e.l[i] = (uint32_t)d;
            bw = (d >> 32) & 1;
        }
        Fe r = one();
        for (int i = 255; i >= 0; --i) {
            r = mul(r, r);
            if ((e.l[i / 32] >> (i % 32)) & 1) r = mul(r, a);
        }
        return r;
    }
    // 256-bit little-endian integer mod p -> Montgomery (from_le_bytes_mod_order)
    Fe from_le_bytes_mod_order(const uint8_t b[32]) const {
        Fe c;
        std::memcpy(c.l, b, 32);
        Fe p = modulus();
        // value < 2^256 < 6p (BN254) / 3p (BLS): subtract p while >= p
        for (;;) {
            bool ge = true;
            for (int i = 7; i >= 0; --i) {
                if (c.l[i] > p.l[i]) break;
                if (c.l[i] < p.l[i]) { ge = false; break; }
            }
            if (!ge) break;
            uint64_t bw = 0;
            for (int i = 0; i < 8; ++i) {
                uint64_t d = (uint64_t)c.l[i] - p.l[i] - bw;
                c.l[i] = (uint32_t)d;
                bw = (d >> 32) & 1;
            }
        }
        return to_mont(c);
    }
    // Sum of up to 2^32 canonical-range residues delivered as 8 zero-extended u64 limbs
    // (the NCCL allreduce operand): carry-propagate, then reduce mod p.
    Fe from_wide_limbs(const unsigned long long w[8]) const {
        uint32_t v[10] = {0};
        unsigned __int128 carry = 0;
        for (int i = 0; i < 8; ++i) {
            carry += w[i];
            v[i] = (uint32_t)carry;
            carry >>= 32;
        }
        v[8] = (uint32_t)carry;
        v[9] = (uint32_t)(carry >> 32);
        Fe p = modulus();
        for (;;) {  // v < world * p: at most world-1 subtractions
            bool ge = (v[8] | v[9]) != 0;
            if (!ge) {
                ge = true;
                for (int i = 7; i >= 0; --i) {
                    if (v[i] > p.l[i]) break;
                    if (v[i] < p.l[i]) { ge = false; break; }
                }
            }
            if (!ge) break;
            uint64_t bw = 0;
            for (int i = 0; i < 10; ++i) {
                uint64_t d = (uint64_t)v[i] - (i < 8 ? p.l[i] : 0u) - bw;
                v[i] = (uint32_t)d;
                bw = (d >> 32) & 1;
            }
        }
        Fe r;
        for (int i = 0; i < 8; ++i) r.l[i] = v[i];
        return r;
    }
};

// Per-launch multiplication table for a fixed multiplicand r (fr.cuh FixedMul):
// t[i] = r * 2^(32 i + 64) * R^-1 mod p = mul(r, 2^(32 i + 64) mod p).
struct FixedMulBuilder {
    Fe c[8];  // 2^(32 i + 64) mod p as plain integers
    void init(const HostField& H) {
        Fe v = H.zero();
        v.l[0] = 1;
        for (int k = 0; k < 64; ++k) v = H.add(v, v);
        for (int i = 0; i < 8; ++i) {
            c[i] = v;
            for (int k = 0; k < 32; ++k) v = H.add(v, v);
        }
    }
    void make(const HostField& H, const Fe& r, FixedMul* out) const {
        for (int i = 0; i < 8; ++i) {
            Fe t = H.mul(r, c[i]);
            std::memcpy(out->t[i], t.l, 32);
        }
    }
};

// The same challenge for the tensor-core fold (fr.cuh TcFoldMats, tcfold.cuh): column k of the two byte matrices is
// T1_k = (1 - r) 2^(8 k + 32) mod p resp. T2_k = r 2^(8 k + 32) mod p; c[32] = ONE (Montgomery form).
struct TcMatsBuilder {
    Fe c[33];  // 2^(8 k + 32) mod p as plain integers, then ONE
    void init(const HostField& H) {
        Fe v = H.zero();
        v.l[1] = 1;
        for (int k = 0; k < 32; ++k) {
            c[k] = v;
            for (int d = 0; d < 8; ++d) v = H.add(v, v);
        }
        c[32] = H.one();
    }
    // one B operand for an arbitrary weight w (Montgomery form): byte n of w 2^(8 k + 32) mod p
    void make_weight(const HostField& H, const Fe& w, uint8_t* out1024) const {
        for (int k = 0; k < 32; ++k) {
            const Fe t = H.mul(w, c[k]);
            uint8_t* m = out1024 + (k / 16) * 512 + k % 16;
            for (int n = 0; n < 32; ++n) m[n * 16] = (uint8_t)(t.l[n / 4] >> (8 * (n % 4)));
        }
    }
    void make(const HostField& H, const Fe& r, TcFoldMats* out) const {
        const Fe omr = H.sub(c[32], r);
        for (int k = 0; k < 32; ++k) {
            const Fe t1 = H.mul(omr, c[k]), t2 = H.mul(r, c[k]);
            uint8_t* m = out->b[0] + (k / 16) * 512 + k % 16;
            for (int n = 0; n < 32; ++n) {
                m[n * 16] = (uint8_t)(t1.l[n / 4] >> (8 * (n % 4)));
                m[1024 + n * 16] = (uint8_t)(t2.l[n / 4] >> (8 * (n % 4)));
            }
        }
    }
};

// -------------------------------------------------------------- Transcript
// fiat_shamir_transcript.rs:5-30
struct TranscriptImpl {
    HostField H;
    Keccak256 hasher;
    void append(const uint8_t* d, size_t n) { hasher.update(d, n); }  // :19-21
    // append(&fq_vec_to_bytes(values)) -- 32-byte LE canonical per element (:32-37)
    void append_elements(const Fe* v, size_t n) {
        for (size_t i = 0; i < n; ++i) {
            Fe c = H.from_mont(v[i]);
            hasher.update(reinterpret_cast<const uint8_t*>(c.l), 32);
        }
    }
    Fe challenge() {  // :23-29
        uint8_t dg[32];
        hasher.finalize_reset(dg);
        hasher.update(dg, 32);  // the new sponge is seeded with the digest (:25)
        return H.from_le_bytes_mod_order(dg);
    }
};

// ---------------------------------------------------------- UnivariatePoly
// interpolate (univariate_polynomial_dense.rs:48-74) for the nodes 0..n-1 that
// get_round_partial_polynomial_proof_gkr uses (sum_check_protocol.rs:157-160),
// with the per-node factors 1/prod_{j!=i}(i-j) precomputed once per ctx.
struct RoundInterpolator {
    HostField H;
    int n = 0;
    std::vector<Fe> basis;  // basis[i*n + k] = coefficient k of the i-th Lagrange basis polynomial

    void init(const HostField& h, int npts) {
        H = h;
        n = npts;
        basis.assign((size_t)n * n, H.zero());
        std::vector<Fe> xs(n);
        for (int i = 0; i < n; ++i) xs[i] = H.from_u64((uint64_t)i);
        for (int i = 0; i < n; ++i) {
            std::vector<Fe> li(1, H.one());
            Fe denom = H.one();
            for (int j = 0; j < n; ++j) {
                if (j == i) continue;
                std::vector<Fe> nx(li.size() + 1, H.zero());
                Fe neg = H.sub(H.zero(), xs[j]);
                for (size_t k = 0; k < li.size(); ++k) {
                    nx[k] = H.add(nx[k], H.mul(li[k], neg));
                    nx[k + 1] = H.add(nx[k + 1], li[k]);
                }
                li.swap(nx);
                denom = H.mul(denom, H.sub(xs[i], xs[j]));
            }
            Fe w = H.inv(denom);
            for (int k = 0; k < n; ++k) basis[(size_t)i * n + k] = H.mul(li[k], w);
        }
    }
    // ys[0..n) -> ascending coefficients; returns the trimmed length (:14-18,:71)
    int interpolate(const Fe* ys, Fe* coeffs) const {
        for (int k = 0; k < n; ++k) coeffs[k] = H.zero();
        for (int i = 0; i < n; ++i)
            for (int k = 0; k < n; ++k) coeffs[k] = H.add(coeffs[k], H.mul(basis[(size_t)i * n + k], ys[i]));
        int len = n;
        while (len > 0 && H.is_zero(coeffs[len - 1])) --len;
        return len;
    }
};

// General-node Lagrange interpolation (the public UnivariatePoly::interpolate).
inline int uni_interpolate(const HostField& H, const Fe* xs, const Fe* ys, int n, Fe* coeffs) {
    std::vector<Fe> acc(n, H.zero());
    for (int i = 0; i < n; ++i) {
        std::vector<Fe> li(1, H.one());
        Fe denom = H.one();
        for (int j = 0; j < n; ++j) {
            if (j == i) continue;
            std::vector<Fe> nx(li.size() + 1, H.zero());
            Fe neg = H.sub(H.zero(), xs[j]);
            for (size_t k = 0; k < li.size(); ++k) {
                nx[k] = H.add(nx[k], H.mul(li[k], neg));
                nx[k + 1] = H.add(nx[k + 1], li[k]);
            }
            li.swap(nx);
            denom = H.mul(denom, H.sub(xs[i], xs[j]));
        }
        Fe s = H.mul(ys[i], H.inv(denom));
        for (int k = 0; k < n; ++k) acc[k] = H.add(acc[k], H.mul(li[k], s));
    }
    int len = n;
    while (len > 0 && H.is_zero(acc[len - 1])) --len;
    for (int k = 0; k < len; ++k) coeffs[k] = acc[k];
    return len;
}
// evaluate (:20-26): sum c_i x^i
inline Fe uni_evaluate(const HostField& H, const Fe* c, int len, const Fe& x) {
    Fe s = H.zero(), xp = H.one();
    for (int i = 0; i < len; ++i) {
        s = H.add(s, H.mul(c[i], xp));
        xp = H.mul(xp, x);
    }
    return s;
}

}  // namespace zkb
