// g1.cuh -- BLS12-381 base-field (Fq, 381 bits, 12 x 32-bit limbs, Montgomery R = 2^384) and G1 arithmetic for the
// input-layer commitment of the GKR prover: the multilinear KZG of pcs/src/kzg_pcs/kzg.rs:17-95, which
// gkr/src/gkr_protocol.rs:92-118 runs on the input MLE after the last layer sumcheck.
//
// Replaces the ark-bls12-381 0.5.0 / ark-ec 0.5.0 operators the reference calls (`G1Projective` add, `mul_bigint`,
// `sum`): short Weierstrass y^2 = x^3 + 4, Jacobian coordinates, the standard generator.  The multiplier is the same
// even/odd-column CIOS row as fr.cuh (every 32x32->64 product lands on a 64-bit aligned limb pair, one
// IMAD.WIDE.U32(.X) each), generalised to N limbs.  Compiles for the host as well (carry flag emulated, fr.cuh), which
// is how the limb code is tested without a GPU (tests/test_g1_host.py).
#pragma once
#include "fr.cuh"

// The 12-limb product is ~600 instructions: as a real function (not inlined at its ~16 call sites per point addition)
// it costs a few per cent at run time and cuts the compile time of the MSM kernels by an order of magnitude.
#if defined(__CUDACC__)
#define ZK_HD_CALL __host__ __device__ __noinline__
#else
#define ZK_HD_CALL inline
#endif

namespace zkb {

struct Fq {
    uint32_t l[12];
};

struct Bls12381Fq {
    static constexpr int N = 12;
    static constexpr uint32_t INV = 0xfffcfffdu;  // -p^-1 mod 2^32
    ZK_HD static constexpr uint32_t P(int i) {
        constexpr uint32_t v[12] = {0xffffaaabu, 0xb9feffffu, 0xb153ffffu, 0x1eabfffeu, 0xf6b0f624u, 0x6730d2a0u,
                                    0xf38512bfu, 0x64774b84u, 0x434bacd7u, 0x4b1ba7b6u, 0x397fe69au, 0x1a0111eau};
        return v[i];
    }
    // R mod p and R^2 mod p (R = 2^384), filled in by tools/gen_fq_constants.py (checked in tests/test_g1_host.py)
    ZK_HD static constexpr uint32_t ONE(int i) {
        constexpr uint32_t v[12] = {0x0002fffdu, 0x76090000u, 0xc40c0002u, 0xebf4000bu, 0x53c758bau, 0x5f489857u,
                                    0x70525745u, 0x77ce5853u, 0xa256ec6du, 0x5c071a97u, 0xfa80e493u, 0x15f65ec3u};
        return v[i];
    }
    ZK_HD static constexpr uint32_t R2(int i) {
        constexpr uint32_t v[12] = {0x1c341746u, 0xf4df1f34u, 0x09d104f1u, 0x0a76e6a6u, 0x4c95b6d5u, 0x8de5476cu,
                                    0x939d83c0u, 0x67eb88a9u, 0xb519952du, 0x9a793e85u, 0x92cae3aau, 0x11988fe5u};
        return v[i];
    }
};

template <class P_>
struct BigField {
    static constexpr int N = P_::N;
    static constexpr int H = N / 2;
    typedef Fq E;
    ZK_HD static E zero() {
        E r;
#pragma unroll
        for (int i = 0; i < N; ++i) r.l[i] = 0;
        return r;
    }
    ZK_HD static E one() {
        E r;
#pragma unroll
        for (int i = 0; i < N; ++i) r.l[i] = P_::ONE(i);
        return r;
    }
    ZK_HD static E r2() {
        E r;
#pragma unroll
        for (int i = 0; i < N; ++i) r.l[i] = P_::R2(i);
        return r;
    }
    ZK_HD static bool is_zero(const E& a) {
        uint32_t o = 0;
#pragma unroll
        for (int i = 0; i < N; ++i) o |= a.l[i];
        return o == 0;
    }
    ZK_HD static bool eq(const E& a, const E& b) {
        uint32_t o = 0;
#pragma unroll
        for (int i = 0; i < N; ++i) o |= a.l[i] ^ b.l[i];
        return o == 0;
    }
    ZK_HD static E reduce_once(const E& a) {  // a < 2p
        E t;
        t.l[0] = sub_cc(a.l[0], P_::P(0));
#pragma unroll
        for (int i = 1; i < N; ++i) t.l[i] = subc_cc(a.l[i], P_::P(i));
        const uint32_t borrow = subc(0u, 0u);
        E r;
#pragma unroll
        for (int i = 0; i < N; ++i) r.l[i] = borrow ? a.l[i] : t.l[i];
        return r;
    }
    ZK_HD static E add(const E& a, const E& b) {  // 2p < 2^384
        E s;
        s.l[0] = add_cc(a.l[0], b.l[0]);
#pragma unroll
        for (int i = 1; i < N - 1; ++i) s.l[i] = addc_cc(a.l[i], b.l[i]);
        s.l[N - 1] = addc(a.l[N - 1], b.l[N - 1]);
        return reduce_once(s);
    }
    ZK_HD static E sub(const E& a, const E& b) {
        E d;
        d.l[0] = sub_cc(a.l[0], b.l[0]);
#pragma unroll
        for (int i = 1; i < N; ++i) d.l[i] = subc_cc(a.l[i], b.l[i]);
        const uint32_t m = subc(0u, 0u);
        E r;
        r.l[0] = add_cc(d.l[0], P_::P(0) & m);
#pragma unroll
        for (int i = 1; i < N - 1; ++i) r.l[i] = addc_cc(d.l[i], P_::P(i) & m);
        r.l[N - 1] = addc(d.l[N - 1], P_::P(N - 1) & m);
        return r;
    }
    ZK_HD static E neg(const E& a) { return sub(zero(), a); }
    ZK_HD static E dbl(const E& a) { return add(a, a); }

    // one CIOS row on H x u64 accumulator pairs (see fr.cuh roww for the invariants)
    ZK_HD static void roww(uint64_t* ev, uint64_t* od, const uint32_t* a, uint32_t bi, bool first) {
        if (first) {
#pragma unroll
            for (int k = 0; k < H; ++k) {
                od[k] = mul_wide(a[2 * k + 1], bi);
                ev[k] = mul_wide(a[2 * k], bi);
            }
        } else {
            const uint32_t e0 = add_cc(lo32(ev[0]), hi32(od[0]));
#pragma unroll
            for (int k = 0; k < H - 1; ++k) od[k] = madwc_cc(a[2 * k + 1], bi, od[k + 1]);
            od[H - 1] = madwc(a[N - 1], bi, 0ull);
            ev[0] = pack64(e0, hi32(ev[0]));
            ev[0] = madw_cc(a[0], bi, ev[0]);
#pragma unroll
            for (int k = 1; k < H; ++k) ev[k] = madwc_cc(a[2 * k], bi, ev[k]);
            od[H - 1] = pack64(lo32(od[H - 1]), addc(hi32(od[H - 1]), 0u));
        }
        const uint32_t m = mul_lo(lo32(ev[0]), P_::INV);
        od[0] = madw_cc(P_::P(1), m, od[0]);
#pragma unroll
        for (int k = 1; k < H; ++k) od[k] = madwc_cc(P_::P(2 * k + 1), m, od[k]);
        ev[0] = madw_cc(P_::P(0), m, ev[0]);
#pragma unroll
        for (int k = 1; k < H; ++k) ev[k] = madwc_cc(P_::P(2 * k), m, ev[k]);
        od[H - 1] = pack64(lo32(od[H - 1]), addc(hi32(od[H - 1]), 0u));
    }
    ZK_HD_CALL static E mul(const E& a, const E& b) {
        uint64_t ev[H], od[H];
#pragma unroll
        for (int i = 0; i < N; i += 2) {
            roww(ev, od, a.l, b.l[i], i == 0);
            roww(od, ev, a.l, b.l[i + 1], false);
        }
        E r;  // ev + (od >> 32); od's limb 0 is zero
        r.l[0] = add_cc(lo32(ev[0]), hi32(od[0]));
#pragma unroll
        for (int k = 1; k < N - 1; ++k) {
            const uint32_t e = (k & 1) ? hi32(ev[k >> 1]) : lo32(ev[k >> 1]);
            const uint32_t o = (k & 1) ? lo32(od[(k + 1) >> 1]) : hi32(od[k >> 1]);
            r.l[k] = addc_cc(e, o);
        }
        r.l[N - 1] = addc(hi32(ev[H - 1]), 0u);
        return reduce_once(r);
    }
    ZK_HD static E sqr(const E& a) { return mul(a, a); }
    ZK_HD static E to_mont(const E& c) { return mul(c, r2()); }
    ZK_HD static E from_mont(const E& m) {
        E o = zero();
        o.l[0] = 1;
        return mul(m, o);
    }
    // a^(p-2): the inverse (a != 0)
    ZK_HD static E inv(const E& a) {
        E r = one();
        for (int i = N - 1; i >= 0; --i) {
            uint32_t w = P_::P(i);
            if (i == 0) w -= 2;  // p - 2: the low limb of p is 0x...aaab, no borrow
            for (int b = 31; b >= 0; --b) {
                r = sqr(r);
                if ((w >> b) & 1u) r = mul(r, a);
            }
        }
        return r;
    }
};
typedef BigField<Bls12381Fq> Fqf;

// ---------------------------------------------------------------- G1 points
// Affine points are stored as Montgomery (x, y); (0, 0) encodes the point at infinity (not on y^2 = x^3 + 4).
struct G1Affine {
    Fq x, y;
};
struct G1Jac {
    Fq x, y, z;  // z == 0: infinity
};

struct G1 {
    ZK_HD static G1Jac inf() {
        G1Jac r;
        r.x = Fqf::one();
        r.y = Fqf::one();
        r.z = Fqf::zero();
        return r;
    }
    ZK_HD static bool is_inf(const G1Jac& p) { return Fqf::is_zero(p.z); }
    ZK_HD static bool is_inf(const G1Affine& p) { return Fqf::is_zero(p.x) && Fqf::is_zero(p.y); }
    ZK_HD static G1Jac from_affine(const G1Affine& p) {
        if (is_inf(p)) return inf();
        G1Jac r;
        r.x = p.x;
        r.y = p.y;
        r.z = Fqf::one();
        return r;
    }
    // dbl-2009-l (a = 0): 2M + 5S
    ZK_HD static G1Jac dbl(const G1Jac& p) {
        if (is_inf(p)) return p;
        const Fq A = Fqf::sqr(p.x), B = Fqf::sqr(p.y), C = Fqf::sqr(B);
        Fq D = Fqf::sub(Fqf::sub(Fqf::sqr(Fqf::add(p.x, B)), A), C);
        D = Fqf::dbl(D);
        const Fq E = Fqf::add(Fqf::dbl(A), A), F = Fqf::sqr(E);
        G1Jac r;
        r.x = Fqf::sub(F, Fqf::dbl(D));
        Fq C8 = Fqf::dbl(Fqf::dbl(Fqf::dbl(C)));
        r.y = Fqf::sub(Fqf::mul(E, Fqf::sub(D, r.x)), C8);
        r.z = Fqf::dbl(Fqf::mul(p.y, p.z));
        return r;
    }
    // madd-2007-bl: Jacobian + affine, 7M + 4S; handles infinity, doubling and inverse points
    ZK_HD static G1Jac madd(const G1Jac& p, const G1Affine& q) {
        if (is_inf(q)) return p;
        if (is_inf(p)) return from_affine(q);
        const Fq Z1Z1 = Fqf::sqr(p.z);
        const Fq U2 = Fqf::mul(q.x, Z1Z1);
        const Fq S2 = Fqf::mul(Fqf::mul(q.y, p.z), Z1Z1);
        const Fq Hh = Fqf::sub(U2, p.x);
        Fq rr = Fqf::sub(S2, p.y);
        if (Fqf::is_zero(Hh)) {
            if (Fqf::is_zero(rr)) return dbl(p);
            return inf();
        }
        rr = Fqf::dbl(rr);
        const Fq HH = Fqf::sqr(Hh);
        const Fq I = Fqf::dbl(Fqf::dbl(HH));
        const Fq J = Fqf::mul(Hh, I);
        const Fq V = Fqf::mul(p.x, I);
        G1Jac r;
        r.x = Fqf::sub(Fqf::sub(Fqf::sqr(rr), J), Fqf::dbl(V));
        r.y = Fqf::sub(Fqf::mul(rr, Fqf::sub(V, r.x)), Fqf::dbl(Fqf::mul(p.y, J)));
        r.z = Fqf::sub(Fqf::sub(Fqf::sqr(Fqf::add(p.z, Hh)), Z1Z1), HH);
        return r;
    }
    // add-2007-bl: Jacobian + Jacobian, 11M + 5S
    ZK_HD static G1Jac add(const G1Jac& p, const G1Jac& q) {
        if (is_inf(p)) return q;
        if (is_inf(q)) return p;
        const Fq Z1Z1 = Fqf::sqr(p.z), Z2Z2 = Fqf::sqr(q.z);
        const Fq U1 = Fqf::mul(p.x, Z2Z2), U2 = Fqf::mul(q.x, Z1Z1);
        const Fq S1 = Fqf::mul(Fqf::mul(p.y, q.z), Z2Z2), S2 = Fqf::mul(Fqf::mul(q.y, p.z), Z1Z1);
        const Fq Hh = Fqf::sub(U2, U1);
        Fq rr = Fqf::sub(S2, S1);
        if (Fqf::is_zero(Hh)) {
            if (Fqf::is_zero(rr)) return dbl(p);
            return inf();
        }
        rr = Fqf::dbl(rr);
        const Fq I = Fqf::sqr(Fqf::dbl(Hh));
        const Fq J = Fqf::mul(Hh, I);
        const Fq V = Fqf::mul(U1, I);
        G1Jac r;
        r.x = Fqf::sub(Fqf::sub(Fqf::sqr(rr), J), Fqf::dbl(V));
        r.y = Fqf::sub(Fqf::mul(rr, Fqf::sub(V, r.x)), Fqf::dbl(Fqf::mul(S1, J)));
        r.z = Fqf::mul(Fqf::sub(Fqf::sub(Fqf::sqr(Fqf::add(p.z, q.z)), Z1Z1), Z2Z2), Hh);
        return r;
    }
    ZK_HD static G1Affine to_affine(const G1Jac& p) {
        G1Affine r;
        if (is_inf(p)) {
            r.x = Fqf::zero();
            r.y = Fqf::zero();
            return r;
        }
        const Fq zi = Fqf::inv(p.z), zi2 = Fqf::sqr(zi);
        r.x = Fqf::mul(p.x, zi2);
        r.y = Fqf::mul(p.y, Fqf::mul(zi2, zi));
        return r;
    }
    // k * p for a small unsigned k (double-and-add from the top bit)
    ZK_HD static G1Jac mul_small(const G1Jac& p, uint32_t k) {
        G1Jac acc = inf();
        for (int b = 31; b >= 0; --b) {
            acc = dbl(acc);
            if ((k >> b) & 1u) acc = add(acc, p);
        }
        return acc;
    }
    // the standard generator, Montgomery form (tools/gen_fq_constants.py)
    ZK_HD static G1Affine generator() {
        G1Affine g;
        constexpr uint32_t gx[12] = {0xfd530c16u, 0x5cb38790u, 0x9976fff5u, 0x7817fc67u, 0x143ba1c1u, 0x154f95c7u,
                                     0xf3d0e747u, 0xf0ae6acdu, 0x21dbf440u, 0xedce6eccu, 0x9e0bfb75u, 0x12017741u};
        constexpr uint32_t gy[12] = {0x0ce72271u, 0xbaac93d5u, 0x7918fd8eu, 0x8c22631au, 0x570725ceu, 0xdd595f13u,
                                     0x50405194u, 0x51ac5829u, 0xad0059c0u, 0x0e1c8c3fu, 0x5008a26au, 0x0bbc3efcu};
#pragma unroll
        for (int i = 0; i < 12; ++i) {
            g.x.l[i] = gx[i];
            g.y.l[i] = gy[i];
        }
        return g;
    }
};

}  // namespace zkb
