// Host build of zk-research-implementations_b200/csrc/g1.cuh (carry flag emulated): the BLS12-381 Fq multiplier and
// the G1 group law the KZG kernels run, checked against Python integers on a machine with no GPU.
#include <cstddef>
#include <cstring>
#include "g1.cuh"
using namespace zkb;

extern "C" {
// op: 0 add, 1 sub, 2 mul (Montgomery: a*b*R^-1), 3 to_mont, 4 from_mont, 5 inv (Montgomery: returns a^-1 * R)
void host_fq_op(int op, const uint32_t* a, const uint32_t* b, uint32_t* out, size_t n) {
    for (size_t i = 0; i < n; ++i) {
        Fq x, y, r;
        memcpy(x.l, a + 12 * i, 48);
        memcpy(y.l, b + 12 * i, 48);
        switch (op) {
            case 0: r = Fqf::add(x, y); break;
            case 1: r = Fqf::sub(x, y); break;
            case 2: r = Fqf::mul(x, y); break;
            case 3: r = Fqf::to_mont(x); break;
            case 4: r = Fqf::from_mont(x); break;
            default: r = Fqf::inv(x); break;
        }
        memcpy(out + 12 * i, r.l, 48);
    }
}
// Points cross as canonical affine (x, y), 24 words; (0, 0) = infinity.
static G1Affine load(const uint32_t* p) {
    G1Affine a;
    memcpy(a.x.l, p, 48);
    memcpy(a.y.l, p + 12, 48);
    if (!G1::is_inf(a)) {
        a.x = Fqf::to_mont(a.x);
        a.y = Fqf::to_mont(a.y);
    }
    return a;
}
static void store(const G1Jac& j, uint32_t* p) {
    G1Affine a = G1::to_affine(j);
    if (!G1::is_inf(a)) {
        a.x = Fqf::from_mont(a.x);
        a.y = Fqf::from_mont(a.y);
    }
    memcpy(p, a.x.l, 48);
    memcpy(p + 12, a.y.l, 48);
}
// op: 0 = P + Q (Jacobian add), 1 = P + Q (mixed add), 2 = 2P, 3 = k*P with k = b[0] (mul_small), 4 = generator,
//     5 = ((P + Q) + Q) + P through Jacobian intermediates with non-trivial z
void host_g1_op(int op, const uint32_t* a, const uint32_t* b, uint32_t* out, size_t n) {
    for (size_t i = 0; i < n; ++i) {
        const G1Affine P = load(a + 24 * i), Q = load(b + 24 * i);
        G1Jac r;
        switch (op) {
            case 0: r = G1::add(G1::from_affine(P), G1::from_affine(Q)); break;
            case 1: r = G1::madd(G1::from_affine(P), Q); break;
            case 2: r = G1::dbl(G1::from_affine(P)); break;
            case 3: r = G1::mul_small(G1::from_affine(P), b[24 * i]); break;
            case 4: r = G1::from_affine(G1::generator()); break;
            default: {
                G1Jac s = G1::madd(G1::from_affine(P), Q);
                s = G1::madd(s, Q);
                r = G1::add(s, G1::dbl(G1::from_affine(P)));
                r = G1::add(r, G1::mul_small(G1::from_affine(P), 0xffffffffu));  // + (2^32 - 1) P  => (2^32 + 2) P + 2 Q
                break;
            }
        }
        store(r, out + 24 * i);
    }
}
}
