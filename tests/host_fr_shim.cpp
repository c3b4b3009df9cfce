// Host build of zk-research-implementations_b200/csrc/fr.cuh (carry flag emulated) so the
// exact limb algorithms the kernels run can be checked on a machine with no GPU.
#include <cstddef>
#include <cstring>
#include "fr.cuh"
using namespace zkb;

template <class F>
static void run(int op, const uint32_t* a, const uint32_t* b, uint32_t* out, size_t n) {
    for (size_t i = 0; i < n; ++i) {
        Fe x, y, r;
        memcpy(x.l, a + 8 * i, 32);
        memcpy(y.l, b + 8 * i, 32);
        switch (op) {
            case 0: r = Field<F>::add(x, y); break;
            case 1: r = Field<F>::sub(x, y); break;
            case 2: r = Field<F>::mul(x, y); break;
            case 3: r = Field<F>::mul_split(x, y); break;
            case 4: r = Field<F>::to_mont(x); break;
            case 5: r = Field<F>::from_mont(x); break;
            case 6: r = Field<F>::fold(x, y, Field<F>::to_mont(y)); break;
            case 7: {  // x * y * R^-1 through the fixed-multiplicand path (table built like host_math.hpp)
                FixedMul T;
                Fe v = Field<F>::zero();
                v.l[0] = 1;
                for (int k = 0; k < 64; ++k) v = Field<F>::add(v, v);
                for (int j = 0; j < 8; ++j) {
                    Fe t = Field<F>::mul(y, v);
                    memcpy(T.t[j], t.l, 32);
                    for (int k = 0; k < 32; ++k) v = Field<F>::add(v, v);
                }
                r = Field<F>::mul_fixed(x, T);
                break;
            }
            case 8: {  // lazy sum of the products of this and the next 3 pairs (wrapping), reduced once
                Wide w = Field<F>::wide_zero();
                for (size_t k = 0; k < 4; ++k) {
                    Fe p, q;
                    memcpy(p.l, a + 8 * ((i + k) % n), 32);
                    memcpy(q.l, b + 8 * ((i + k) % n), 32);
                    Field<F>::mac_wide(w, p, q);
                }
                r = Field<F>::reduce_wide(w);
                break;
            }
            case 9: {  // the same product accumulated 5000 times (exercises the guard limb)
                Wide w = Field<F>::wide_zero();
                for (int k = 0; k < 5000; ++k) Field<F>::mac_wide(w, x, y);
                r = Field<F>::reduce_wide(w);
                break;
            }
            case 10: r = Field<F>::reduce_once(Field<F>::sub_lazy(x, y)); break;  // x - y + p, then canonical
            case 11: {  // (2y - x + p) * (2y - x + p) lazily, reduced once (only where 3p < 2^256)
                Wide w = Field<F>::wide_zero();
                Fe v = F::SLACK3P ? Field<F>::line2_lazy(x, y) : Field<F>::sub(Field<F>::dbl(y), x);
                Field<F>::mac_wide(w, v, v);
                r = Field<F>::reduce_wide(w);
                break;
            }
            default: r = Field<F>::zero();
        }
        memcpy(out + 8 * i, r.l, 32);
    }
}
extern "C" void host_fr_op(int field, int op, const uint32_t* a, const uint32_t* b, uint32_t* out, size_t n) {
    if (field == 0) run<Bn254Fr>(op, a, b, out, n);
    else if (field == 1) run<Bn254Fq>(op, a, b, out, n);
    else run<Bls12381Fr>(op, a, b, out, n);
}
