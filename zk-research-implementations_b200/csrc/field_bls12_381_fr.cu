#define ZKB_FIELD Bls12381Fr
#define ZKB_FIELD_FN field_kernels_bls12_381_fr
#include "field_impl.cuh"
