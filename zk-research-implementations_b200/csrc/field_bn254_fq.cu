#define ZKB_FIELD Bn254Fq
#define ZKB_FIELD_FN field_kernels_bn254_fq
#include "field_impl.cuh"
