/* abi_smoke.c -- the header of the drop-in boundary compiled as plain C and linked against libzkb200.so: catches type
 * drift between include/zkb200.h and the library (a ctypes or regex check only sees names).  Host-only calls; on a
 * machine without a GPU zkb_ctx_create must fail with ZKB_ERR_CUDA (there is no CPU fallback).
 *   gcc -std=c11 -Wall -Wextra -Werror -I include tests/abi_smoke.c -L <pkg> -lzkb200 -Wl,-rpath,<pkg> -o abi_smoke */
#include <stdio.h>
#include <string.h>

#include "zkb200.h"

/* every entry point with its exact prototype: assigning to a typed function pointer fails to compile on any drift */
static int32_t (*const p_ctx_create)(int32_t, int32_t, int32_t, zkb_ctx**) = zkb_ctx_create;
static int32_t (*const p_mle_upload)(zkb_ctx*, const uint64_t*, uint64_t, zkb_mle*) = zkb_mle_upload;
static int32_t (*const p_mle_pe)(zkb_ctx*, zkb_mle, uint32_t, const uint64_t[4], zkb_mle*) = zkb_mle_partial_evaluate;
static int32_t (*const p_mle_eval)(zkb_ctx*, zkb_mle, const uint64_t*, uint32_t, uint64_t[4]) = zkb_mle_evaluate;
static int32_t (*const p_sp_create)(zkb_ctx*, const zkb_mle*, uint32_t, uint32_t, zkb_sp*) = zkb_sumpoly_create;
static int32_t (*const p_sc_prove)(zkb_ctx*, zkb_mle, uint32_t, uint64_t[4], uint64_t*, uint64_t*) = zkb_sumcheck_prove;
static int32_t (*const p_sc_verify)(zkb_ctx*, zkb_mle, uint32_t, const uint64_t[4], const uint64_t*, uint32_t, int32_t*) = zkb_sumcheck_verify;
static int32_t (*const p_gkr_sc_prove)(zkb_ctx*, zkb_transcript*, const uint64_t[4], zkb_sp, uint64_t*, int32_t*, uint64_t*, uint64_t*) = zkb_gkr_sumcheck_prove;
static int32_t (*const p_gkr_sc_verify)(zkb_transcript*, uint32_t, uint32_t, const uint64_t*, const int32_t*, const uint64_t[4], int32_t*,
                                        uint64_t[4], uint64_t*) = zkb_gkr_sumcheck_verify;
static int32_t (*const p_circ_create)(zkb_ctx*, uint32_t, const uint32_t*, const uint8_t*, zkb_circ*) = zkb_circuit_create;

static int fails = 0;
#define EXPECT(c)                                                        \
    do {                                                                 \
        if (!(c)) {                                                      \
            fprintf(stderr, "abi_smoke: %s:%d: %s\n", __FILE__, __LINE__, #c); \
            ++fails;                                                     \
        }                                                                \
    } while (0)

int main(void) {
    (void)p_ctx_create; (void)p_mle_upload; (void)p_mle_pe; (void)p_mle_eval; (void)p_sp_create; (void)p_sc_prove;
    (void)p_sc_verify; (void)p_gkr_sc_prove; (void)p_gkr_sc_verify; (void)p_circ_create;
    EXPECT(strstr(zkb_version(), "sm_100a") != NULL);
    EXPECT(strcmp(zkb_strerror(ZKB_ERR_NOT_POW2), "Invalid evaluations") == 0);
    EXPECT(strcmp(zkb_strerror(ZKB_ERR_ARITY), "Invalid number of values") == 0);
    /* Keccak-256 (original padding) known answer: keccak256("") */
    uint8_t dg[32];
    EXPECT(zkb_keccak256((const uint8_t*)"", 0, dg) == ZKB_OK);
    EXPECT(dg[0] == 0xc5 && dg[1] == 0xd2 && dg[31] == 0x70);
    /* host transcript + field conversion round trip */
    zkb_transcript* t = NULL;
    EXPECT(zkb_transcript_new(ZKB_FIELD_BN254_FR, &t) == ZKB_OK && t != NULL);
    EXPECT(zkb_transcript_append(t, (const uint8_t*)"zero knowledge", 14) == ZKB_OK);
    uint64_t r[4], canon[4], back[4];
    EXPECT(zkb_transcript_challenge(t, r) == ZKB_OK);
    EXPECT(zkb_fe_from_mont(ZKB_FIELD_BN254_FR, r, canon, 1) == ZKB_OK);
    EXPECT(zkb_fe_to_mont(ZKB_FIELD_BN254_FR, canon, back, 1) == ZKB_OK);
    EXPECT(memcmp(r, back, 32) == 0);
    /* first challenge after append("zero knowledge") over BN254 Fr */
    EXPECT(canon[3] == 0x020d8026e5dccbcaull && canon[0] == 0x04566af81a460761ull);
    EXPECT(zkb_transcript_free(t) == ZKB_OK);
    /* interpolate {(0,2),(1,4),(2,6)} -> [2,2] (univariate_polynomial_dense.rs tests): trimmed to two coefficients */
    uint64_t xs[12] = {0}, ys[12] = {0}, xm[12], ym[12], co[12];
    uint32_t len = 0;
    xs[4] = 1; xs[8] = 2; ys[0] = 2; ys[4] = 4; ys[8] = 6;
    EXPECT(zkb_fe_to_mont(ZKB_FIELD_BN254_FQ, xs, xm, 3) == ZKB_OK && zkb_fe_to_mont(ZKB_FIELD_BN254_FQ, ys, ym, 3) == ZKB_OK);
    EXPECT(zkb_uni_interpolate(ZKB_FIELD_BN254_FQ, xm, ym, 3, co, &len) == ZKB_OK && len == 2);
    EXPECT(zkb_fe_from_mont(ZKB_FIELD_BN254_FQ, co, xs, 2) == ZKB_OK && xs[0] == 2 && xs[4] == 2);
    /* no CPU fallback: without a device the context cannot be created */
    zkb_ctx* ctx = NULL;
    int32_t st = zkb_ctx_create(ZKB_FIELD_BN254_FR, 0, ZKB_MODE_FULL, &ctx);
    if (st == ZKB_OK) {
        EXPECT(ctx != NULL);
        EXPECT(zkb_ctx_destroy(ctx) == ZKB_OK);
        printf("abi_smoke: device present\n");
    } else {
        EXPECT(st == ZKB_ERR_CUDA);
        printf("abi_smoke: no device (ZKB_ERR_CUDA), as expected on a CPU box\n");
    }
    printf("abi_smoke: %s\n", fails ? "FAILED" : "ok");
    return fails ? 1 : 0;
}
