/* The C ABI used from plain C, no Python and no torch: echo times -> table -> forward model -> least-squares solve, and the
 * solve must give the water / fat maps back (SURVEY §8c KAT iii: get_rho(IDEAL_model(maps)) == maps[:, :2]).
 *   gcc -std=c99 -I include -I /usr/local/cuda/include tests/c_abi/roundtrip.c -L ideal-gan_b200/idealgan -lidealgan -L /usr/local/cuda/lib64 -lcudart -lm
 * Exit code 0 and "C ABI round trip ok" on success. */
#include <cuda_runtime_api.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "idealgan.h"

#define CHECK(call)                                                              \
    do {                                                                         \
        int rc_ = (call);                                                        \
        if (rc_ != 0) {                                                          \
            fprintf(stderr, "%s -> %d: %s\n", #call, rc_, ig_last_error());      \
            return 1;                                                            \
        }                                                                        \
    } while (0)

int main(void) {
    const int nb = 2, ne = 6, H = 24, W = 32, nv = H * W, rows = 3;
    if (ig_version() != IG_VERSION) { fprintf(stderr, "header / library version mismatch\n"); return 1; }
    if (ig_device_ok() != 1) { fprintf(stderr, "not an sm_100 device\n"); return 2; }
    float *te = malloc(sizeof(float) * nb * ne), *maps = malloc(sizeof(float) * nb * rows * nv * 2), *rho = malloc(sizeof(float) * nb * 2 * nv * 2);
    for (int b = 0; b < nb; ++b)
        for (int e = 0; e < ne; ++e) te[b * ne + e] = 1.3e-3f + 2.1e-3f * e + 1e-4f * b;          /* one echo train per sample */
    unsigned s = 12345u;
    for (int i = 0; i < nb * rows * nv * 2; ++i) {
        s = s * 1664525u + 1013904223u;
        maps[i] = (float)(s >> 8) / 16777216.0f - 0.5f;                                         /* U(-0.5, 0.5) */
    }
    for (int b = 0; b < nb; ++b)                                                                 /* R2* map >= 0: the relu gate is the identity */
        for (int v = 0; v < nv; ++v) maps[((b * rows + 2) * nv + v) * 2 + 1] = fabsf(maps[((b * rows + 2) * nv + v) * 2 + 1]);
    float *te_d, *tab_d, *maps_d, *sig_d, *rho_d;
    if (cudaMalloc((void **)&te_d, sizeof(float) * nb * ne) || cudaMalloc((void **)&tab_d, sizeof(float) * nb * IG_TAB_FLOATS) ||
        cudaMalloc((void **)&maps_d, sizeof(float) * nb * rows * nv * 2) || cudaMalloc((void **)&sig_d, sizeof(float) * nb * ne * nv * 2) ||
        cudaMalloc((void **)&rho_d, sizeof(float) * nb * 2 * nv * 2)) { fprintf(stderr, "cudaMalloc failed\n"); return 3; }
    cudaMemcpy(te_d, te, sizeof(float) * nb * ne, cudaMemcpyHostToDevice);
    cudaMemcpy(maps_d, maps, sizeof(float) * nb * rows * nv * 2, cudaMemcpyHostToDevice);
    CHECK(ig_gen_tables(te_d, nb, ne, 1.5f, tab_d, NULL));
    CHECK(ig_ideal_fwd(IG_MODEL_WFPM, maps_d, rows, tab_d, nb, ne, nv, 200.0f, 0, sig_d, NULL));
    /* the (phi, R2*) row of sample 0 is row 2 of the maps; consecutive samples are rows * nv * 2 floats apart */
    CHECK(ig_get_rho_fwd(sig_d, maps_d + 2 * nv * 2, (long)rows * nv * 2, NULL, 0, tab_d, nb, ne, nv, 200.0f, 0, rho_d, NULL, NULL));
    if (cudaMemcpy(rho, rho_d, sizeof(float) * nb * 2 * nv * 2, cudaMemcpyDeviceToHost) != cudaSuccess) { fprintf(stderr, "copy back failed\n"); return 4; }
    double worst = 0.0, scale = 0.0;
    for (int b = 0; b < nb; ++b)
        for (int r = 0; r < 2; ++r)
            for (int i = 0; i < nv * 2; ++i) {
                const double want = maps[(b * rows + r) * nv * 2 + i], got = rho[(b * 2 + r) * nv * 2 + i];
                if (fabs(got - want) > worst) worst = fabs(got - want);
                if (fabs(want) > scale) scale = fabs(want);
            }
    /* argument validation without a launch */
    if (ig_ideal_fwd(IG_MODEL_WFPM, maps_d, 2, tab_d, nb, ne, nv, 200.0f, 0, sig_d, NULL) != IG_E_ARG) { fprintf(stderr, "rows = 2 was accepted\n"); return 5; }
    if (ig_gen_tables(te_d, nb, IG_MAX_NE + 1, 1.5f, tab_d, NULL) != IG_E_NE) { fprintf(stderr, "ne = 17 was accepted\n"); return 5; }
    printf("round trip error %.3e of %.3f\n", worst, scale);
    if (!(worst <= 1e-5 * scale)) { fprintf(stderr, "round trip off by %.3e\n", worst / scale); return 6; }
    printf("C ABI round trip ok\n");
    return 0;
}
