/*
 * conv_driver.c -- a plain C caller of libhec's C ABI (include/hec.h): what a cgo / JNI / FFI binding does,
 * without Python or torch anywhere.  It reads one evalConv_BN problem (parameters, a level-1 ciphertext,
 * kernel plaintexts, monomials, bias, rotation keys) from a little-endian binary file, runs
 * hec_conv_then_pack (conv.go:522-546 + eval.go:258) on the GPU, and writes the level-0 result.
 *
 *   gcc -O2 -std=c11 -Iinclude examples/conv_driver.c -Loptimal_conv_b200 -lhec \
 *       -Wl,-rpath,$PWD/optimal_conv_b200 -o examples/conv_driver
 *   examples/conv_driver problem.bin result.bin [oplevel]
 *
 * File layout (uint64 words): magic 0x4845435f434f4e56, logN, nQ, nP, Q[nQ], P[nP], B, norm, bits(ct_scale),
 * bits(pt_scale), bits(out_scale), has_bias, nkeys, then c0[2][N], c1[2][N], pt_ker[B][2][N], pt_idx[logN][N],
 * bias[N] (if has_bias), and per key: galEl, swk[beta][2][nQ+nP][N] with beta = ceil(nQ/nP).
 * Result: c0[N], c1[N], bits(scale), level.   tests/test_c_abi_driver.py writes the problem and checks the result.
 */
#include <inttypes.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "hec.h"

#define CHECK(call)                                                                              \
    do {                                                                                         \
        int rc_ = (call);                                                                        \
        if (rc_ != HEC_OK) {                                                                     \
            fprintf(stderr, "%s failed: %d (%s)\n", #call, rc_, ctx ? hec_last_error(ctx) : ""); \
            return 2;                                                                            \
        }                                                                                        \
    } while (0)

static double as_double(uint64_t bits) { double d; memcpy(&d, &bits, 8); return d; }
static uint64_t as_bits(double d) { uint64_t b; memcpy(&b, &d, 8); return b; }

int main(int argc, char **argv) {
    if (argc < 3) { fprintf(stderr, "usage: %s problem.bin result.bin [oplevel]\n", argv[0]); return 1; }
    FILE *f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 1; }
    fseek(f, 0, SEEK_END);
    long bytes = ftell(f);
    fseek(f, 0, SEEK_SET);
    uint64_t *w = malloc((size_t)bytes);
    if (!w || fread(w, 1, (size_t)bytes, f) != (size_t)bytes) { fprintf(stderr, "short read\n"); return 1; }
    fclose(f);
    const uint64_t *p = w;
    if (*p++ != 0x4845435f434f4e56ull) { fprintf(stderr, "bad magic\n"); return 1; }
    int logN = (int)*p++, nQ = (int)*p++, nP = (int)*p++;
    const uint64_t *Q = p; p += nQ;
    const uint64_t *P = p; p += nP;
    int B = (int)*p++, norm = (int)*p++;
    double ct_scale = as_double(*p++), pt_scale = as_double(*p++), out_scale = as_double(*p++);
    int has_bias = (int)*p++, nkeys = (int)*p++;
    size_t N = (size_t)1 << logN;
    int full = nQ + nP, beta = (nQ + nP - 1) / nP;

    hec_ctx *ctx = NULL;
    CHECK(hec_ctx_create(&ctx, logN, Q, nQ, P, nP, 0));

    /* the input ciphertext: two polynomials of two limbs, as ring.Poly.Coeffs[limb] */
    const uint64_t *c0[2] = {p, p + N}, *c1[2] = {p + 2 * N, p + 3 * N};
    p += 4 * N;
    hec_ct *ct = NULL, *res = NULL;
    CHECK(hec_ct_upload(ctx, 1, c0, c1, ct_scale, &ct));

    hec_pt **ker = calloc((size_t)B, sizeof(*ker)), **idx = calloc((size_t)logN, sizeof(*idx)), *bias = NULL;
    for (int i = 0; i < B; i++, p += 2 * N) {
        const uint64_t *limbs[2] = {p, p + N};
        if (i % norm == 0) CHECK(hec_pt_upload(ctx, 1, limbs, pt_scale, &ker[i]));
    }
    for (int i = 0; i < logN; i++, p += N) {
        const uint64_t *limbs[1] = {p};
        CHECK(hec_pt_upload(ctx, 0, limbs, 1.0, &idx[i]));
    }
    if (has_bias) {
        const uint64_t *limbs[1] = {p};
        CHECK(hec_pt_upload(ctx, 0, limbs, out_scale, &bias));
        p += N;
    }
    /* rotation keys: limbs[(d*2 + k)*(nQ+nP) + t] = Value[d][k].Coeffs[t]; only the level-0 slice is kept */
    const uint64_t **kl = malloc((size_t)beta * 2 * full * sizeof(*kl));
    for (int k = 0; k < nkeys; k++) {
        uint64_t galEl = *p++;
        for (int j = 0; j < beta * 2 * full; j++, p += N) kl[j] = p;
        CHECK(hec_swk_upload(ctx, galEl, 0, kl));
    }

    int flags = (argc > 3 && strcmp(argv[3], "oplevel") == 0) ? HEC_CONV_OPLEVEL : HEC_CONV_FUSED;
    CHECK(hec_conv_then_pack(ctx, ct, (const hec_pt *const *)ker, B, norm, out_scale, (const hec_pt *const *)idx, bias, flags, &res));

    uint64_t *o = malloc((2 * N + 2) * 8);
    uint64_t *o0[1] = {o}, *o1[1] = {o + N};
    CHECK(hec_ct_download(ctx, res, o0, o1));
    o[2 * N] = as_bits(hec_ct_scale(res));
    o[2 * N + 1] = (uint64_t)hec_ct_level(res);
    f = fopen(argv[2], "wb");
    if (!f || fwrite(o, 8, 2 * N + 2, f) != 2 * N + 2) { perror(argv[2]); return 1; }
    fclose(f);
    printf("conv_then_pack B=%d norm=%d -> level %d, scale %.17g, %" PRIu64 " kernels launched (%s)\n", B, norm,
           hec_ct_level(res), hec_ct_scale(res), hec_launch_count(ctx), hec_version());

    hec_ct_free(ctx, res); hec_ct_free(ctx, ct);
    for (int i = 0; i < B; i++) if (ker[i]) hec_pt_free(ctx, ker[i]);
    for (int i = 0; i < logN; i++) hec_pt_free(ctx, idx[i]);
    if (bias) hec_pt_free(ctx, bias);
    hec_ctx_destroy(ctx);
    free(kl); free(ker); free(idx); free(o); free(w);
    return 0;
}
