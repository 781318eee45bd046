// hec_encode.cu -- Encoder.EncodeCoeffs + Encoder.ToNTT on the device: the float coefficient vectors that
// prep_Ker (conv.go:487-516) and evalConv_BN (eval.go:238-243) build per layer become evaluator plaintexts
// without a host-side RNS reduction or transform.  Part of libhec.so (included by libhec.cu).
//
// Reference behaviour followed (pinned through tests/golden/ref_eval_vectors.json "encode_coeffs"):
//   L:ckks/encoder.go EncodeCoeffs  -> scaleUpVecExact(values, pt.Scale, Q[:level+1], pt.Coeffs); IsNTT = false
//   L:ckks/utils.go   scaleUpVecExact: per value v, with n = scale
//       n*|v| >  2^64 : X = trunc(double(n*|v|) + 0.5) as an exact integer (big.Float at 53 bits), r = X mod q
//       otherwise     : X = uint64(n*|v| + 0.5) with Go's amd64 float->uint64 conversion,         r = X mod q
//       v < 0 stores q - r, v >= 0 (also -0.0) stores r; coefficients past len(values) are cleared
//   L:ckks/encoder.go ToNTT         -> ring.NTTLvl(level, pt, pt)
// A negative value that rounds to zero leaves the word q (not 0) in the reference's coefficient-domain
// plaintext; after ToNTT that is not observable (the transform's outputs are canonical), so the kernel writes 0.

#define HEC_ENC_MAXPT 16

struct EncJobs {
    u64 *out[HEC_ENC_MAXPT]; // plaintext buffers [level+1][N]
    u64 r2[32];              // R^2 mod q_j (to Montgomery form before the transform; the NTT is linear)
};

// Go's float64 -> uint64 as compiled for amd64: CVTTSD2SI below 2^63, else CVTTSD2SI(x - 2^63) | 2^63, where an
// out-of-range CVTTSD2SI returns 0x8000000000000000
__device__ __forceinline__ u64 enc_cvttsd2si(double y) {
    const double two63 = 9223372036854775808.0;
    if (!(y < two63) || y < -two63) return 1ull << 63;
    return (u64)(long long)y;
}
__device__ __forceinline__ u64 enc_go_f2u(double x) {
    const double two63 = 9223372036854775808.0;
    return x < two63 ? enc_cvttsd2si(x) : (enc_cvttsd2si(__dsub_rn(x, two63)) | (1ull << 63));
}

// grid = (N / HEC_THREADS, plaintexts in this launch); one thread per coefficient, all limbs
__global__ void __launch_bounds__(HEC_THREADS) k_encode_coeffs(EncJobs J, const double *__restrict__ vals, int n_values,
                                                                int level, double scale, const ModC *__restrict__ mods,
                                                                int *__restrict__ bad) {
    const int i = blockIdx.x * HEC_THREADS + threadIdx.x;
    u64 *out = J.out[blockIdx.y];
    if (i >= n_values) {
        for (int j = 0; j <= level; j++) out[(size_t)j * HEC_N + i] = 0;
        return;
    }
    const double v = vals[(size_t)blockIdx.y * n_values + i];
    if (!isfinite(v)) atomicOr(bad, 1); // reported by the host as HEC_E_INVAL (big.Float panics on NaN)
    const bool neg = v < 0.0;
    const double ax = __dmul_rn(scale, fabs(v)); // == (-n)*v resp. n*v: IEEE products are sign-symmetric
    const double y = __dadd_rn(ax, 0.5);
    const bool big = ax > 18446744073709551616.0;
    u64 x = 0;
    int e = 0;
    if (big) { // y >= 2^64 is an integer: mantissa * 2^e with e >= 12
        const u64 bits = (u64)__double_as_longlong(y);
        x = (bits & ((1ull << 52) - 1)) | (1ull << 52);
        e = (int)((bits >> 52) & 0x7ff) - 1075;
    } else {
        x = enc_go_f2u(y);
    }
    for (int j = 0; j <= level; j++) {
        const ModC M = mods[j];
        u64 r = mred(x, J.r2[j], M.q, M.qinv); // x * R mod q, canonical, for any x < 2^64: no division
        for (int k = 0; k < e; k++) { // rare path (doubling commutes with the Montgomery factor); q < 2^63
            r <<= 1;
            if (r >= M.q) r -= M.q;
        }
        if (neg && r) r = M.q - r;
        out[(size_t)j * HEC_N + i] = r;
    }
}

static int encode_many(hec_ctx *c, const double *values, int count, int n_values, int level, double scale, hec_pt **out) {
    if (!c || !values || !out || count < 1 || level < 0 || level >= c->nQ || level >= 32)
        return c ? c->fail(HEC_E_INVAL, "encode_coeffs args") : HEC_E_INVAL;
    if (n_values < 0 || n_values > HEC_N) return c->fail(HEC_E_INVAL, "encode_coeffs: too many values for the ring degree"); // the reference panics
    if (!std::isfinite(scale)) return c->fail(HEC_E_INVAL, "encode_coeffs: scale not finite");
    const size_t total = (size_t)count * n_values, vlimbs = (total + HEC_N - 1) / HEC_N;
    cudaSetDevice(c->device);
    int rc = reserve(c, vlimbs + 1);
    if (rc) return rc;
    double *dv = reinterpret_cast<double *>(c->scratch(vlimbs));
    int *dbad = reinterpret_cast<int *>(c->scratch(1)), bad = 0;
    HEC_CUDA(c, cudaMemsetAsync(dbad, 0, sizeof(int), c->stream));
    if (total) HEC_CUDA(c, cudaMemcpyAsync(dv, values, total * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    std::vector<hec_pt *> pts;
    auto drop = [&]() { for (hec_pt *p : pts) { cudaFreeAsync(p->buf, c->stream); delete p; } };
    for (int p = 0; p < count; p++) {
        hec_pt *pt = new hec_pt();
        pt->level = level; pt->scale = scale; pt->serial = c->next_serial++;
        if (cudaMallocFromPoolAsync(&pt->buf, (size_t)(level + 1) * HEC_N * sizeof(u64), c->pool, c->stream) != cudaSuccess) { delete pt; drop(); return c->fail(HEC_E_NOMEM, "cudaMallocAsync plaintext"); }
        pts.push_back(pt);
    }
    EncJobs J;
    for (int j = 0; j <= level; j++) J.r2[j] = mform(c->hm[j].rmod, c->q(j));
    std::vector<LimbJob> nj;
    for (int p0 = 0; p0 < count; p0 += HEC_ENC_MAXPT) {
        int n = std::min(HEC_ENC_MAXPT, count - p0);
        for (int p = 0; p < n; p++) {
            J.out[p] = pts[p0 + p]->buf;
            for (int j = 0; j <= level; j++) nj.push_back({pts[p0 + p]->buf + (size_t)j * HEC_N, pts[p0 + p]->buf + (size_t)j * HEC_N, c->modQ(j), 0});
        }
        k_encode_coeffs<<<dim3(HEC_N / HEC_THREADS, n), HEC_THREADS, 0, c->stream>>>(J, dv + (size_t)p0 * n_values, n_values, level, scale, c->dmods, dbad);
        c->launches += 1;
    }
    if ((rc = check_launch(c, "encode_coeffs")) || (rc = hec_launch_ntt(c, nj, false))) { drop(); return rc; }
    cudaError_t e = cudaMemcpyAsync(&bad, dbad, sizeof(int), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream); // the caller may reuse `values` after the return
    if (e != cudaSuccess) { drop(); return c->fail(HEC_E_CUDA, std::string("encode_coeffs: ") + cudaGetErrorString(e)); }
    if (bad) { drop(); return c->fail(HEC_E_INVAL, "encode_coeffs: value not finite"); }
    for (int p = 0; p < count; p++) out[p] = pts[p];
    return HEC_OK;
}

extern "C" int hec_encode_coeffs(hec_ctx *c, const double *values, int n_values, int level, double scale, hec_pt **out) {
    return encode_many(c, values, 1, n_values, level, scale, out);
}
extern "C" int hec_encode_coeffs_many(hec_ctx *c, const double *values, int count, int n_values, int level, double scale, hec_pt **out) {
    return encode_many(c, values, count, n_values, level, scale, out);
}

// the plaintext's limbs as canonical residues in the NTT domain (what Plaintext.Value.Coeffs holds after ToNTT), e.g.
// to let the caller cache the encoded kernels of a layer across images
extern "C" int hec_pt_download(hec_ctx *c, const hec_pt *pt, uint64_t *const *limbs) {
    if (!c || !pt || !limbs) return c ? c->fail(HEC_E_INVAL, "pt_download args") : HEC_E_INVAL;
    cudaSetDevice(c->device);
    int L = pt->level + 1, rc = reserve(c, L);
    if (rc) return rc;
    u64 *t = c->scratch(L);
    std::vector<EwJob> jobs;
    for (int j = 0; j < L; j++) jobs.push_back(ewjob(pt->buf + (size_t)j * HEC_N, nullptr, t + (size_t)j * HEC_N, c->modQ(j), 1)); // * 1 * R^-1
    if ((rc = launch_ew<EW_MULSCALAR>(c, jobs))) return rc;
    for (int j = 0; j < L; j++) HEC_CUDA(c, cudaMemcpyAsync(limbs[j], t + (size_t)j * HEC_N, HEC_N * sizeof(u64), cudaMemcpyDeviceToHost, c->stream));
    HEC_CUDA(c, cudaStreamSynchronize(c->stream));
    return HEC_OK;
}
