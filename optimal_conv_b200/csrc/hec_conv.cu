// hec_conv.cu -- the conv_then_pack path (conv.go:522-546 incl. pack_ctxts conv.go:266-300,
// + the bias Add of evalConv_BN eval.go:258):
//   * op-level replay through the evaluator ops (reads like the Go code), and
//   * the fused plan: 3 kernels for Stage A + 5 per pack-tree level, captured in a CUDA graph.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <functional>
#include "hec_host.cuh"

using namespace hec;

int hec_ct_alloc(hec_ctx *c, int level, double scale, hec_ct **out);
double hec_const_limbs(const hec_ctx *c, int level, double constant, std::vector<u64> &k);

// =========================================================================================
// op-level replay (host side mirrors the reference line by line)
// =========================================================================================
// pack_ctxts (conv.go:266-300)
static int pack_ctxts(hec_ctx *ev, std::vector<hec_ct *> &ctxts_in, int max_cnum, int real_cnum,
                      const hec_pt *const *idx, hec_ct **result) {
    int step = max_cnum / 2;
    int norm = max_cnum / real_cnum;
    int rc;
    std::vector<hec_ct *> ctxts(max_cnum, nullptr);
    hec_ct *tmp1 = nullptr, *tmp2 = nullptr;
    auto fail = [&](int code) { // every early return frees what this call allocated
        for (auto &p : ctxts) if (p) { hec_ct_free(ev, p); p = nullptr; }
        if (tmp1) hec_ct_free(ev, tmp1);
        if (tmp2) hec_ct_free(ev, tmp2);
        return code;
    };
    for (int i = 0; i < max_cnum; i++)
        if (i % norm == 0) {
            if ((rc = hec_ct_copy_new(ev, ctxts_in[i], &ctxts[i]))) return fail(rc);
            hec_ct_set_scale(ctxts[i], hec_ct_scale(ctxts[i]) * (double)real_cnum);
        }
    int logStep = 0;
    for (int i = step; i > 1; i /= 2) logStep++;
    int j = HEC_LOGN - logStep;
    while (step >= norm && step >= 1) {
        for (int i = 0; i < step; i += norm) {
            if ((rc = hec_mul_pt_new(ev, ctxts[i + step], idx[logStep], &tmp1))) return fail(rc);
            if ((rc = hec_sub_new(ev, ctxts[i], tmp1, &tmp2))) return fail(rc);
            if ((rc = hec_add(ev, ctxts[i], tmp1, tmp1))) return fail(rc);
            if ((rc = hec_rotate_gal(ev, tmp2, (1ull << j) + 1, tmp2))) return fail(rc);
            if ((rc = hec_add(ev, tmp1, tmp2, ctxts[i]))) return fail(rc);
            hec_ct_free(ev, tmp1);
            hec_ct_free(ev, tmp2);
            tmp1 = tmp2 = nullptr;
        }
        step /= 2;
        logStep--;
        j++;
    }
    *result = ctxts[0];
    ctxts[0] = nullptr;
    fail(0);
    return HEC_OK;
}

// conv_then_pack (conv.go:522-546)
static int conv_then_pack_oplevel(hec_ctx *ev, const hec_ct *ctxt_in, const hec_pt *const *pl_ker, int max_ob, int norm,
                                  double out_scale, const hec_pt *const *plain_idx, hec_ct **out) {
    std::vector<hec_ct *> ctxt_out(max_ob, nullptr);
    int rc;
    auto cleanup = [&]() { for (auto &p : ctxt_out) if (p) { hec_ct_free(ev, p); p = nullptr; } };
    for (int i = 0; i < max_ob; i++)
        if (i % norm == 0) {
            if ((rc = hec_mul_pt_new(ev, ctxt_in, pl_ker[i], &ctxt_out[i]))) { cleanup(); return rc; }
            if ((rc = hec_set_scale(ev, ctxt_out[i], out_scale / (double)(max_ob / norm)))) { cleanup(); return rc; }
        }
    hec_ct *res = nullptr;
    rc = pack_ctxts(ev, ctxt_out, max_ob, max_ob / norm, plain_idx, &res);
    cleanup();
    if (rc) return rc;
    if (out_scale != hec_ct_scale(res) || 0 != hec_ct_level(res)) {
        hec_ct_free(ev, res);
        return ev->fail(HEC_E_SCALE, "LV or scale after conv then pack, inconsistent");
    }
    *out = res;
    return HEC_OK;
}

// =========================================================================================
// baseline (rotation-per-tap) convolution, op-level: evalConv_BN_BL_test's timed interval
// (eval.go:108-131) = preConv_BL (conv.go:120-143) + rot_iters x [postConv_BL (conv.go:146-178)
// + RotateNew + Add (eval.go:118-125)] + bias Add (eval.go:130).  The kernel plaintexts that
// postConv_BL encodes on the host inside its loop (conv.go:165-166) arrive pre-encoded.
// =========================================================================================
int hec_rotate_many(hec_ctx *c, const std::vector<const hec_ct *> &ct, const std::vector<u64> &galEl,
                    const std::vector<hec_ct *> &out, bool hoisted);

extern "C" int hec_conv_bl(hec_ctx *ev, const hec_ct *ct_input, int in_wid, int ker_wid, int rot_iters, int rot_step,
                           const hec_pt *const *pt_taps, const hec_pt *pl_bn_b, hec_ct **out) {
    if (!ev || !ct_input || !pt_taps || !out || ker_wid < 1 || !(ker_wid & 1) || rot_iters < 1) return ev ? ev->fail(HEC_E_INVAL, "conv_bl args") : HEC_E_INVAL;
    cudaSetDevice(ev->device);
    const int ker_size = ker_wid * ker_wid;
    int rc;
    for (int i = 0; i < rot_iters * ker_size; i++)
        if (!pt_taps[i]) return ev->fail(HEC_E_INVAL, "conv_bl: missing tap plaintext");
    // preConv_BL (conv.go:120-143): one hoisted decomposition, k^2 rotations i*in_wid + j
    std::vector<int> rotations;
    for (int i = -(ker_wid / 2); i <= ker_wid / 2; i++)
        for (int j = -(ker_wid / 2); j <= ker_wid / 2; j++) rotations.push_back(i * in_wid + j);
    std::vector<hec_ct *> ct_in_rots(ker_size, nullptr), ct_tmp(rot_iters, nullptr), ct_rot(rot_iters, nullptr);
    auto cleanup = [&]() {
        for (auto p : ct_in_rots) if (p) hec_ct_free(ev, p);
        for (auto p : ct_tmp) if (p) hec_ct_free(ev, p);
        for (auto p : ct_rot) if (p) hec_ct_free(ev, p);
    };
    if ((rc = hec_rotate_hoisted(ev, ct_input, rotations.data(), ker_size, ct_in_rots.data()))) { cleanup(); return rc; }
    // postConv_BL (conv.go:146-178) for every output rotation i: ct_tmp[i] = sum_tap MulNew(ct_in_rots[tap], pl[i][tap])
    // (the MulNew/Add chain is a sum of Montgomery products; one launch covers all i, limbs and both polys)
    int level = ct_input->level;
    for (int i = 0; i < rot_iters * ker_size; i++) level = std::min(level, pt_taps[i]->level);
    const double scale = ct_input->scale * pt_taps[0]->scale;
    std::vector<DotSpec> specs;
    for (int i = 0; i < rot_iters; i++) {
        if ((rc = hec_ct_alloc(ev, level, scale, &ct_tmp[i]))) { cleanup(); return rc; }
        for (int p = 0; p < 2; p++)
            for (int l = 0; l <= level; l++) {
                DotSpec sp;
                for (int t = 0; t < ker_size; t++) {
                    sp.a.push_back(ct_in_rots[t]->limb(p, l));
                    sp.b.push_back(pt_taps[(size_t)i * ker_size + t]->buf + (size_t)l * HEC_N);
                }
                sp.out = ct_tmp[i]->limb(p, l); sp.mod = ev->modQ(l);
                specs.push_back(sp);
            }
    }
    if ((rc = launch_dot(ev, specs))) { cleanup(); return rc; }
    // eval.go:118-125: ct_res = ct_tmp[0] + sum_{i>0} RotateNew(ct_tmp[i], i*rot_step); the rotations are
    // independent of each other and run as one batch
    std::vector<const hec_ct *> rin;
    std::vector<u64> gal;
    std::vector<hec_ct *> rout;
    for (int i = 1; i < rot_iters; i++) {
        if ((rc = hec_ct_alloc(ev, level, scale, &ct_rot[i]))) { cleanup(); return rc; }
        rin.push_back(ct_tmp[i]); gal.push_back(hec_galois_for_rotation(ev, i * rot_step)); rout.push_back(ct_rot[i]);
    }
    if ((rc = hec_rotate_many(ev, rin, gal, rout, false))) { cleanup(); return rc; }
    hec_ct *ct_res = nullptr;
    if ((rc = hec_ct_alloc(ev, level, scale, &ct_res))) { cleanup(); return rc; }
    specs.clear();
    for (int p = 0; p < 2; p++)
        for (int l = 0; l <= level; l++) {
            DotSpec sp;
            sp.a.push_back(ct_tmp[0]->limb(p, l));
            for (int i = 1; i < rot_iters; i++) sp.a.push_back(ct_rot[i]->limb(p, l));
            sp.out = ct_res->limb(p, l); sp.mod = ev->modQ(l);
            specs.push_back(sp);
        }
    rc = launch_dot(ev, specs);
    cleanup();
    if (rc) { hec_ct_free(ev, ct_res); return rc; }
    if (pl_bn_b) {
        if (hec_ct_scale(ct_res) != pl_bn_b->scale) { // eval.go:127-129
            hec_ct_free(ev, ct_res);
            return ev->fail(HEC_E_SCALE, "Different scale between pl_bn_b and ctxt");
        }
        if ((rc = hec_add_pt(ev, ct_res, pl_bn_b))) { hec_ct_free(ev, ct_res); return rc; }
    }
    *out = ct_res;
    return HEC_OK;
}

// =========================================================================================
// "next" row: the mask / rotate / rescale helpers between layers, op-level,
// at any level and alpha.  Masks arrive as host-encoded plaintexts (EncodeNTT, conv.go:312,357).
// =========================================================================================
// ext_ctxt (conv.go:347-371) / one half of bsgs_ctxt (conv.go:303-344) / ext_double_ctxt
// (conv.go:374-414): result = sum_i RotateNew(MulNew(input, pt_i), rot_i) [; Rescale(result, min_scale)]
extern "C" int hec_ext_ctxt(hec_ctx *ev, const hec_ct *input, int n, const int *rots, const hec_pt *const *pts,
                            int do_rescale, double min_scale, hec_ct **out) {
    if (!ev || !input || !rots || !pts || !out || n < 1) return ev ? ev->fail(HEC_E_INVAL, "ext_ctxt args") : HEC_E_INVAL;
    cudaSetDevice(ev->device);
    int rc;
    std::vector<hec_ct *> prod(n, nullptr), rot(n, nullptr);
    auto cleanup = [&]() {
        for (auto p : prod) if (p) hec_ct_free(ev, p);
        for (auto p : rot) if (p) hec_ct_free(ev, p);
    };
    // the n products and the n rotations are independent of each other: one batch each, then one sum
    std::vector<const hec_ct *> rin;
    std::vector<u64> gal;
    for (int i = 0; i < n; i++) {
        if (!pts[i]) { cleanup(); return ev->fail(HEC_E_INVAL, "ext_ctxt: missing mask plaintext"); }
        if ((rc = hec_mul_pt_new(ev, input, pts[i], &prod[i]))) { cleanup(); return rc; }
        if ((rc = hec_ct_alloc(ev, prod[i]->level, prod[i]->scale, &rot[i]))) { cleanup(); return rc; }
        rin.push_back(prod[i]);
        gal.push_back(hec_galois_for_rotation(ev, rots[i]));
    }
    if ((rc = hec_rotate_many(ev, rin, gal, rot, false))) { cleanup(); return rc; }
    hec_ct *result = nullptr;
    int level = rot[0]->level;
    for (int i = 1; i < n; i++) level = std::min(level, rot[i]->level);
    if ((rc = hec_ct_alloc(ev, level, rot[0]->scale, &result))) { cleanup(); return rc; }
    std::vector<DotSpec> specs;
    for (int p = 0; p < 2; p++)
        for (int l = 0; l <= level; l++) {
            DotSpec sp;
            for (int i = 0; i < n; i++) sp.a.push_back(rot[i]->limb(p, l));
            sp.out = result->limb(p, l); sp.mod = ev->modQ(l);
            specs.push_back(sp);
        }
    rc = launch_dot(ev, specs);
    cleanup();
    if (rc) { hec_ct_free(ev, result); return rc; }
    if (do_rescale && (rc = hec_rescale(ev, result, min_scale))) { hec_ct_free(ev, result); return rc; }
    *out = result;
    return HEC_OK;
}
// keep_ctxt (conv.go:417-431): MulNew(input, mask) then Rescale(result, params.Scale())
extern "C" int hec_keep_ctxt(hec_ctx *ev, const hec_ct *input, const hec_pt *mask, double min_scale, hec_ct **out) {
    if (!ev || !input || !mask || !out) return ev ? ev->fail(HEC_E_INVAL, "keep_ctxt args") : HEC_E_INVAL;
    cudaSetDevice(ev->device);
    hec_ct *result = nullptr;
    int rc;
    if ((rc = hec_mul_pt_new(ev, input, mask, &result))) return rc;
    if ((rc = hec_rescale(ev, result, min_scale))) { hec_ct_free(ev, result); return rc; }
    *out = result;
    return HEC_OK;
}

// =========================================================================================
// fused plan
// =========================================================================================
#ifndef HEC_PLAN_CHUNK_DEFAULT
#define HEC_PLAN_CHUNK_DEFAULT 0
#endif
#ifndef HEC_PLAN_CHAINS_DEFAULT
#define HEC_PLAN_CHAINS_DEFAULT 2
#endif
// The exact basis extension P -> Q with ONE special prime computes v = uint64(float64(y) / float64(p)) for the
// canonical residue y < p (L:ring/ring_basis_extension.go:670-713 with a single term): int -> double conversion and
// IEEE division are monotone, so v is 0 up to a threshold and 1 from there on (float64(y)/float64(p) rounds up to 1.0
// for the last few y below p once p > 2^53).  Returns the smallest such y, or ~0 if there is none.  Host arithmetic
// only: the same IEEE operations the reference executes.
extern "C" uint64_t hec_float_quotient_threshold(uint64_t p) {
    auto v = [p](uint64_t y) { volatile double a = (double)y, b = (double)p; volatile double f = a / b; return (uint64_t)f; };
    if (p == 0 || v(p - 1) == 0) return ~0ull;
    uint64_t lo = 0, hi = p - 1; // v(lo) == 0 (y = 0), v(hi) == 1
    while (hi - lo > 1) {
        uint64_t mid = lo + (hi - lo) / 2;
        if (v(mid)) hi = mid; else lo = mid;
    }
    return hi;
}

struct hec_plan {
    hec_ctx *c = nullptr;
    int B = 0, norm = 1, na = 0, M = 0, levels = 0;
    double in_scale = 0, out_scale = 0;
    const u64 **d_ctin = nullptr;
    const ulonglong2 **d_ptk = nullptr;
    u64 *ptk_scaled = nullptr;        // [na][2][N]: kernel plaintexts with the MultByConst constant folded in (set-up only)
    ulonglong2 *ptk_pairs = nullptr;  // the same as Shoup pairs: what the kernels read
    ulonglong2 *key_pairs = nullptr;  // [levels][2 polys][N]: the Q limb of the level-0 key slices as Shoup pairs
    ulonglong2 *mono_pairs = nullptr; // [levels][N]: the pack monomials NTT(X^step) as Shoup pairs
    u64 *bias_plain = nullptr;        // [N]: the bias plaintext as plain residues
    u64 *pool = nullptr; // all scratch / level buffers
    u64 *stage_in = nullptr, *xfinal = nullptr;
    ConvA pa;
    std::vector<ConvB> pb;
    // chunked execution: the M ciphertexts of a run are worked through in chunks of Mc on `nchains` concurrent chains
    // (forked streams inside the captured graph), each chain with scratch of its own, so that what one kernel of a chunk
    // writes is still in L2 when the next kernel of that chunk reads it
    // deferred-transform plan (hec_kernels.cuh (3)): level-0 polynomials as pairs (U, e); 2 + 5 per level + 2 launches
    bool defer = false;
    u64 *efinal = nullptr, *ufinal = nullptr, *wfin = nullptr;
    ulonglong2 *ptk_first = nullptr;  // [2][na/2][N]: per butterfly of the first pack level, ptk'[a] -+ X^step ptk'[b] as pairs
    int Mc = 0, nchains = 1;
    std::vector<cudaStream_t> chain_streams;
    std::vector<cudaEvent_t> chain_events;
    cudaEvent_t ev_fork = nullptr;
    const u64 *bias = nullptr; int bias_mod = 0;
    cudaGraphExec_t exec = nullptr;
    int launches_per_run = 0;
    // pipelined host runs: copies on their own streams, two batches in flight
    cudaStream_t s_in = nullptr, s_out = nullptr;
    cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_comp[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
    cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr;
    u64 *stage_in2[2] = {nullptr, nullptr}, *stage_out2[2] = {nullptr, nullptr};
    int next_ticket = 0;
    // what the captured graph reads besides the plan's own memory (the plaintexts are copied at creation): the
    // rotation keys, kept alive until the plan goes (see hec_swk_drop)
    std::vector<KeyBuf *> ref_keys;
    std::vector<uint64_t> cache_key; // identity of a plan hec_conv_then_pack cached (serials); empty: a caller's plan
};
static void plan_unref_all(hec_plan *p) {
    for (KeyBuf *kb : p->ref_keys)
        if (--kb->plan_refs == 0 && kb->dead) { cudaFree(kb->buf); delete kb; }
    p->ref_keys.clear();
}

// descriptors of chunk `j` (ciphertexts [j*Mc, (j+1)*Mc) of the run) on scratch set `chain`
static void plan_chunk_args(const hec_plan *p, int j, int chain, ConvA &A, std::vector<ConvB> &Bs) {
    const int Mc = p->Mc, na = p->na;
    const size_t jobsA = (size_t)Mc * na * 2, nb0 = (size_t)Mc * std::max(1, na / 2);
    const size_t per_chain = 2 * jobsA + nb0 * 7;
    u64 *sc = p->pa.w1 + (size_t)chain * per_chain * HEC_N; // chain scratch: w1, w2, wb1, wb2, wb3 x2, wb4 x2, z
    A = p->pa;
    A.ctin = p->pa.ctin + (size_t)j * Mc;
    A.w1 = sc; A.w2 = sc + jobsA * HEC_N;
    A.xout = p->pa.xout + (size_t)j * jobsA * HEC_N;
    u64 *wb = sc + 2 * jobsA * HEC_N;
    Bs = p->pb;
    for (size_t l = 0; l < Bs.size(); l++) {
        ConvB &b = Bs[l];
        const size_t nin = (size_t)Mc * (na >> l) * 2, nout = (size_t)Mc * (na >> (l + 1)) * 2;
        b.xin = p->pb[l].xin + (size_t)j * nin * HEC_N;
        b.xout = p->pb[l].xout + (size_t)j * nout * HEC_N;
        b.w1 = wb; b.w2 = wb + nb0 * HEC_N; b.w3 = wb + 2 * nb0 * HEC_N; b.w4 = wb + 4 * nb0 * HEC_N; b.z = wb + 6 * nb0 * HEC_N;
    }
}
static int plan_launch_chunk(hec_plan *p, const ConvA &A, const std::vector<ConvB> &Bs, int Mc, cudaStream_t s,
                             const std::function<void()> &after) {
    hec_ctx *c = p->c;
    dim3 gA = HEC_GRID(HEC_TILES_PER_LIMB, Mc * p->na * 2);
    if (p->defer) {
        // with at least one pack level the U halves of stage A are never written: the first level forms them from ct_in
        if (Bs.empty()) k_defA1<<<gA, HEC_THREADS, 0, s>>>(A, c->dmods);
        else k_convA1<<<gA, HEC_THREADS, 0, s>>>(A, c->dmods);
        after();
        k_defA2<<<gA, HEC_THREADS, 0, s>>>(A, c->dmods);
        after();
        bool first = true;
        for (auto &b : Bs) {
            int nb = b.n / 2;
            dim3 g1 = HEC_GRID(HEC_TILES_PER_LIMB, Mc * nb), g2 = HEC_GRID(HEC_TILES_PER_LIMB, Mc * nb * 2);
            if (first) k_defB1f<<<g1, HEC_THREADS, 0, s>>>(b, c->dmods);
            else k_convB1<<<g1, HEC_THREADS, 0, s>>>(b, c->dmods);
            after();
            k_defB2<<<g1, HEC_THREADS, HEC_DB2_SMEM, s>>>(b, c->dmods);
            after();
            k_convB3<<<g1, HEC_THREADS, HEC_B3_SMEM, s>>>(b, c->dmods);
            after();
            k_defB4<<<g2, HEC_THREADS, HEC_DB4_SMEM, s>>>(b, c->dmods);
            after();
            if (first) k_defB5<true><<<g1, HEC_THREADS, HEC_DB5_SMEM, s>>>(b, c->dmods);
            else k_defB5<false><<<g1, HEC_THREADS, HEC_DB5_SMEM, s>>>(b, c->dmods);
            after();
            first = false;
        }
        dim3 gF = HEC_GRID(HEC_TILES_PER_LIMB, Mc * 2);
        k_defF1<<<gF, HEC_THREADS, 0, s>>>(p->efinal, p->wfin, A.mq0, c->dmods);
        after();
        k_defF2<<<gF, HEC_THREADS, 0, s>>>(p->wfin, p->ufinal, p->bias, p->xfinal, A.mq0, c->dmods);
        after();
        return HEC_OK;
    }
    k_convA1<<<gA, HEC_THREADS, 0, s>>>(A, c->dmods);
    after();
    k_convA2<<<gA, HEC_THREADS, 0, s>>>(A, c->dmods);
    after();
    k_convA3<<<gA, HEC_THREADS, HEC_A3_SMEM, s>>>(A, c->dmods);
    after();
    for (auto &b : Bs) {
        int nb = b.n / 2;
        dim3 g1 = HEC_GRID(HEC_TILES_PER_LIMB, Mc * nb), g2 = HEC_GRID(HEC_TILES_PER_LIMB, Mc * nb * 2);
        k_convB1<<<g1, HEC_THREADS, 0, s>>>(b, c->dmods);
        after();
        k_convB2<<<g1, HEC_THREADS, 0, s>>>(b, c->dmods);
        after();
        k_convB3<<<g1, HEC_THREADS, HEC_B3_SMEM, s>>>(b, c->dmods);
        after();
        k_convB4<<<g2, HEC_THREADS, 0, s>>>(b, c->dmods);
        after();
        k_convB5<<<g2, HEC_THREADS, HEC_B5_SMEM, s>>>(b, c->dmods);
        after();
    }
    return HEC_OK;
}
// all kernels of a run.  serial = true: every chunk on the context's stream, one after the other (profiling); else the
// chains run side by side on forked streams (inside a stream capture this becomes a graph with parallel branches)
static int plan_launch_all(hec_plan *p, const std::function<void()> &after = [] {}, bool serial = false) {
    hec_ctx *c = p->c;
    cudaStream_t s = c->stream;
    const int nchunks = p->M / p->Mc;
    ConvA A;
    std::vector<ConvB> Bs;
    if (nchunks == 1) {
        plan_chunk_args(p, 0, 0, A, Bs);
        plan_launch_chunk(p, A, Bs, p->Mc, s, after);
    } else if (serial) {
        for (int j = 0; j < nchunks; j++) {
            plan_chunk_args(p, j, 0, A, Bs);
            plan_launch_chunk(p, A, Bs, p->Mc, s, after);
        }
    } else {
        cudaEventRecord(p->ev_fork, s);
        for (int ch = 0; ch < p->nchains; ch++) {
            cudaStream_t cs = p->chain_streams[ch];
            cudaStreamWaitEvent(cs, p->ev_fork, 0);
            for (int j = ch; j < nchunks; j += p->nchains) {
                plan_chunk_args(p, j, ch, A, Bs);
                plan_launch_chunk(p, A, Bs, p->Mc, cs, after);
            }
            cudaEventRecord(p->chain_events[ch], cs);
            cudaStreamWaitEvent(s, p->chain_events[ch], 0);
        }
    }
    if (p->levels == 0 && p->bias && !p->defer) { // single channel: bias not folded into a B5 epilogue
        EwJobs J;
        for (int m = 0; m < p->M; m++) {
            u64 *x = p->xfinal + (size_t)m * 2 * HEC_N;
            J.j[0].a = x; J.j[0].b = p->bias; J.j[0].out = x; J.j[0].mod = p->pa.mq0; J.j[0].g = 0; J.j[0].s0 = 0; J.j[0].s1 = 0;
            k_ew<EW_ADD><<<dim3(32, 1), 256, 0, s>>>(J, c->dmods);
            after();
        }
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return c->fail(HEC_E_CUDA, std::string("conv launch: ") + cudaGetErrorString(e));
    return HEC_OK;
}

extern "C" void hec_plan_destroy(hec_plan *p) {
    if (!p || !p->c) return;
    cudaSetDevice(p->c->device);
    cudaStreamSynchronize(p->c->stream);
    if (p->exec) cudaGraphExecDestroy(p->exec);
    for (int i = 0; i < 2; i++) {
        if (p->ev_in[i]) cudaEventDestroy(p->ev_in[i]);
        if (p->ev_comp[i]) cudaEventDestroy(p->ev_comp[i]);
        if (p->ev_out[i]) cudaEventDestroy(p->ev_out[i]);
        if (p->stage_in2[i]) cudaFree(p->stage_in2[i]);
        if (p->stage_out2[i]) cudaFree(p->stage_out2[i]);
    }
    for (auto e : p->chain_events) cudaEventDestroy(e);
    for (auto st : p->chain_streams) cudaStreamDestroy(st);
    if (p->ev_fork) cudaEventDestroy(p->ev_fork);
    if (p->ev_t0) cudaEventDestroy(p->ev_t0);
    if (p->ev_t1) cudaEventDestroy(p->ev_t1);
    if (p->s_in) cudaStreamDestroy(p->s_in);
    if (p->s_out) cudaStreamDestroy(p->s_out);
    if (p->pool) cudaFreeAsync(p->pool, p->c->stream);
    if (p->ptk_scaled) cudaFreeAsync(p->ptk_scaled, p->c->stream);
    if (p->ptk_pairs) cudaFreeAsync(p->ptk_pairs, p->c->stream);
    if (p->ptk_first) cudaFreeAsync(p->ptk_first, p->c->stream);
    if (p->key_pairs) cudaFreeAsync(p->key_pairs, p->c->stream);
    if (p->mono_pairs) cudaFreeAsync(p->mono_pairs, p->c->stream);
    if (p->bias_plain) cudaFreeAsync(p->bias_plain, p->c->stream);
    if (p->d_ctin) cudaFreeAsync((void *)p->d_ctin, p->c->stream);
    if (p->d_ptk) cudaFreeAsync((void *)p->d_ptk, p->c->stream);
    plan_unref_all(p);
    delete p;
}

// The deferred-transform plan turns the products with the pack monomials into index shifts of coefficients, so the
// plaintexts the caller passes as pl_idx must BE the monomials X^step of gen_idxNlogs (conv.go:241-254): compare each
// with the transform of the unit vector.  Returns 1 / 0, or a negative error code.
// plan memory comes from the context's stream-ordered pool (freed blocks stay cached there): a caller that builds a plan
// per convolution -- hec_conv_then_pack with fresh kernel plaintexts for every image -- would otherwise pay cudaMalloc /
// cudaFree round trips of hundreds of megabytes to the driver per layer
static cudaError_t plan_alloc(hec_ctx *c, void **p, size_t bytes) { return cudaMallocFromPoolAsync(p, bytes, c->pool, c->stream); }

static int pack_monomials_check(hec_ctx *c, const hec_pt *const *pt_idx, int B, int na) {
    u64 *tmp = nullptr;
    int *flag = nullptr;
    if (plan_alloc(c, (void **)&tmp, 3 * HEC_N * sizeof(u64)) != cudaSuccess) return HEC_E_NOMEM;
    if (plan_alloc(c, (void **)&flag, sizeof(int)) != cudaSuccess) { cudaFreeAsync(tmp, c->stream); return HEC_E_NOMEM; }
    cudaMemsetAsync(flag, 0, sizeof(int), c->stream);
    static const u64 one = 1;
    int step = B / 2, logStep = 0, rc = HEC_OK;
    for (int i = step; i > 1; i /= 2) logStep++;
    for (int t = na; t > 1 && !rc; t >>= 1, step /= 2, logStep--) {
        if (!pt_idx[logStep]) { rc = HEC_E_INVAL; break; }
        cudaMemsetAsync(tmp, 0, HEC_N * sizeof(u64), c->stream);
        cudaMemcpyAsync(tmp + step, &one, sizeof(u64), cudaMemcpyHostToDevice, c->stream);
        std::vector<LimbJob> nj;
        nj.push_back({tmp, tmp + HEC_N, c->modQ(0), 0, nullptr, nullptr, 0, 0, nullptr, 0});
        if ((rc = hec_launch_ntt(c, nj, false))) break;
        k_plan_tables<<<64, 256, 0, c->stream>>>(pt_idx[logStep]->buf, nullptr, tmp + 2 * HEC_N, c->modQ(0), c->dmods);
        k_count_diff<<<64, 256, 0, c->stream>>>(tmp + HEC_N, tmp + 2 * HEC_N, flag);
        c->launches += 2;
    }
    int h = 0;
    if (!rc && cudaMemcpyAsync(&h, flag, sizeof(int), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) rc = HEC_E_CUDA;
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) rc = HEC_E_CUDA;
    cudaFreeAsync(tmp, c->stream);
    cudaFreeAsync(flag, c->stream);
    return rc ? rc : (h == 0 ? 1 : 0);
}

static int plan_create_impl(hec_ctx *c, const hec_pt *const *pt_ker, int max_ob, int norm, double in_scale,
                            double out_scale, const hec_pt *const *pt_idx, const hec_pt *pt_bias, int batch,
                            hec_plan **out, bool allow_defer);
extern "C" int hec_plan_create(hec_ctx *c, const hec_pt *const *pt_ker, int max_ob, int norm, double in_scale,
                               double out_scale, const hec_pt *const *pt_idx, const hec_pt *pt_bias, int batch,
                               hec_plan **out) {
    return plan_create_impl(c, pt_ker, max_ob, norm, in_scale, out_scale, pt_idx, pt_bias, batch, out, true);
}
// allow_defer = false: the plan hec_conv_then_pack builds for ONE ciphertext.  Its caller may come back with fresh
// plaintext handles for every image (the network harness does: a new prep_Ker result per layer and image), so the plan
// may run once -- and a deferred plan costs more to set up (monomial check with a synchronisation, per-butterfly tables)
// than one run of it saves (0.2 ms at B = 256).
static int plan_create_impl(hec_ctx *c, const hec_pt *const *pt_ker, int max_ob, int norm, double in_scale,
                            double out_scale, const hec_pt *const *pt_idx, const hec_pt *pt_bias, int batch,
                            hec_plan **out, bool allow_defer) {
    if (!c || !pt_ker || !pt_idx || !out || batch < 1 || max_ob < 1 || norm < 1) return c ? c->fail(HEC_E_INVAL, "plan args") : HEC_E_INVAL;
    cudaSetDevice(c->device);
    if ((max_ob & (max_ob - 1)) || (norm & (norm - 1)) || norm > max_ob || max_ob > 4096)
        return c->fail(HEC_E_UNSUPPORTED, "max_ob and norm must be powers of two, max_ob <= 4096");
    if (c->nQ < 2 || c->nP != 1) return c->fail(HEC_E_UNSUPPORTED, "fused conv needs >= 2 Q limbs and exactly one special prime (main.go:446-454)");
    // the epilogues of A3 / B5 add a few q to an uncorrected forward transform (< 68q) and B5 sums up to 9 lazy terms
    // before one reduction: needs q0 < 2^57; the quotient estimate of reduce_lazy needs q0 > 2^40
    if (c->q(c->modQ(0)) >= (1ull << 57) || c->q(c->modQ(0)) <= (1ull << 40))
        return c->fail(HEC_E_UNSUPPORTED, "fused conv needs a first modulus between 2^40 and 2^57 (use HEC_CONV_OPLEVEL)");
    const int B = max_ob, na = B / norm, M = batch;
    for (int i = 0; i < B; i += norm)
        if (!pt_ker[i] || pt_ker[i]->level < 1) return c->fail(HEC_E_LEVEL, "kernel plaintexts must be at level >= 1 (ECD_LV)");
    // ---- SetScale bookkeeping (L:ckks/evaluator.go:1194-1209, 1291-1325) ----
    const double pt_scale = pt_ker[0]->scale;
    const double target = out_scale / (double)(B / norm);
    double s = in_scale * pt_scale;
    std::vector<u64> k;
    double up = hec_const_limbs(c, 1, target / s, k);
    s *= up;
    int nb = 0;
    while (1 - nb > 0 && s / (double)c->q(1 - nb) >= target / 2) { s /= (double)c->q(1 - nb); nb++; }
    double res_scale = target * (double)(B / norm); // pack_ctxts conv.go:274
    if (nb != 1 || res_scale != out_scale) return c->fail(HEC_E_SCALE, "LV or scale after conv then pack, inconsistent");
    // eval.go:252-257: the bias plaintext is encoded at out_scale; any other scale is the reference's second panic
    if (pt_bias && pt_bias->scale != out_scale) return c->fail(HEC_E_SCALE, "LV or scale after conv then pack, inconsistent (bias plaintext scale)");

    hec_plan *p = new hec_plan();
    p->c = c; p->B = B; p->norm = norm; p->na = na; p->M = M; p->in_scale = in_scale; p->out_scale = out_scale;
    int levels = 0;
    for (int t = na; t > 1; t >>= 1) levels++;
    p->levels = levels;
    auto bail = [&](int code, const char *msg) { hec_plan_destroy(p); return c->fail(code, msg); };
    {   // HEC_DEFER=0: every mod-down / rescale ends in the NTT domain (the round-1/2 kernels, (2) in hec_kernels.cuh);
        // 2: deferred whatever the batch.  Default: deferred once a run has at least 64 channel-ciphertexts to spread over
        // the SMs -- below that a run is a chain of under-filled launches and the longer CTAs of k_defB2 / k_defB5 cost more
        // than the transforms saved (measured, B = 16: 0.41 / 0.52 / 0.73 ms deferred against 0.37 / 0.50 / 0.74 ms for
        // 1 / 2 / 4 ciphertexts per run, 3.84 against 4.32 ms for 32)
        static const int env_defer = getenv("HEC_DEFER") ? atoi(getenv("HEC_DEFER")) : 1;
        if (env_defer && B <= 256 && (env_defer == 2 || (allow_defer && (size_t)M * na >= 64))) {
            int ok = pack_monomials_check(c, pt_idx, B, na);
            if (ok < 0) return bail(ok, "checking the pack monomials");
            p->defer = ok == 1;
        }
    }
    const int mq0_ = c->modQ(0);
    const u64 q1inv0 = invmod(c->q(c->modQ(1)) % c->q(mq0_), c->q(mq0_)); // q1^-1 mod q0
    // ---- device memory: pointer tables + one pool ----
    if (plan_alloc(c, (void **)&p->d_ctin, M * sizeof(u64 *)) != cudaSuccess) return bail(HEC_E_NOMEM, "cudaMalloc");
    if (plan_alloc(c, (void **)&p->d_ptk, B * sizeof(ulonglong2 *)) != cudaSuccess) return bail(HEC_E_NOMEM, "cudaMalloc");
    // plan-owned copies of the kernel plaintexts with the MultByConst constants folded in:
    // (ct*pt)*k == ct*(pt*k); one multiply per coefficient at plan creation instead of one per conv
    if (plan_alloc(c, (void **)&p->ptk_scaled, (size_t)na * 2 * HEC_N * sizeof(u64)) != cudaSuccess) return bail(HEC_E_NOMEM, "cudaMalloc");
    if (plan_alloc(c, (void **)&p->ptk_pairs, (size_t)na * 2 * HEC_N * sizeof(ulonglong2)) != cudaSuccess) return bail(HEC_E_NOMEM, "cudaMalloc");
    std::vector<const ulonglong2 *> hk(B, nullptr);
    {
        std::vector<EwJob> ej;
        for (int a = 0; a < na; a++) {
            u64 *dst = p->ptk_scaled + (size_t)a * 2 * HEC_N;
            hk[a * norm] = p->ptk_pairs + (size_t)a * 2 * HEC_N;
            for (int l = 0; l < 2; l++) {
                // deferred plan: the q0 limb also carries the 1/q1 of the rescale (U = ct*pt*k0/q1)
                const u64 ql = c->q(c->modQ(l)), kl = (p->defer && l == 0) ? mulmod(k[0] % ql, q1inv0, ql) : k[l];
                ej.push_back(ewjob(pt_ker[a * norm]->buf + (size_t)l * HEC_N, nullptr, dst + (size_t)l * HEC_N, c->modQ(l), mform(kl, ql)));
            }
        }
        if (launch_ew<EW_MULSCALAR>(c, ej)) return bail(HEC_E_CUDA, "scaling kernel plaintexts");
        for (int a = 0; a < na; a++)
            for (int l = 0; l < 2; l++) {
                k_plan_tables<<<64, 256, 0, c->stream>>>(p->ptk_scaled + ((size_t)a * 2 + l) * HEC_N, p->ptk_pairs + ((size_t)a * 2 + l) * HEC_N,
                                                          nullptr, c->modQ(l), c->dmods);
                c->launches++;
            }
        if (p->defer && levels >= 1) {
            // first pack level: both inputs of butterfly u are ct_in times a kernel plaintext, so their sum and difference
            // (X^step folded in) are ct_in times ONE table each; built here in Montgomery form, kept as pairs
            const int nbf = na / 2;
            int step0 = B / 2, logStep0 = 0;
            for (int i = step0; i > 1; i /= 2) logStep0++;
            if (!pt_idx[logStep0]) return bail(HEC_E_INVAL, "pt_idx entry missing");
            u64 *tmp = nullptr;
            if (plan_alloc(c, (void **)&p->ptk_first, (size_t)2 * nbf * HEC_N * sizeof(ulonglong2)) != cudaSuccess ||
                plan_alloc(c, (void **)&tmp, (size_t)3 * nbf * HEC_N * sizeof(u64)) != cudaSuccess)
                return bail(HEC_E_NOMEM, "cudaMalloc");
            const int m0 = c->modQ(0);
            std::vector<EwJob> jm, js, ja;
            for (int u = 0; u < nbf; u++) {
                const u64 *pa = p->ptk_scaled + (size_t)u * 2 * HEC_N, *pb = p->ptk_scaled + (size_t)(u + nbf) * 2 * HEC_N;
                u64 *t = tmp + (size_t)3 * u * HEC_N;
                jm.push_back(ewjob(pb, pt_idx[logStep0]->buf, t, m0));
                js.push_back(ewjob(pa, t, t + HEC_N, m0));
                ja.push_back(ewjob(pa, t, t + 2 * HEC_N, m0));
            }
            if (launch_ew<EW_MULMONT>(c, jm) || launch_ew<EW_SUB>(c, js) || launch_ew<EW_ADD>(c, ja)) {
                cudaFreeAsync(tmp, c->stream);
                return bail(HEC_E_CUDA, "first-level plaintext tables");
            }
            for (int u = 0; u < nbf; u++)
                for (int t = 0; t < 2; t++) {
                    k_plan_tables<<<64, 256, 0, c->stream>>>(tmp + (size_t)(3 * u + 1 + t) * HEC_N, p->ptk_first + ((size_t)t * nbf + u) * HEC_N,
                                                              nullptr, m0, c->dmods);
                    c->launches++;
                }
            cudaFreeAsync(tmp, c->stream);
        }
        cudaFreeAsync(p->ptk_scaled, c->stream); // only the pairs are read from here on
        p->ptk_scaled = nullptr;
    }
    // small pageable copy, staged by the driver before the call returns; on the context's stream like the allocation
    if (cudaMemcpyAsync((void *)p->d_ptk, hk.data(), B * sizeof(ulonglong2 *), cudaMemcpyHostToDevice, c->stream) != cudaSuccess) return bail(HEC_E_CUDA, "memcpy ptk");
    // chunking (see hec_plan): HEC_PLAN_CHUNK ciphertexts per chunk (0 / not a divisor of M: the whole batch at once),
    // HEC_PLAN_CHAINS chunks in flight
    {
        static const int env_chunk = getenv("HEC_PLAN_CHUNK") ? atoi(getenv("HEC_PLAN_CHUNK")) : HEC_PLAN_CHUNK_DEFAULT;
        static const int env_chains = getenv("HEC_PLAN_CHAINS") ? atoi(getenv("HEC_PLAN_CHAINS")) : HEC_PLAN_CHAINS_DEFAULT;
        p->Mc = (!p->defer && env_chunk > 0 && env_chunk < M && M % env_chunk == 0) ? env_chunk : M;
        p->nchains = std::max(1, std::min(env_chains, M / p->Mc));
    }
    const int Mc = p->Mc;
    size_t jobsA = (size_t)M * na * 2;          // limbs per level-0 buffer
    size_t jobsAc = (size_t)Mc * na * 2, nb0c = (size_t)Mc * std::max(1, na / 2);
    size_t per_chain = 2 * jobsAc + nb0c * 7;    // w1, w2; wb1, wb2, wb3 x2, wb4 x2, z
    size_t limbs = 4 * (size_t)M                 // staged inputs [M][2][2]
                 + (size_t)p->nchains * per_chain
                 + 2 * jobsA                     // X_0 .. X_last (geometric, < 2x)
                 + (p->defer ? 2 * jobsA + 4 * (size_t)M : 0); // deferred: the e halves of every level, the last transform's scratch, the results
    if (plan_alloc(c, (void **)&p->pool, limbs * HEC_N * sizeof(u64)) != cudaSuccess) return bail(HEC_E_NOMEM, "cudaMalloc plan pool");
    u64 *cur = p->pool;
    auto take = [&](size_t n) { u64 *r = cur; cur += n * HEC_N; return r; };
    p->stage_in = take(4 * (size_t)M);
    u64 *w1 = take((size_t)p->nchains * per_chain), *w2 = nullptr; // chain scratch sets; carved up in plan_chunk_args
    std::vector<u64 *> X(levels + 1);
    for (int l = 0; l <= levels; l++) X[l] = take((size_t)M * (na >> l) * 2);
    std::vector<u64 *> XE(levels + 1, nullptr);
    if (p->defer) {
        for (int l = 0; l <= levels; l++) XE[l] = take((size_t)M * (na >> l) * 2);
        p->wfin = take(2 * (size_t)M);
    }
    u64 *xres = p->defer ? take(2 * (size_t)M) : nullptr;
    u64 *wb1 = nullptr, *wb2 = nullptr, *wb3 = nullptr, *wb4 = nullptr, *wbz = nullptr;
    if (p->M / p->Mc > 1) {
        if (cudaEventCreateWithFlags(&p->ev_fork, cudaEventDisableTiming) != cudaSuccess) return bail(HEC_E_CUDA, "event");
        for (int ch = 0; ch < p->nchains; ch++) {
            cudaStream_t st; cudaEvent_t ev;
            if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) return bail(HEC_E_CUDA, "stream");
            p->chain_streams.push_back(st);
            if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) return bail(HEC_E_CUDA, "event");
            p->chain_events.push_back(ev);
        }
    }
    p->xfinal = p->defer ? xres : X[levels];
    p->ufinal = X[levels]; p->efinal = XE[levels];
    // ---- Stage A constants ----
    const int mq0 = c->modQ(0), mq1 = c->modQ(1), mp0 = c->modP(0);
    const u64 q0 = c->q(mq0), q1 = c->q(mq1), p0 = c->q(mp0);
    ConvA &A = p->pa;
    A.ctin = p->d_ctin; A.ptk = p->d_ptk; A.w1 = w1; A.w2 = w2; A.xout = X[0];
    A.na = na; A.norm = norm; A.mq0 = mq0; A.mq1 = mq1;
    A.half1 = (q1 - 1) >> 1;
    A.hneg0 = q0 - A.half1 % q0;
    auto pair = [](u64 w, u64 q) { return make_ulonglong2(w, (u64)(((u128)w << 64) / q)); };
    A.resc0 = pair(q0 - invmod(q1 % q0, q0), q0);
    A.uout = X[0]; A.eout = XE[0]; A.q1inv = pair(q1inv0, q0);
    if (levels > 0 && (plan_alloc(c, (void **)&p->mono_pairs, (size_t)levels * HEC_N * sizeof(ulonglong2)) != cudaSuccess ||
                       plan_alloc(c, (void **)&p->key_pairs, (size_t)levels * 2 * HEC_N * sizeof(ulonglong2)) != cudaSuccess))
        return bail(HEC_E_NOMEM, "cudaMalloc");
    if (pt_bias) {
        if (plan_alloc(c, (void **)&p->bias_plain, HEC_N * sizeof(u64)) != cudaSuccess) return bail(HEC_E_NOMEM, "cudaMalloc");
        k_plan_tables<<<64, 256, 0, c->stream>>>(pt_bias->buf, nullptr, p->bias_plain, mq0, c->dmods);
        c->launches++;
    }
    // ---- tree levels ----
    int step = B / 2, logStep = 0;
    for (int i = step; i > 1; i /= 2) logStep++;
    int j = HEC_LOGN - logStep;
    if (pt_bias) { p->bias = p->bias_plain; p->bias_mod = mq0; }
    for (int l = 0; l < levels; l++, step /= 2, logStep--, j++) {
        u64 g = (1ull << j) + 1;
        auto it = c->keys.find(g);
        if (it == c->keys.end()) return bail(HEC_E_NOKEY, "rotation key for a pack level is missing");
        if (j < 5) return bail(HEC_E_UNSUPPORTED, "automorphism not local to a 4096-word tile");
        if (!pt_idx[logStep]) return bail(HEC_E_INVAL, "pt_idx entry missing");
        ConvB b;
        memset(&b, 0, sizeof b);
        b.xin = X[l]; b.xout = X[l + 1];
        ulonglong2 *mp = p->mono_pairs + (size_t)l * HEC_N;
        k_plan_tables<<<64, 256, 0, c->stream>>>(pt_idx[logStep]->buf, mp, nullptr, mq0, c->dmods);
        c->launches++;
        b.mono = mp;
        it->second.kb->plan_refs++;
        p->ref_keys.push_back(it->second.kb);
        {   // key slice of this level: the Q limb as pairs (B5), the P limb as it is (B3)
            const int kl = it->second.Lk + c->nP, poff = it->second.Lk;
            ulonglong2 *kp = p->key_pairs + (size_t)l * 2 * HEC_N;
            for (int pc = 0; pc < 2; pc++) {
                // deferred plan: the Q limb of the key carries the 1/P of the mod-down (U = digit * key / P)
                k_plan_tables<<<64, 256, 0, c->stream>>>(it->second.buf + (size_t)(pc * kl) * HEC_N, kp + (size_t)pc * HEC_N, nullptr, mq0, c->dmods,
                                                          p->defer ? invmod(p0 % q0, q0) : 1ull);
                c->launches++;
            }
            b.key = kp;
            b.keyP = it->second.buf + (size_t)poff * HEC_N;
            b.keyPstride = (size_t)kl * HEC_N;
        }
        b.bias = (l == levels - 1 && pt_bias) ? p->bias_plain : nullptr;
        b.w1 = wb1; b.w2 = wb2; b.w3 = wb3; b.w4 = wb4; b.z = wbz;
        b.n = na >> l; b.mq0 = mq0; b.mp0 = mp0;
        b.galEl = (u32)g;
        b.negpinv = pair(q0 - invmod(p0 % q0, q0), q0);
        b.qpj1 = c->pq.qpjinv[mq0][1];
        b.vthr = hec_float_quotient_threshold(p0);
        b.mu0 = (u32)(((u128)1 << 64) / q0);
        if (p->defer) {
            b.ein = XE[l]; b.eout = XE[l + 1];
            u32 gi = 1;
            for (int it2 = 0; it2 < 6; it2++) gi *= 2u - (u32)g * gi; // Newton: g^-1 mod 2^32
            b.ginv = gi & (2u * HEC_N - 1u);
            b.step = (u32)step;
            b.pinv = pair(invmod(p0 % q0, q0), q0);
            b.bias = nullptr; // added by the last transform (k_defF2)
            if (l == 0) {
                b.ctin = p->d_ctin;
                b.ptkz = p->ptk_first;
                b.ptks = p->ptk_first + (size_t)(na / 2) * HEC_N;
            }
        }
        p->pb.push_back(b);
    }
    p->launches_per_run = p->defer ? 4 + 5 * levels : (3 + 5 * levels) * (M / p->Mc) + ((levels == 0 && pt_bias) ? M : 0);
    if (cudaFuncSetAttribute(k_defB2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HEC_DB2_SMEM) != cudaSuccess ||
        cudaFuncSetAttribute(k_defB4, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HEC_DB4_SMEM) != cudaSuccess ||
        cudaFuncSetAttribute(k_defB5<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HEC_DB5_SMEM) != cudaSuccess ||
        cudaFuncSetAttribute(k_defB5<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HEC_DB5_SMEM) != cudaSuccess)
        return bail(HEC_E_CUDA, "cudaFuncSetAttribute(k_defB2/B4/B5)");
    if (cudaFuncSetAttribute(k_convB3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HEC_B3_SMEM) != cudaSuccess ||
        cudaFuncSetAttribute(k_convB5, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HEC_B5_SMEM) != cudaSuccess ||
        cudaFuncSetAttribute(k_convA3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HEC_A3_SMEM) != cudaSuccess)
        return bail(HEC_E_CUDA, "cudaFuncSetAttribute(k_convB3/B5/A3)");
    // ---- capture the kernel sequence in a CUDA graph ----
    cudaGraph_t graph = nullptr;
    if (cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) return bail(HEC_E_CUDA, "begin capture");
    int rc = plan_launch_all(p);
    cudaError_t e = cudaStreamEndCapture(c->stream, &graph);
    if (rc || e != cudaSuccess || !graph) return bail(HEC_E_CUDA, "graph capture failed");
    e = cudaGraphInstantiate(&p->exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) return bail(HEC_E_CUDA, "graph instantiate failed");
    *out = p;
    return HEC_OK;
}

static int plan_run_graph(hec_plan *p, const std::vector<const u64 *> &ptrs) {
    hec_ctx *c = p->c;
    // small pageable copy: staged by the driver at call time, so `ptrs` may die on return
    HEC_CUDA(c, cudaMemcpyAsync((void *)p->d_ctin, ptrs.data(), p->M * sizeof(u64 *), cudaMemcpyHostToDevice, c->stream));
    HEC_CUDA(c, cudaGraphLaunch(p->exec, c->stream));
    c->launches += p->launches_per_run;
    return HEC_OK;
}

extern "C" int hec_plan_run(hec_plan *p, const hec_ct *const *ins, hec_ct **outs) {
    if (!p || !ins || !outs) return HEC_E_INVAL;
    hec_ctx *c = p->c;
    cudaSetDevice(c->device);
    std::vector<const u64 *> ptrs(p->M);
    for (int m = 0; m < p->M; m++) {
        if (!ins[m] || ins[m]->level != 1 || ins[m]->alloc != 2) return c->fail(HEC_E_LEVEL, "plan inputs must be level-1 ciphertexts");
        if (ins[m]->scale != p->in_scale) return c->fail(HEC_E_SCALE, "input scale differs from the plan's");
        ptrs[m] = ins[m]->buf;
    }
    int rc = plan_run_graph(p, ptrs);
    if (rc) return rc;
    for (int m = 0; m < p->M; m++) {
        if (!outs[m] && (rc = hec_ct_alloc(c, 0, p->out_scale, &outs[m]))) return rc;
        if (outs[m]->alloc != 1) return c->fail(HEC_E_LEVEL, "plan outputs must be level-0 ciphertexts");
        outs[m]->level = 0;
        outs[m]->scale = p->out_scale;
        HEC_CUDA(c, cudaMemcpyAsync(outs[m]->buf, p->xfinal + (size_t)m * 2 * HEC_N, 2 * HEC_N * sizeof(u64), cudaMemcpyDeviceToDevice, c->stream));
    }
    return HEC_OK;
}

extern "C" int hec_plan_run_host(hec_plan *p, const uint64_t *const *in_c0, const uint64_t *const *in_c1,
                                 uint64_t *const *out_c0, uint64_t *const *out_c1) {
    if (!p || !in_c0 || !in_c1 || !out_c0 || !out_c1) return HEC_E_INVAL;
    hec_ctx *c = p->c;
    cudaSetDevice(c->device);
    const size_t LB = HEC_N * sizeof(u64);
    std::vector<const u64 *> ptrs(p->M);
    for (int m = 0; m < p->M; m++) {
        u64 *d = p->stage_in + (size_t)m * 4 * HEC_N; // [2 polys][2 limbs][N]
        for (int i = 0; i < 2; i++) {
            HEC_CUDA(c, cudaMemcpyAsync(d + (size_t)(0 * 2 + i) * HEC_N, in_c0[m * 2 + i], LB, cudaMemcpyHostToDevice, c->stream));
            HEC_CUDA(c, cudaMemcpyAsync(d + (size_t)(1 * 2 + i) * HEC_N, in_c1[m * 2 + i], LB, cudaMemcpyHostToDevice, c->stream));
        }
        ptrs[m] = d;
    }
    int rc = plan_run_graph(p, ptrs);
    if (rc) return rc;
    for (int m = 0; m < p->M; m++) {
        const u64 *x = p->xfinal + (size_t)m * 2 * HEC_N;
        HEC_CUDA(c, cudaMemcpyAsync(out_c0[m], x, LB, cudaMemcpyDeviceToHost, c->stream));
        HEC_CUDA(c, cudaMemcpyAsync(out_c1[m], x + HEC_N, LB, cudaMemcpyDeviceToHost, c->stream));
    }
    HEC_CUDA(c, cudaStreamSynchronize(c->stream));
    return HEC_OK;
}

// per-launch device times of one run, kernels launched one by one (not through the graph):
// order A1,A2,A3,(B1..B5) x levels.  For bench.py's roofline of the dominant kernel.
extern "C" int hec_plan_profile(hec_plan *p, const hec_ct *const *ins, float *ms, int cap, int *n) {
    if (!p || !ins || !ms || !n) return HEC_E_INVAL;
    hec_ctx *c = p->c;
    cudaSetDevice(c->device);
    const int nk = p->launches_per_run, nchunks = p->M / p->Mc;
    const int per_chunk = (p->defer ? 4 : 3) + 5 * p->levels, nout = nchunks > 1 ? per_chunk : nk; // chunks are folded: one time per kernel position
    if (cap < nout) return c->fail(HEC_E_INVAL, "profile buffer too small");
    std::vector<const u64 *> ptrs(p->M);
    for (int m = 0; m < p->M; m++) {
        if (!ins[m] || ins[m]->level != 1 || ins[m]->alloc != 2) return c->fail(HEC_E_LEVEL, "plan inputs must be level-1 ciphertexts");
        ptrs[m] = ins[m]->buf;
    }
    HEC_CUDA(c, cudaMemcpyAsync((void *)p->d_ctin, ptrs.data(), p->M * sizeof(u64 *), cudaMemcpyHostToDevice, c->stream));
    std::vector<cudaEvent_t> ev(nk + 1);
    for (auto &e : ev) cudaEventCreate(&e);
    int k = 0;
    cudaEventRecord(ev[0], c->stream);
    int rc = plan_launch_all(p, [&] { if (k < nk) cudaEventRecord(ev[++k], c->stream); }, true);
    cudaStreamSynchronize(c->stream);
    c->launches += nk;
    for (int i = 0; i < nout; i++) ms[i] = 0;
    for (int i = 0; i < k; i++) {
        float t = 0;
        cudaEventElapsedTime(&t, ev[i], ev[i + 1]);
        ms[nchunks > 1 ? i % per_chunk : i] += t;
    }
    for (auto &e : ev) cudaEventDestroy(e);
    *n = std::min(k, nout);
    return rc;
}

// names of the launches hec_plan_profile times, in order, comma separated ("A1,A2,A3,B1,..." -- or, for a plan with
// deferred transforms, "A1,A2,B1,...,B5,...,F1,F2"); returns their number, or a negative code
extern "C" int hec_plan_kernel_names(const hec_plan *p, char *buf, int cap) {
    if (!p || !buf || cap < 1) return HEC_E_INVAL;
    std::string s = p->defer ? "A1,A2" : "A1,A2,A3";
    int n = p->defer ? 2 : 3;
    for (int l = 0; l < p->levels; l++, n += 5) s += ",B1,B2,B3,B4,B5";
    if (p->defer) { s += ",F1,F2"; n += 2; }
    if ((int)s.size() + 1 > cap) return HEC_E_INVAL;
    memcpy(buf, s.c_str(), s.size() + 1);
    return n;
}
extern "C" int hec_plan_is_deferred(const hec_plan *p) { return p && p->defer ? 1 : 0; }
// development aid (not in hec.h): the (U, e) halves a deferred plan left at pack level `level` after its last run,
// [M * (na >> level)][2][N] words each, copied to host memory
extern "C" int hec_plan_debug_level(hec_plan *p, int level, int half, uint64_t *host) {
    if (!p || !p->defer || level < 0 || level > p->levels || !host) return HEC_E_INVAL;
    const u64 *src = level == p->levels ? (half ? p->efinal : p->ufinal) : (half ? p->pb[level].ein : p->pb[level].xin);
    cudaSetDevice(p->c->device);
    cudaStreamSynchronize(p->c->stream);
    size_t words = (size_t)p->M * (p->na >> level) * 2 * HEC_N;
    return cudaMemcpy(host, src, words * sizeof(u64), cudaMemcpyDeviceToHost) == cudaSuccess ? HEC_OK : HEC_E_CUDA;
}

// ---- pipelined host runs -------------------------------------------------------------------
// submit: H2D (stream s_in) -> graph (context stream) -> D2H (stream s_out), double buffered, so
// the copies of one batch overlap the kernels of its neighbours.  wait: block until that
// batch's outputs are in the caller's buffers.  At most two batches in flight.
static int plan_pipeline_init(hec_plan *p) {
    hec_ctx *c = p->c;
    if (p->s_in) return HEC_OK;
    HEC_CUDA(c, cudaStreamCreateWithFlags(&p->s_in, cudaStreamNonBlocking));
    HEC_CUDA(c, cudaStreamCreateWithFlags(&p->s_out, cudaStreamNonBlocking));
    for (int i = 0; i < 2; i++) {
        HEC_CUDA(c, cudaEventCreateWithFlags(&p->ev_in[i], cudaEventDisableTiming));
        HEC_CUDA(c, cudaEventCreateWithFlags(&p->ev_comp[i], cudaEventDisableTiming));
        HEC_CUDA(c, cudaEventCreateWithFlags(&p->ev_out[i], cudaEventDisableTiming));
        HEC_CUDA(c, cudaMalloc(&p->stage_in2[i], (size_t)p->M * 4 * HEC_N * sizeof(u64)));
        HEC_CUDA(c, cudaMalloc(&p->stage_out2[i], (size_t)p->M * 2 * HEC_N * sizeof(u64)));
    }
    HEC_CUDA(c, cudaEventCreate(&p->ev_t0));
    HEC_CUDA(c, cudaEventCreate(&p->ev_t1));
    return HEC_OK;
}
extern "C" int hec_plan_submit_host(hec_plan *p, const uint64_t *const *in_c0, const uint64_t *const *in_c1,
                                    uint64_t *const *out_c0, uint64_t *const *out_c1, int *ticket) {
    if (!p || !in_c0 || !in_c1 || !out_c0 || !out_c1 || !ticket) return HEC_E_INVAL;
    hec_ctx *c = p->c;
    cudaSetDevice(c->device);
    int rc = plan_pipeline_init(p);
    if (rc) return rc;
    const int t = p->next_ticket++, slot = t & 1;
    const size_t LB = HEC_N * sizeof(u64);
    if (t >= 2) HEC_CUDA(c, cudaEventSynchronize(p->ev_out[slot])); // batch t-2 fully delivered
    HEC_CUDA(c, cudaStreamWaitEvent(p->s_in, p->ev_comp[slot], 0)); // kernels of t-2 done with this input slot
    std::vector<const u64 *> ptrs(p->M);
    for (int m = 0; m < p->M; m++) {
        u64 *d = p->stage_in2[slot] + (size_t)m * 4 * HEC_N;
        for (int i = 0; i < 2; i++) {
            HEC_CUDA(c, cudaMemcpyAsync(d + (size_t)i * HEC_N, in_c0[m * 2 + i], LB, cudaMemcpyHostToDevice, p->s_in));
            HEC_CUDA(c, cudaMemcpyAsync(d + (size_t)(2 + i) * HEC_N, in_c1[m * 2 + i], LB, cudaMemcpyHostToDevice, p->s_in));
        }
        ptrs[m] = d;
    }
    HEC_CUDA(c, cudaEventRecord(p->ev_in[slot], p->s_in));
    HEC_CUDA(c, cudaStreamWaitEvent(c->stream, p->ev_in[slot], 0));
    if ((rc = plan_run_graph(p, ptrs))) return rc;
    HEC_CUDA(c, cudaMemcpyAsync(p->stage_out2[slot], p->xfinal, (size_t)p->M * 2 * LB, cudaMemcpyDeviceToDevice, c->stream));
    HEC_CUDA(c, cudaEventRecord(p->ev_comp[slot], c->stream));
    HEC_CUDA(c, cudaStreamWaitEvent(p->s_out, p->ev_comp[slot], 0));
    for (int m = 0; m < p->M; m++) {
        const u64 *x = p->stage_out2[slot] + (size_t)m * 2 * HEC_N;
        HEC_CUDA(c, cudaMemcpyAsync(out_c0[m], x, LB, cudaMemcpyDeviceToHost, p->s_out));
        HEC_CUDA(c, cudaMemcpyAsync(out_c1[m], x + HEC_N, LB, cudaMemcpyDeviceToHost, p->s_out));
    }
    HEC_CUDA(c, cudaEventRecord(p->ev_out[slot], p->s_out));
    *ticket = t;
    return HEC_OK;
}
extern "C" int hec_plan_wait(hec_plan *p, int ticket) {
    if (!p || ticket < 0 || ticket >= p->next_ticket) return HEC_E_INVAL;
    hec_ctx *c = p->c;
    cudaSetDevice(c->device);
    if (ticket < p->next_ticket - 2) return HEC_OK; // already forced complete by a later submit
    HEC_CUDA(c, cudaEventSynchronize(p->ev_out[ticket & 1]));
    return HEC_OK;
}
// device-side span of a pipelined sequence: begin records on the H2D stream, end on the D2H stream
extern "C" int hec_plan_span_begin(hec_plan *p) {
    if (!p) return HEC_E_INVAL;
    hec_ctx *c = p->c;
    cudaSetDevice(c->device);
    int rc = plan_pipeline_init(p);
    if (rc) return rc;
    HEC_CUDA(c, cudaEventRecord(p->ev_t0, p->s_in));
    return HEC_OK;
}
extern "C" int hec_plan_span_end_ms(hec_plan *p, float *ms) {
    if (!p || !ms || !p->s_out) return HEC_E_INVAL;
    hec_ctx *c = p->c;
    cudaSetDevice(c->device);
    HEC_CUDA(c, cudaEventRecord(p->ev_t1, p->s_out));
    HEC_CUDA(c, cudaEventSynchronize(p->ev_t1));
    HEC_CUDA(c, cudaEventElapsedTime(ms, p->ev_t0, p->ev_t1));
    return HEC_OK;
}

// =========================================================================================
// hec_conv_then_pack
// =========================================================================================
// The single-call entry is what a Go caller binds in place of conv_then_pack: one call per
// convolution, the same kernel plaintexts for every image of a layer.  Building a plan (4 allocations, the plaintext
// rescale, a stream capture and a graph instantiation) costs more than the convolution itself, so the plans built here
// are kept in the context, keyed by what the graph reads -- the serials of the kernel / monomial / bias plaintexts and
// of the rotation keys, B, norm and the two scales.  Freeing or replacing any of those objects evicts the plans that
// read it (hec_pt_free, hec_swk_drop, hec_swk_upload).
#define HEC_PLAN_CACHE_MAX 8
void hec_plan_cache_clear(hec_ctx *c) {
    std::vector<hec_plan *> all;
    all.swap(c->plan_cache);
    for (hec_plan *p : all) hec_plan_destroy(p);
}
void hec_plan_cache_evict(hec_ctx *c, uint64_t serial) {
    for (size_t i = 0; i < c->plan_cache.size();) {
        hec_plan *p = c->plan_cache[i];
        if (std::find(p->cache_key.begin(), p->cache_key.end(), serial) != p->cache_key.end()) {
            c->plan_cache.erase(c->plan_cache.begin() + i);
            hec_plan_destroy(p);
        } else {
            i++;
        }
    }
}
extern "C" int hec_plan_cache_size(const hec_ctx *c) { return c ? (int)c->plan_cache.size() : 0; }
static int plan_cached(hec_ctx *c, const hec_pt *const *pt_ker, int max_ob, int norm, double in_scale, double out_scale,
                       const hec_pt *const *pt_idx, const hec_pt *pt_bias, hec_plan **out) {
    std::vector<uint64_t> key;
    auto bits = [](double d) { uint64_t u; memcpy(&u, &d, 8); return u; };
    if (max_ob >= 1 && norm >= 1 && norm <= max_ob && max_ob <= 4096) {
        // serials start at 1 and only grow, so the small integers in front cannot be mistaken for one by evict()
        // once they are offset into the top of the range
        key = {~(uint64_t)max_ob, ~(uint64_t)(norm + 1000), bits(in_scale), bits(out_scale), pt_bias ? pt_bias->serial : 0};
        for (int i = 0; i < max_ob; i += norm) key.push_back(pt_ker[i] ? pt_ker[i]->serial : 0);
        int step = max_ob / 2, logStep = 0;
        for (int i = step; i > 1; i /= 2) logStep++;
        int j = HEC_LOGN - logStep;
        for (int t = max_ob / norm; t > 1; t >>= 1, logStep--, j++) {
            key.push_back(logStep >= 0 && pt_idx[logStep] ? pt_idx[logStep]->serial : 0);
            auto it = c->keys.find((1ull << j) + 1);
            key.push_back(it == c->keys.end() ? 0 : it->second.kb->serial);
        }
        for (size_t i = 0; i < c->plan_cache.size(); i++)
            if (c->plan_cache[i]->cache_key == key) {
                hec_plan *p = c->plan_cache[i];
                c->plan_cache.erase(c->plan_cache.begin() + i);
                c->plan_cache.insert(c->plan_cache.begin(), p);
                *out = p;
                return HEC_OK;
            }
    }
    hec_plan *p = nullptr;
    int rc = plan_create_impl(c, pt_ker, max_ob, norm, in_scale, out_scale, pt_idx, pt_bias, 1, &p, false);
    if (rc) return rc;
    p->cache_key = key;
    c->plan_cache.insert(c->plan_cache.begin(), p);
    while (c->plan_cache.size() > HEC_PLAN_CACHE_MAX) {
        hec_plan *last = c->plan_cache.back();
        c->plan_cache.pop_back();
        hec_plan_destroy(last);
    }
    *out = p;
    return HEC_OK;
}
extern "C" int hec_conv_then_pack(hec_ctx *c, const hec_ct *ct_in, const hec_pt *const *pt_ker, int max_ob, int norm,
                                  double out_scale, const hec_pt *const *pt_idx, const hec_pt *pt_bias, int flags,
                                  hec_ct **out) {
    if (!c || !ct_in || !pt_ker || !pt_idx || !out) return HEC_E_INVAL;
    cudaSetDevice(c->device);
    int rc;
    if (flags == HEC_CONV_OPLEVEL) {
        hec_ct *res = nullptr;
        if ((rc = conv_then_pack_oplevel(c, ct_in, pt_ker, max_ob, norm, out_scale, pt_idx, &res))) return rc;
        if (pt_bias && pt_bias->scale != hec_ct_scale(res)) { // eval.go:252-257
            hec_ct_free(c, res);
            return c->fail(HEC_E_SCALE, "LV or scale after conv then pack, inconsistent (bias plaintext scale)");
        }
        if (pt_bias && (rc = hec_add_pt(c, res, pt_bias))) { hec_ct_free(c, res); return rc; } // eval.go:258
        *out = res;
        return HEC_OK;
    }
    if (ct_in->level != 1) return c->fail(HEC_E_UNSUPPORTED, "fused conv_then_pack expects a level-1 input (ECD_LV = 1)");
    hec_plan *plan = nullptr;
    if ((rc = plan_cached(c, pt_ker, max_ob, norm, ct_in->scale, out_scale, pt_idx, pt_bias, &plan))) return rc;
    hec_ct *res = nullptr;
    const hec_ct *ins[1] = {ct_in};
    hec_ct *tmp = nullptr;
    if (ct_in->alloc != 2) { // compact a ciphertext whose buffer still has dropped limbs
        if ((rc = hec_ct_alloc(c, 1, ct_in->scale, &tmp))) return rc;
        for (int p = 0; p < 2; p++)
            cudaMemcpyAsync(tmp->limb(p, 0), ct_in->limb(p, 0), 2 * HEC_N * sizeof(u64), cudaMemcpyDeviceToDevice, c->stream);
        ins[0] = tmp;
    }
    rc = hec_plan_run(plan, ins, &res);
    if (tmp) hec_ct_free(c, tmp);
    if (rc) { if (res) hec_ct_free(c, res); return rc; }
    *out = res;
    return HEC_OK;
}
