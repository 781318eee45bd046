// hec_lt.cu -- hoisted baby-step giant-step linear transform on ciphertexts: LinearTransform(ct, *PtDiagMatrix) ->
// MultiplyByDiagMatrixBSGS (L:ckks/linear_transform.go), the core of the bootstrapper's CoeffsToSlots /
// SlotsToCoeffs.
//
// Every ModDown is a rounding step, so where it happens decides the output bits.  The reference's placement,
// kept here: (1) the baby rotations R_i = sigma_i(KS_noModDown(c1; key_i) + (P c0, 0)) stay in Q||P (one shared
// decomposition: KeyswitchHoistedNoModDown); (2) every giant step j != 0 forms sum_i diag[n1 j + i] * R_i in Q||P,
// ModDown's it, adds the un-rotated term diag[n1 j] * ct in Q, key-switches the c1 part again WITHOUT ModDown and
// accumulates sigma_(n1 j)(.) in Q||P (its c0 part is rotated and added in Q at once); (3) giant step 0 adds its
// inner sum to the same Q||P accumulator; (4) one final ModDown; (5) + diag[0] * ct.
// All steps are batched over rotations / giant steps: one launch sequence each.
#include <set>

struct hec_ptdiag {
    int log_slots = 0, n1 = 0, level = 0, nP = 0;
    double scale = 0;
    std::map<int, int> slot;  // diagonal index -> position in buf
    u64 *buf = nullptr;       // per diagonal: (level+1) Q limbs then nP P limbs, Montgomery + NTT form (PtDiagMatrix.Vec)
    const u64 *q(int d, int l) const { return buf + ((size_t)slot.at(d) * (level + 1 + nP) + l) * HEC_N; }
    const u64 *p(int d, int j) const { return buf + ((size_t)slot.at(d) * (level + 1 + nP) + level + 1 + j) * HEC_N; }
};

// PtDiagMatrix{LogSlots, N1, Level, Scale, Vec map[int][2]*ring.Poly} as the reference's encoder builds it
// (EncodeDiagMatrixAtLvl): limbs[d * (level+1+nP) + t] = Vec[keys[d]][0].Coeffs[t] for t <= level, then
// Vec[keys[d]][1].Coeffs[t - level - 1] (the P part).  keys[d] in [0, 2^log_slots), n1 a power of two.
extern "C" int hec_ptdiag_upload(hec_ctx *c, int log_slots, int n1, int level, double scale, int nd, const int *keys,
                                 const uint64_t *const *limbs, hec_ptdiag **out) {
    if (!c || !keys || !limbs || !out || nd < 1 || n1 < 1 || (n1 & (n1 - 1)) || level < 0 || level >= c->nQ || c->nP == 0)
        return c ? c->fail(HEC_E_INVAL, "ptdiag_upload args") : HEC_E_INVAL;
    cudaSetDevice(c->device);
    hec_ptdiag *m = new hec_ptdiag();
    m->log_slots = log_slots; m->n1 = n1; m->level = level; m->nP = c->nP; m->scale = scale;
    int W = level + 1 + c->nP;
    if (cudaMalloc(&m->buf, (size_t)nd * W * HEC_N * sizeof(u64)) != cudaSuccess) { delete m; return c->fail(HEC_E_NOMEM, "cudaMalloc diagonals"); }
    for (int d = 0; d < nd; d++) {
        if (keys[d] < 0 || keys[d] >= (1 << log_slots) || m->slot.count(keys[d])) { cudaFree(m->buf); delete m; return c->fail(HEC_E_INVAL, "ptdiag_upload: bad diagonal index"); }
        m->slot[keys[d]] = d;
        for (int t = 0; t < W; t++)
            if (cudaMemcpyAsync(m->buf + ((size_t)d * W + t) * HEC_N, limbs[(size_t)d * W + t], HEC_N * sizeof(u64), cudaMemcpyHostToDevice, c->stream) != cudaSuccess) {
                cudaFree(m->buf); delete m;
                return c->fail(HEC_E_CUDA, "cudaMemcpyAsync diagonal");
            }
    }
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) {
        cudaFree(m->buf); delete m;
        return c->fail(HEC_E_CUDA, "ptdiag_upload: copy failed");
    }
    *out = m;
    return HEC_OK;
}
extern "C" void hec_ptdiag_free(hec_ctx *c, hec_ptdiag *m) {
    if (!m) return;
    if (c) { cudaSetDevice(c->device); cudaStreamSynchronize(c->stream); }
    cudaFree(m->buf);
    delete m;
}

extern "C" int hec_linear_transform(hec_ctx *c, const hec_ct *ct, const hec_ptdiag *m, hec_ct **out) {
    if (!c || !ct || !m || !out) return HEC_E_INVAL;
    cudaSetDevice(c->device);
    const int level = std::min(ct->level, m->level), L = level + 1, nP = c->nP, W = L + nP, n1 = m->n1;
    int rc;
    // bsgsIndex: giant step j = k / n1, baby step i = k & (n1 - 1)
    std::map<int, std::vector<int>> index;
    std::set<int> babyset;
    for (auto &kv : m->slot) {
        index[kv.first / n1].push_back(kv.first & (n1 - 1));
        if (kv.first & (n1 - 1)) babyset.insert(kv.first & (n1 - 1));
    }
    std::vector<int> babies(babyset.begin(), babyset.end()), giants, inner_js;
    for (auto &kv : index) {
        if (kv.first) giants.push_back(kv.first);
        for (int i : kv.second)
            if (i) { inner_js.push_back(kv.first); break; }
    }
    const size_t nb = babies.size(), ng = giants.size(), ni = inner_js.size();
    auto find_key = [&](int r, const SwKey *&k) {
        auto it = c->keys.find(hec_galois_for_rotation(c, r));
        if (it == c->keys.end()) return c->fail(HEC_E_NOKEY, "rotation key for step " + std::to_string(r) + " missing");
        k = &it->second;
        return (int)HEC_OK;
    };
    std::vector<const SwKey *> bkeys(nb), gkeys(ng);
    for (size_t b = 0; b < nb; b++) if ((rc = find_key(babies[b], bkeys[b]))) return rc;
    for (size_t g = 0; g < ng; g++) if ((rc = find_key(n1 * giants[g], gkeys[g]))) return rc;

    hec_ct *o = nullptr;
    if ((rc = hec_ct_alloc(c, level, ct->scale * m->scale, &o))) return rc;
    auto bail = [&](int e) { hec_ct_free(c, o); return e; };
    size_t need = L + (1 + ng) * decomp_limbs(c, level) + (2 * nb + ni + 2 * ng + 1) * 2 * (size_t)W + ng * L
                  + (2 * ng + 2) * (size_t)L /* mod-down scratch */ + 2 * ng * (size_t)L /* giant steps without an inner sum */ + 4;
    if ((rc = reserve(c, need))) return bail(rc);
    // a Q||P quadruple: Q0 [L], Q1 [L], P0 [nP], P1 [nP]
    struct QP { u64 *q[2], *p[2]; };
    auto alloc_qp = [&]() { QP x; x.q[0] = c->scratch(L); x.q[1] = c->scratch(L); x.p[0] = c->scratch(nP); x.p[1] = c->scratch(nP); return x; };
    auto limb_of = [&](const QP &x, int poly, int t) { return t < L ? x.q[poly] + (size_t)t * HEC_N : x.p[poly] + (size_t)(t - L) * HEC_N; };
    auto mod_of = [&](int t) { return t < L ? c->modQ(t) : c->modP(t - L); };

    // ---- (1) baby rotations in Q||P
    std::vector<QP> R(nb);
    if (nb) {
        u64 *pc0 = c->scratch(L);
        std::vector<EwJob> ej;
        for (int l = 0; l < L; l++) {
            u64 q = c->q(l), pm = 1 % q;
            for (int j = 0; j < nP; j++) pm = (u64)(((u128)pm * (c->q(c->modP(j)) % q)) % q);
            ej.push_back(ewjob(ct->limb(0, l), nullptr, pc0 + (size_t)l * HEC_N, l, mform(pm, q)));   // MulScalarBigint(c0, P)
        }
        if ((rc = launch_ew<EW_MULSCALAR>(c, ej))) return bail(rc);
        std::vector<Decomp> dc;
        if ((rc = decompose_many(c, level, {ct->limb(1, 0)}, dc))) return bail(rc);
        std::vector<QP> acc(nb);
        std::vector<u64 *> aQ, aP;
        for (size_t b = 0; b < nb; b++) {
            acc[b] = alloc_qp(); R[b] = alloc_qp();
            aQ.push_back(acc[b].q[0]); aP.push_back(acc[b].p[0]);
            aQ.push_back(acc[b].q[1]); aP.push_back(acc[b].p[1]);
        }
        if ((rc = ks_mac_many(c, level, dc, bkeys, aQ, aP))) return bail(rc);
        ej.clear();
        std::vector<EwJob> perm;
        for (size_t b = 0; b < nb; b++) {
            u32 g = (u32)hec_galois_for_rotation(c, babies[b]);
            for (int l = 0; l < L; l++) ej.push_back(ewjob(acc[b].q[0] + (size_t)l * HEC_N, pc0 + (size_t)l * HEC_N, acc[b].q[0] + (size_t)l * HEC_N, l));
            for (int poly = 0; poly < 2; poly++)
                for (int t = 0; t < W; t++) perm.push_back(ewjob(limb_of(acc[b], poly, t), nullptr, limb_of(R[b], poly, t), mod_of(t), 0, g));
        }
        if ((rc = launch_ew<EW_ADD>(c, ej))) return bail(rc);
        if ((rc = launch_ew<EW_PERMUTE>(c, perm))) return bail(rc);
    }
    auto baby_pos = [&](int i) { return (size_t)(std::lower_bound(babies.begin(), babies.end(), i) - babies.begin()); };

    // ---- (2a) inner sums of every giant step, all in one launch
    std::map<int, QP> inner;
    if (ni) {
        std::vector<DotSpec> specs;
        for (int j : inner_js) {
            QP x = alloc_qp();
            inner[j] = x;
            for (int poly = 0; poly < 2; poly++)
                for (int t = 0; t < W; t++) {
                    DotSpec sp;
                    for (int i : index[j]) {
                        if (!i) continue;
                        sp.a.push_back(limb_of(R[baby_pos(i)], poly, t));
                        sp.b.push_back(t < L ? m->q(n1 * j + i, t) : m->p(n1 * j + i, t - L));
                    }
                    sp.out = limb_of(x, poly, t); sp.mod = mod_of(t);
                    specs.push_back(sp);
                }
        }
        if ((rc = launch_dot(c, specs))) return bail(rc);
    }
    // ---- (2b) per giant step: ModDown, un-rotated term, second key switch (no ModDown), rotate
    std::vector<u64 *> t0(ng), t1(ng);
    std::vector<QP> G(ng), GP(ng);
    std::vector<u64 *> pt0(ng);
    if (ng) {
        std::vector<u64 *> mdQ, mdP;
        for (size_t g = 0; g < ng; g++) {
            int j = giants[g];
            if (inner.count(j)) {
                t0[g] = inner[j].q[0]; t1[g] = inner[j].q[1];
                mdQ.push_back(t0[g]); mdP.push_back(inner[j].p[0]);
                mdQ.push_back(t1[g]); mdP.push_back(inner[j].p[1]);
            } else { // only the un-rotated diagonal in this giant step
                t0[g] = c->scratch(L); t1[g] = c->scratch(L);
                if (cudaMemsetAsync(t0[g], 0, (size_t)L * HEC_N * sizeof(u64), c->stream) != cudaSuccess ||
                    cudaMemsetAsync(t1[g], 0, (size_t)L * HEC_N * sizeof(u64), c->stream) != cudaSuccess)
                    return bail(c->fail(HEC_E_CUDA, "linear_transform: cudaMemsetAsync"));
            }
        }
        if (!mdQ.empty() && (rc = moddown_many(c, level, mdQ, mdP))) return bail(rc);
        std::vector<EwJob> zt;
        for (size_t g = 0; g < ng; g++) {
            int j = giants[g];
            if (std::find(index[j].begin(), index[j].end(), 0) == index[j].end()) continue;
            for (int l = 0; l < L; l++) {
                zt.push_back(ewjob(ct->limb(0, l), m->q(n1 * j, l), t0[g] + (size_t)l * HEC_N, l));
                zt.push_back(ewjob(ct->limb(1, l), m->q(n1 * j, l), t1[g] + (size_t)l * HEC_N, l));
            }
        }
        if (!zt.empty() && (rc = launch_ew<EW_MAC>(c, zt))) return bail(rc);
        std::vector<const u64 *> src(t1.begin(), t1.end());
        std::vector<Decomp> dc;
        if ((rc = decompose_many(c, level, src, dc))) return bail(rc);
        std::vector<u64 *> aQ, aP;
        for (size_t g = 0; g < ng; g++) {
            G[g] = alloc_qp(); GP[g] = alloc_qp(); pt0[g] = c->scratch(L);
            aQ.push_back(G[g].q[0]); aP.push_back(G[g].p[0]);
            aQ.push_back(G[g].q[1]); aP.push_back(G[g].p[1]);
        }
        if ((rc = ks_mac_many(c, level, dc, gkeys, aQ, aP))) return bail(rc);
        std::vector<EwJob> perm;
        for (size_t g = 0; g < ng; g++) {
            u32 gal = (u32)hec_galois_for_rotation(c, n1 * giants[g]);
            for (int l = 0; l < L; l++) perm.push_back(ewjob(t0[g] + (size_t)l * HEC_N, nullptr, pt0[g] + (size_t)l * HEC_N, l, 0, gal));
            for (int poly = 0; poly < 2; poly++)
                for (int t = 0; t < W; t++) perm.push_back(ewjob(limb_of(G[g], poly, t), nullptr, limb_of(GP[g], poly, t), mod_of(t), 0, gal));
        }
        if ((rc = launch_ew<EW_PERMUTE>(c, perm))) return bail(rc);
    }
    // ---- (3) + (4): the Q||P accumulator and its single ModDown
    bool have_outer = ng > 0 || inner.count(0);
    QP outer{};
    if (have_outer) {
        outer = alloc_qp();
        std::vector<DotSpec> specs;
        for (int poly = 0; poly < 2; poly++)
            for (int t = 0; t < W; t++) {
                DotSpec sp;
                for (size_t g = 0; g < ng; g++) sp.a.push_back(limb_of(GP[g], poly, t));
                if (inner.count(0)) sp.a.push_back(limb_of(inner[0], poly, t));
                sp.out = limb_of(outer, poly, t); sp.mod = mod_of(t);
                specs.push_back(sp);
            }
        if ((rc = launch_dot(c, specs))) return bail(rc);
        if ((rc = moddown_many(c, level, {outer.q[0], outer.q[1]}, {outer.p[0], outer.p[1]}))) return bail(rc);
    }
    // ---- result: c0 = sum_j sigma_j(t0_j) + ModDown(outer0), c1 = ModDown(outer1); then (5) + diag[0] * ct
    {
        std::vector<DotSpec> specs;
        for (int poly = 0; poly < 2; poly++)
            for (int l = 0; l < L; l++) {
                DotSpec sp;
                if (poly == 0) for (size_t g = 0; g < ng; g++) sp.a.push_back(pt0[g] + (size_t)l * HEC_N);
                if (have_outer) sp.a.push_back(outer.q[poly] + (size_t)l * HEC_N);
                sp.out = o->limb(poly, l); sp.mod = l;
                if (sp.a.empty() && cudaMemsetAsync(sp.out, 0, HEC_N * sizeof(u64), c->stream) != cudaSuccess)
                    return bail(c->fail(HEC_E_CUDA, "linear_transform: cudaMemsetAsync"));
                else specs.push_back(sp);
            }
        if (!specs.empty() && (rc = launch_dot(c, specs))) return bail(rc);
        if (m->slot.count(0)) {
            std::vector<EwJob> zt;
            for (int poly = 0; poly < 2; poly++)
                for (int l = 0; l < L; l++) zt.push_back(ewjob(ct->limb(poly, l), m->q(0, l), o->limb(poly, l), l));
            if ((rc = launch_ew<EW_MAC>(c, zt))) return bail(rc);
        }
    }
    *out = o;
    return HEC_OK;
}

// dft (L:ckks/bootstrap.go): vec = Rescale(LinearTransform(vec, M_k), scale before) over the factor matrices
static int dft_chain(hec_ctx *c, const hec_ct *in, const hec_ptdiag *const *mats, int n, hec_ct **out) {
    hec_ct *cur = nullptr;
    int rc = HEC_OK;
    for (int k = 0; k < n && !rc; k++) {
        const hec_ct *src = cur ? cur : in;
        double scale = src->scale;
        hec_ct *nxt = nullptr;
        rc = hec_linear_transform(c, src, mats[k], &nxt);
        if (!rc) rc = hec_rescale(c, nxt, scale);
        hec_ct_free(c, cur);
        cur = nxt;
        if (rc) { hec_ct_free(c, cur); cur = nullptr; }
    }
    if (rc) return rc;
    if (!cur) return hec_ct_copy_new(c, in, out);
    *out = cur;
    return HEC_OK;
}
// CoeffsToSlots(vec, pDFTInv, eval) (L:ckks/bootstrap.go), full packing (LogSlots = LogN - 1): z = dft(vec);
// ct0 = z + conj(z) (real parts), ct1 = (z - conj(z)) / i (imaginary parts).  Needs the conjugation key (galEl 2N-1).
extern "C" int hec_coeffs_to_slots(hec_ctx *c, const hec_ct *ct, const hec_ptdiag *const *mats, int n, hec_ct **ct0, hec_ct **ct1) {
    if (!c || !ct || !mats || n < 1 || !ct0 || !ct1) return c ? c->fail(HEC_E_INVAL, "coeffs_to_slots args") : HEC_E_INVAL;
    cudaSetDevice(c->device);
    hec_ct *z = nullptr, *zc = nullptr, *re = nullptr, *im = nullptr;
    int rc = dft_chain(c, ct, mats, n, &z);
    if (!rc) rc = hec_ct_copy_new(c, z, &zc);
    if (!rc) rc = hec_conjugate(c, z, zc);
    if (!rc) rc = hec_add_new(c, z, zc, &re);
    if (!rc) rc = hec_sub_new(c, z, zc, &im);
    if (!rc) rc = hec_mult_by_i(c, im, 1);
    if (!rc && mats[0]->log_slots < HEC_LOGN - 1) {
        // sparse packing: the right n/2 slots of both parts are zero -- rotate the imaginary part into them and return
        // one ciphertext (ct1 = NULL), as the reference does
        hec_ct *rot = nullptr;
        rc = hec_rotate_new(c, im, 1 << mats[0]->log_slots, &rot);
        if (!rc) rc = hec_add(c, re, rot, re);
        hec_ct_free(c, rot);
        hec_ct_free(c, im);
        im = nullptr;
    }
    hec_ct_free(c, z); hec_ct_free(c, zc);
    if (rc) { hec_ct_free(c, re); hec_ct_free(c, im); return rc; }
    *ct0 = re; *ct1 = im;
    return HEC_OK;
}
// Bootstrapper.subSum (0x5071e0): ct += Rotate(ct, 2^i) for i = log_slots .. LogN - 2; nothing to do at full packing
extern "C" int hec_sub_sum(hec_ctx *c, hec_ct *ct, int log_slots) {
    if (!c || !ct || log_slots < 0 || log_slots > HEC_LOGN - 1) return c ? c->fail(HEC_E_INVAL, "sub_sum args") : HEC_E_INVAL;
    cudaSetDevice(c->device);
    int rc = HEC_OK;
    for (int i = log_slots; i < HEC_LOGN - 1 && !rc; i++) {
        hec_ct *rot = nullptr;
        rc = hec_rotate_new(c, ct, 1 << i, &rot);
        if (!rc) rc = hec_add(c, ct, rot, ct);
        hec_ct_free(c, rot);
    }
    return rc;
}
// SlotsToCoeffs(ct0, ct1, pDFT, eval): dft(ct0 + i * ct1); ct1 may be NULL
extern "C" int hec_slots_to_coeffs(hec_ctx *c, const hec_ct *ct0, const hec_ct *ct1, const hec_ptdiag *const *mats, int n, hec_ct **out) {
    if (!c || !ct0 || !mats || n < 1 || !out) return c ? c->fail(HEC_E_INVAL, "slots_to_coeffs args") : HEC_E_INVAL;
    cudaSetDevice(c->device);
    hec_ct *sum = nullptr, *t = nullptr;
    int rc = hec_ct_copy_new(c, ct0, &sum);
    if (!rc && ct1) {
        rc = hec_ct_copy_new(c, ct1, &t);
        if (!rc) rc = hec_mult_by_i(c, t, 0);
        if (!rc) rc = hec_add(c, sum, t, sum);
    }
    if (!rc) rc = dft_chain(c, sum, mats, n, out);
    hec_ct_free(c, sum); hec_ct_free(c, t);
    return rc;
}

// =========================================================================================
// Split bootstrapping, first half: BootstrappConv_CtoS of the fork's ckks/bootstrap.go (eval.go:447-459), restated
// from the disassembly of the reference binary (0x506800) and checked against its interpreted output (tests/golden).
// =========================================================================================
// Bootstrapper.modUp (0x507400): the level-0 ciphertext's coefficients, centred around q0, re-expressed modulo every
// q_i of the chain; back to the NTT domain.  Returns a new ciphertext at the top level.
extern "C" int hec_mod_up(hec_ctx *c, const hec_ct *ct, hec_ct **out) {
    if (!c || !ct || !out) return HEC_E_INVAL;
    cudaSetDevice(c->device);
    if (ct->level != 0) return c->fail(HEC_E_LEVEL, "modUp expects a level-0 ciphertext");
    int top = c->nQ - 1, rc;
    hec_ct *o = nullptr;
    if ((rc = hec_ct_alloc(c, top, ct->scale, &o))) return rc;
    auto bail = [&](int e) { hec_ct_free(c, o); return e; };
    if ((rc = reserve(c, 2))) return bail(rc);
    u64 *t = c->scratch(2);
    std::vector<LimbJob> nj;
    for (int p = 0; p < 2; p++) nj.push_back({ct->limb(p, 0), t + (size_t)p * HEC_N, 0, 0});
    if ((rc = hec_launch_ntt(c, nj, true))) return bail(rc);
    u64 q0 = c->q(0);
    std::vector<EwJob> ej;
    nj.clear();
    for (int p = 0; p < 2; p++)
        for (int i = 0; i <= top; i++) {
            EwJob j = ewjob(t + (size_t)p * HEC_N, nullptr, o->limb(p, i), i, q0 % c->q(i));
            j.s1 = q0 >> 1;
            ej.push_back(j);
            nj.push_back({o->limb(p, i), o->limb(p, i), i, 0});
        }
    if ((rc = launch_ew<EW_CENTER_LIFT>(c, ej))) return bail(rc);
    if ((rc = hec_launch_ntt(c, nj, false))) return bail(rc);
    *out = o;
    return HEC_OK;
}

static double go_round(double x) { return x < 0 ? -floor(-x + 0.5) : floor(x + 0.5); } // math.Round: half away from zero

// Bootstrapper.evaluateCheby (0x508540): change of variable, Chebyshev evaluation, SinRescal double-angle steps; the
// evaluator's rescale threshold is the sine scale throughout (evaluateSine sets eval.scale = sinescale)
static int btp_evaluate_cheby(hec_ctx *c, hec_ct **pct, const hec_btp_params *b) {
    hec_ct *ct = *pct;
    volatile double target = b->sinescale;
    for (int i = 0; i < b->sin_rescal; i++) { volatile double prod = target * (double)b->sine_qi[i]; target = sqrt(prod); }
    int rc = HEC_OK;
    if (b->sin_type == 1 || b->sin_type == 2) {
        volatile double den = b->sc_fac * (b->cheby_b - b->cheby_a);
        rc = hec_add_const(c, ct, -0.5 / den);
    }
    hec_ct *r = nullptr;
    if (!rc) rc = hec_evaluate_cheby(c, ct, b->cheby, b->n_cheby, target, b->sinescale, &r);
    if (rc) return rc;
    hec_ct_free(c, ct);
    *pct = ct = r;
    volatile double s = b->sqrt2pi;
    for (int i = 0; i < b->sin_rescal && !rc; i++) {
        s = s * s;
        hec_ct *sq = nullptr;
        rc = hec_mul_relin_new(c, ct, ct, &sq);
        if (rc) break;
        hec_ct_free(c, ct);
        *pct = ct = sq;
        rc = hec_add(c, ct, ct, ct);
        if (!rc) rc = hec_add_const(c, ct, -s);
        if (!rc) rc = hec_rescale(c, ct, b->sinescale);
    }
    return rc;
}

// the same on both halves of a fully packed ciphertext at once (they share level and scale): every step is one launch
// sequence for the two -- half the launches of the sine evaluation, and grids twice as full (a single ciphertext leaves
// most of the 148 SMs idle in the small kernels)
static int btp_evaluate_cheby_pair(hec_ctx *c, hec_ct **p0, hec_ct **p1, const hec_btp_params *b) {
    std::vector<hec_ct *> cts = {*p0, *p1};
    volatile double target = b->sinescale;
    for (int i = 0; i < b->sin_rescal; i++) { volatile double prod = target * (double)b->sine_qi[i]; target = sqrt(prod); }
    int rc = HEC_OK;
    if (b->sin_type == 1 || b->sin_type == 2) {
        volatile double den = b->sc_fac * (b->cheby_b - b->cheby_a);
        rc = add_const_many(c, cts, -0.5 / den);
    }
    std::vector<hec_ct *> r;
    if (!rc) rc = evaluate_poly_many(c, std::vector<const hec_ct *>(cts.begin(), cts.end()), b->cheby, b->n_cheby, target, b->sinescale, r, true);
    if (rc) return rc;
    hec_ct_free(c, cts[0]); hec_ct_free(c, cts[1]);
    *p0 = r[0]; *p1 = r[1];
    cts = r;
    volatile double s = b->sqrt2pi;
    for (int i = 0; i < b->sin_rescal && !rc; i++) {
        s = s * s;
        std::vector<hec_ct *> sq;
        std::vector<const hec_ct *> cc(cts.begin(), cts.end());
        rc = hec_mul_relin_many(c, cc, cc, sq);
        if (rc) break;
        hec_ct_free(c, cts[0]); hec_ct_free(c, cts[1]);
        *p0 = sq[0]; *p1 = sq[1];
        cts = sq;
        rc = add_inplace_many(c, cts, std::vector<const hec_ct *>(cts.begin(), cts.end()));
        if (!rc) rc = add_const_many(c, cts, -s);
        if (!rc) rc = hec_rescale_many(c, cts, b->sinescale);
    }
    return rc;
}

// the common head of Bootstrapp (0x505d00) and BootstrappConv_CtoS (0x506800): to the bootstrapping scale at level 0,
// modUp, ScaleUp, CoeffsToSlots, evaluateSine on both halves
static int btp_until_sine(hec_ctx *c, const hec_ct *ct_in, const hec_btp_params *b, const hec_ptdiag *const *pdftinv, int nmat,
                          hec_ct **ct0, hec_ct **ct1) {
    if (!c || !ct_in || !b || !pdftinv || nmat < 1 || !ct0 || !ct1 || !b->cheby || !b->sine_qi || b->n_sine_qi < b->sin_rescal)
        return c ? c->fail(HEC_E_INVAL, "bootstrap args") : HEC_E_INVAL;
    cudaSetDevice(c->device);
    if (b->arcsine_deg > 0) return c->fail(HEC_E_UNSUPPORTED, "ArcSineDeg > 0 is not implemented");
    hec_ct *ct = nullptr, *up = nullptr, *r0 = nullptr, *r1 = nullptr;
    int rc = hec_ct_copy_new(c, ct_in, &ct);
    if (!rc && ct->level > 1) rc = hec_drop_level(c, ct, ct->level - 1);
    if (!rc) {
        if (ct->level == 1) {                                  // one level available: SetScale to the bootstrapping scale
            rc = hec_set_scale(c, ct, b->prescale);
            if (!rc) rc = hec_drop_level(c, ct, ct->level);
        } else {                                               // level 0: integer ScaleUp
            if (ct->scale > b->prescale) rc = c->fail(HEC_E_SCALE, "ciphetext scale > q/||m||)");
            else {
                double r = go_round(b->prescale / ct->scale);
                rc = hec_mult_by_const(c, ct, r);
                if (!rc) ct->scale = ct->scale * r;
            }
        }
    }
    if (!rc) rc = hec_mod_up(c, ct, &up);
    if (!rc) {
        double r = go_round(b->postscale / up->scale);
        rc = hec_mult_by_const(c, up, r);
        if (!rc) up->scale = up->scale * r;
    }
    if (!rc) rc = hec_sub_sum(c, up, pdftinv[0]->log_slots);
    if (!rc) rc = hec_coeffs_to_slots(c, up, pdftinv, nmat, &r0, &r1);
    hec_ct **halves[2] = {&r0, &r1};
    static const int pair = getenv("HEC_SINE_PAIR") ? atoi(getenv("HEC_SINE_PAIR")) : 1;
    if (!rc && pair && r0 && r1 && r0->level == r1->level && r0->scale == r1->scale) { // evaluateSine on both halves at once
        r0->scale = r0->scale * b->message_ratio;
        r1->scale = r1->scale * b->message_ratio;
        rc = btp_evaluate_cheby_pair(c, &r0, &r1, b);
        for (int h = 0; h < 2 && !rc; h++) {
            hec_ct *x = *halves[h];
            volatile double d = b->postscale * b->message_ratio / b->params_scale;
            x->scale = x->scale / d;
        }
    } else
    for (int h = 0; h < 2 && !rc; h++) {                       // evaluateSine (0x508380); ct1 is NULL under sparse packing
        hec_ct *x = *halves[h];
        if (!x) continue;
        x->scale = x->scale * b->message_ratio;
        rc = btp_evaluate_cheby(c, halves[h], b);
        if (!rc) { x = *halves[h]; volatile double d = b->postscale * b->message_ratio / b->params_scale; x->scale = x->scale / d; }
    }
    hec_ct_free(c, ct); hec_ct_free(c, up);
    if (rc) { hec_ct_free(c, r0); hec_ct_free(c, r1); return rc; }
    *ct0 = r0; *ct1 = r1;
    return HEC_OK;
}

extern "C" int hec_bootstrap_ctos(hec_ctx *c, const hec_ct *ct_in, const hec_btp_params *b, const hec_ptdiag *const *pdftinv, int nmat,
                                  hec_ct **ct0, hec_ct **ct1, double *constant) {
    hec_ct *r0 = nullptr, *r1 = nullptr;
    int rc = btp_until_sine(c, ct_in, b, pdftinv, nmat, &r0, &r1);
    if (rc) return rc;
    // the fork's tail: both halves times (q0 / 2^round(log2 q0)) * params.Scale / postscale, then Rescale
    double q0 = (double)c->q(0);
    volatile double k = q0 / exp2(go_round(log2(q0)));
    k = k * b->params_scale;
    k = k / b->postscale;
    // (under sparse packing the reference binary dereferences its nil second half here, 0x506e8e, and dies; the one
    // ciphertext there is gets the same treatment as at full packing)
    hec_ct *halves[2] = {r0, r1};
    for (int h = 0; h < 2 && !rc; h++)
        if (halves[h]) rc = hec_mult_by_const(c, halves[h], k);
    if (!rc && r0 && r1 && r0->level == r1->level && r0->scale == r1->scale) rc = hec_rescale_many(c, {r0, r1}, b->params_scale);
    else
        for (int h = 0; h < 2 && !rc; h++)
            if (halves[h]) rc = hec_rescale(c, halves[h], b->params_scale);
    if (rc) { hec_ct_free(c, r0); hec_ct_free(c, r1); return rc; }
    *ct0 = r0; *ct1 = r1;
    if (constant) *constant = k;
    return HEC_OK;
}

// btp.Bootstrapp(ct) (test_BL.go:133): the head above, then SlotsToCoeffs with the pDFT factors; the fork returns that
// result as it is (no final rounding of the scale)
extern "C" int hec_bootstrapp(hec_ctx *c, const hec_ct *ct_in, const hec_btp_params *b, const hec_ptdiag *const *pdftinv, int ninv,
                              const hec_ptdiag *const *pdft, int nfwd, hec_ct **out) {
    if (!out || !pdft || nfwd < 1) return c ? c->fail(HEC_E_INVAL, "bootstrapp args") : HEC_E_INVAL;
    hec_ct *r0 = nullptr, *r1 = nullptr;
    int rc = btp_until_sine(c, ct_in, b, pdftinv, ninv, &r0, &r1);
    if (!rc) rc = hec_slots_to_coeffs(c, r0, r1, pdft, nfwd, out);
    hec_ct_free(c, r0); hec_ct_free(c, r1);
    return rc;
}

// btp.BootstrappConv_StoC(ct0, ct1) (eval.go:540-560): in the reference binary it is SlotsToCoeffs(ct0, ct1, btp.pDFT,
// btp.evaluator) inlined into its callers (0x53eb00); the caller then rescales (eval.go:562)
extern "C" int hec_bootstrap_stoc(hec_ctx *c, const hec_ct *ct0, const hec_ct *ct1, const hec_ptdiag *const *pdft, int nmat, hec_ct **out) {
    return hec_slots_to_coeffs(c, ct0, ct1, pdft, nmat, out);
}
