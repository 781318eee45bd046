// hec_poly.cu -- polynomial evaluation on ciphertexts: the evalReLU row.
//
// Mirrors, call for call, what the reference's evalReLU (conv.go:435-480) makes the Lattigo fork do:
// EvaluatePoly (L:ckks/polynomial_evaluation.go: computePowerBasis, recurse, splitCoeffs,
// evaluatePolyFromPowerBasis) on top of MulRelinNew / Rescale / MultByGaussianIntegerAndAdd / AddConst /
// DropLevel / Add-with-scale-matching.  Every scale is the same sequence of IEEE double operations the Go
// code performs, because the scales decide integer constants (int64(coeff * constScale)) and rescale counts.
// Host orchestration only; the arithmetic runs in the kernels of hec_kernels.cuh.
#include <map>
#include <memory>

extern "C" int hec_drop_level(hec_ctx *c, hec_ct *ct, int levels) {
    if (!c || !ct || levels < 0 || levels > ct->level) return c ? c->fail(HEC_E_LEVEL, "DropLevel below level 0") : HEC_E_INVAL;
    ct->level -= levels; // L:ckks/evaluator.go DropLevel: the top limbs are simply forgotten
    return HEC_OK;
}

// MulByPow2(ct, pow2, ct) (L:ckks/evaluator.go MulByPow2 -> ring.MulByPow2Lvl): every coefficient times 2^pow2 mod q_i,
// scale unchanged (eval.go:476: the ReLU output is scaled back by 2^pow)
extern "C" int hec_mul_by_pow2(hec_ctx *c, hec_ct *ct, int pow2) {
    if (!c || !ct || pow2 < 0) return c ? c->fail(HEC_E_INVAL, "mul_by_pow2 args") : HEC_E_INVAL;
    cudaSetDevice(c->device);
    std::vector<EwJob> jobs;
    for (int i = 0; i <= ct->level; i++) {
        u64 q = c->q(i), k = 1 % q;
        for (int b = 0; b < pow2; b++) k = (u64)(((u128)k * 2) % q);
        for (int p = 0; p < 2; p++) jobs.push_back(ewjob(ct->limb(p, i), nullptr, ct->limb(p, i), i, mform(k, q)));
    }
    return launch_ew<EW_MULSCALAR>(c, jobs);
}

// MultByi(ct, ct) / DivByi(ct, ct) (L:ckks/evaluator.go): the product with X^(N/2) in the NTT domain -- the first
// N/2 slots times psi^(N/2) (NttPsi[i][1]), the rest times its negative; DivByi the other way round
extern "C" int hec_mult_by_i(hec_ctx *c, hec_ct *ct, int divide) {
    if (!c || !ct) return HEC_E_INVAL;
    cudaSetDevice(c->device);
    std::vector<EwJob> jobs;
    for (int i = 0; i <= ct->level; i++) {
        u64 q = c->q(i), im = c->hm[i].psi_half_mont, neg = q - im;
        for (int p = 0; p < 2; p++) {
            EwJob j = ewjob(ct->limb(p, i), nullptr, ct->limb(p, i), i, divide ? neg : im);
            j.s1 = divide ? im : neg;
            jobs.push_back(j);
        }
    }
    return launch_ew<EW_MULSCALAR_HALVES>(c, jobs);
}
// Conjugate(ct, out) (L:ckks/evaluator.go): permuteNTT with GaloisElementForRowRotation = 2N - 1
extern "C" int hec_conjugate(hec_ctx *c, const hec_ct *ct, hec_ct *out) {
    if (!c || !ct || !out) return HEC_E_INVAL;
    return hec_rotate_gal(c, ct, 2ull * HEC_N - 1, out);
}

// AddConst(ct, c, ct) for a real constant (L:ckks/evaluator.go AddConst): the NTT of a constant polynomial is that
// constant in every slot, so c0[j] += scaleUpExact(c, ct.Scale, q_i) for all j; c1 is untouched.  Batch form.
static int add_const_many(hec_ctx *c, const std::vector<hec_ct *> &cts, double constant) {
    if (constant == 0) return HEC_OK;
    std::vector<EwJob> jobs;
    for (hec_ct *ct : cts)
        for (int i = 0; i <= ct->level; i++) {
            u64 q = c->q(i);
            jobs.push_back(ewjob(ct->limb(0, i), nullptr, ct->limb(0, i), i, scale_up_exact(constant, ct->scale, q) % q));
        }
    return launch_ew<EW_CENTER>(c, jobs);
}
extern "C" int hec_add_const(hec_ctx *c, hec_ct *ct, double constant) {
    if (!c || !ct) return HEC_E_INVAL;
    cudaSetDevice(c->device);
    return add_const_many(c, {ct}, constant);
}

// MultByGaussianIntegerAndAdd(ct, cReal, 0, out) (L:ckks/evaluator.go): out += ct * cReal on both polynomials,
// levels 0..min(ct.Level, out.Level); interfaceMod maps a negative constant to q - (|c| mod q).  Batch form.
static int mult_int_add_many(hec_ctx *c, const std::vector<const hec_ct *> &cts, int64_t k, const std::vector<hec_ct *> &outs) {
    std::vector<EwJob> jobs;
    for (size_t m = 0; m < cts.size(); m++) {
        int level = std::min(cts[m]->level, outs[m]->level);
        for (int i = 0; i <= level; i++) {
            u64 q = c->q(i);
            u64 r = k < 0 ? (q - ((u64)(-(k + 1)) + 1) % q) % q : (u64)k % q;
            for (int p = 0; p < 2; p++) jobs.push_back(ewjob(cts[m]->limb(p, i), nullptr, outs[m]->limb(p, i), i, mform(r, q)));
        }
    }
    return launch_ew<EW_MULSCALAR_ADD>(c, jobs);
}
extern "C" int hec_mult_by_int_and_add(hec_ctx *c, const hec_ct *ct, int64_t k, hec_ct *out) {
    if (!c || !ct || !out) return HEC_E_INVAL;
    cudaSetDevice(c->device);
    return mult_int_add_many(c, {ct}, k, {out});
}

// Add / Sub(a, b, a) for a batch whose a's share one scale and whose b's share one scale (receiver == first operand):
// evaluateInPlace's scale matching, one launch for the scale-up and one for the add
template <int OP>
static int addsub_inplace_many(hec_ctx *c, const std::vector<hec_ct *> &a, const std::vector<const hec_ct *> &b) {
    size_t n = a.size();
    int level = std::min(a[0]->level, b[0]->level), L = level + 1, rc;
    double sa = a[0]->scale, sb = b[0]->scale;
    for (size_t m = 0; m < n; m++)
        if (std::min(a[m]->level, b[m]->level) != level || a[m]->scale != sa || b[m]->scale != sb)
            return c->fail(HEC_E_INVAL, "batched Add needs common levels and scales");
    std::vector<const u64 *> xb(n * 2 * L);
    for (size_t m = 0; m < n; m++)
        for (int p = 0; p < 2; p++)
            for (int i = 0; i < L; i++) xb[(m * 2 + p) * L + i] = b[m]->limb(p, i);
    std::vector<u64> k;
    if (sa > sb && floor(sa / sb) > 1) { // b * floor(sa/sb) into scratch
        hec_const_limbs(c, level, floor(sa / sb), k);
        if ((rc = reserve(c, n * 2 * (size_t)L))) return rc;
        std::vector<EwJob> mj;
        for (size_t m = 0; m < n; m++) {
            u64 *tmp = c->scratch(2 * (size_t)L);
            for (int p = 0; p < 2; p++)
                for (int i = 0; i < L; i++) {
                    u64 *d = tmp + (size_t)(p * L + i) * HEC_N;
                    mj.push_back(ewjob(b[m]->limb(p, i), nullptr, d, i, mform(k[i], c->q(i))));
                    xb[(m * 2 + p) * L + i] = d;
                }
        }
        if ((rc = launch_ew<EW_MULSCALAR>(c, mj))) return rc;
    } else if (sb > sa && floor(sb / sa) > 1) { // a *= floor(sb/sa) in place
        hec_const_limbs(c, level, floor(sb / sa), k);
        std::vector<EwJob> mj;
        for (size_t m = 0; m < n; m++)
            for (int p = 0; p < 2; p++)
                for (int i = 0; i < L; i++) mj.push_back(ewjob(a[m]->limb(p, i), nullptr, a[m]->limb(p, i), i, mform(k[i], c->q(i))));
        if ((rc = launch_ew<EW_MULSCALAR>(c, mj))) return rc;
    }
    std::vector<EwJob> jobs;
    for (size_t m = 0; m < n; m++)
        for (int p = 0; p < 2; p++)
            for (int i = 0; i < L; i++) jobs.push_back(ewjob(a[m]->limb(p, i), xb[(m * 2 + p) * L + i], a[m]->limb(p, i), i));
    if ((rc = launch_ew<OP>(c, jobs))) return rc;
    for (size_t m = 0; m < n; m++) { a[m]->level = level; a[m]->scale = std::max(sa, sb); }
    return HEC_OK;
}
static int add_inplace_many(hec_ctx *c, const std::vector<hec_ct *> &a, const std::vector<const hec_ct *> &b) {
    return addsub_inplace_many<EW_ADD>(c, a, b);
}

namespace {
typedef std::shared_ptr<hec_ct> CtP;
typedef std::vector<CtP> CtV; // one logical ciphertext of the algorithm = a batch of independent ones, same level and scale
struct PolyEval {
    hec_ctx *c;
    double eval_scale;
    bool cheby = false; // EvaluateCheby: T_n = 2 T_a T_b - T_|a-b| (computePowerBasisCheby), splitCoeffsCheby
    int rc = HEC_OK;
    std::map<int, CtV> C;
    struct Poly { std::vector<double> co; int max_deg; bool lead; int degree() const { return (int)co.size() - 1; } };

    CtP hold(hec_ct *p) { hec_ctx *cc = c; return CtP(p, [cc](hec_ct *x) { hec_ct_free(cc, x); }); }
    static std::vector<hec_ct *> raw(const CtV &v) { std::vector<hec_ct *> r; for (auto &p : v) r.push_back(p.get()); return r; }
    static std::vector<const hec_ct *> craw(const CtV &v) { std::vector<const hec_ct *> r; for (auto &p : v) r.push_back(p.get()); return r; }
    static int level(const CtV &v) { return v[0]->level; }
    static double scale(const CtV &v) { return v[0]->scale; }
    CtV zero(size_t n, int lvl, double sc) {
        CtV v;
        for (size_t m = 0; m < n; m++) {
            hec_ct *o = nullptr;
            if ((rc = hec_ct_alloc(c, lvl, sc, &o))) return CtV();
            v.push_back(hold(o));
            if (cudaMemsetAsync(o->buf, 0, (size_t)2 * o->alloc * HEC_N * sizeof(u64), c->stream) != cudaSuccess) {
                rc = c->fail(HEC_E_CUDA, "cudaMemsetAsync");
                return CtV();
            }
        }
        return v;
    }
    CtV mul_relin(const CtV &a, const CtV &b) {
        std::vector<hec_ct *> o;
        if ((rc = hec_mul_relin_many(c, craw(a), craw(b), o))) return CtV();
        CtV v;
        for (hec_ct *p : o) v.push_back(hold(p));
        return v;
    }
    // Rescale(MulRelinNew(a, b), eval_scale) with the first division fused into the relinearisation (hec.cu)
    CtV mul_relin_rescale(const CtV &a, const CtV &b) {
        std::vector<hec_ct *> o;
        if ((rc = hec_mul_relin_rescale_many(c, craw(a), craw(b), eval_scale, o))) return CtV();
        CtV v;
        for (hec_ct *p : o) v.push_back(hold(p));
        return v;
    }
    // computePowerBasis: C[n] = Rescale(MulRelinNew(C[ceil(n/2)], C[n/2]))
    bool power(int n) {
        if (C.count(n)) return true;
        int a = (n + 1) / 2, b = n >> 1;
        if (!power(a) || !power(b)) return false;
        if (cheby && a != b && !power(a - b)) return false;
        CtV r = mul_relin_rescale(C[a], C[b]);
        if (r.empty()) return false;
        if (cheby) {
            if ((rc = add_inplace_many(c, raw(r), craw(r)))) return false;                       // 2 T_a T_b
            if (a == b) rc = add_const_many(c, raw(r), -1.0);                                    // - T_0
            else rc = addsub_inplace_many<EW_SUB>(c, raw(r), craw(C[a - b]));                    // - T_(a-b)
            if (rc) return false;
        }
        C[n] = r;
        return true;
    }
    void split(const Poly &p, int sp, Poly &q, Poly &r) const {
        r.co.assign(p.co.begin(), p.co.begin() + sp);
        r.max_deg = p.max_deg == p.degree() ? sp - 1 : p.max_deg - (p.degree() - sp + 1);
        r.lead = false;
        q.co.assign(p.co.begin() + sp, p.co.end());
        q.max_deg = p.max_deg;
        q.lead = p.lead;
        if (cheby) // p = q T_sp + r with T_i T_sp = (T_(sp+i) + T_(sp-i)) / 2
            for (int i = sp + 1; i <= p.degree(); i++) {
                volatile double twice = 2 * p.co[i];
                q.co[i - sp] = twice;
                volatile double diff = r.co[sp - (i - sp)] - p.co[i];
                r.co[sp - (i - sp)] = diff;
            }
    }
    CtV from_basis(double ts, const Poly &p) {
        size_t n = C[1].size();
        if (p.degree() == 0) {
            CtV res = zero(n, level(C[1]), ts);
            if (!res.empty() && fabs(p.co[0]) > 1e-14) rc = add_const_many(c, raw(res), p.co[0]);
            return rc ? CtV() : res;
        }
        int lvl = level(C[p.degree()]);
        double qi = (double)c->q(lvl);
        // res = p.co[0] + sum_key C[key] * k_key: NewCiphertext (zero) + AddConst + MultByGaussianIntegerAndAdd per term,
        // formed in one pass (k_lincomb); every C[key] is at a level >= lvl, so each term covers all limbs of res
        CtV res;
        for (size_t m = 0; m < n; m++) {
            hec_ct *o = nullptr;
            if ((rc = hec_ct_alloc(c, lvl, ts * qi, &o))) return CtV();
            res.push_back(hold(o));
        }
        std::vector<int> keys;
        std::vector<int64_t> ks;
        for (int key = p.degree(); key > 0; key--)
            if (fabs(p.co[key]) > 1e-14) {
                volatile double const_scale = ts * qi / scale(C[key]); // (ts * qi) / scale, as the Go expression associates
                volatile double prod = p.co[key] * const_scale;
                // Go's int64(float64) on amd64 is CVTTSD2SQ: out-of-range (and NaN) give the "integer indefinite" -2^63
                int64_t k = (prod >= -9223372036854775808.0 && prod < 9223372036854775808.0) ? (int64_t)prod : INT64_MIN;
                if (level(C[key]) < lvl) { rc = c->fail(HEC_E_LEVEL, "power basis below the leaf's level"); return CtV(); }
                keys.push_back(key); ks.push_back(k);
            }
        {
            const size_t nt = keys.size();
            std::vector<u64> table;          // per job: nt pointers, then nt scalars
            std::vector<LinJob> jobs;
            for (size_t m = 0; m < n; m++)
                for (int pp = 0; pp < 2; pp++)
                    for (int i = 0; i <= lvl; i++) {
                        const u64 q = c->q(i);
                        LinJob J;
                        J.a_off = (long long)table.size(); J.s_off = J.a_off + (long long)nt;
                        J.out = res[m]->limb(pp, i); J.mod = i; J.T = (int)nt;
                        // AddConst touches c0 only: scaleUpExact(c, scale, q_i) in every slot
                        J.cst = (pp == 0 && fabs(p.co[0]) > 1e-14) ? scale_up_exact(p.co[0], res[m]->scale, q) % q : 0;
                        for (size_t t = 0; t < nt; t++) table.push_back((u64)(uintptr_t)C[keys[t]][m]->limb(pp, i));
                        for (size_t t = 0; t < nt; t++) {
                            const int64_t k = ks[t];
                            const u64 r = k < 0 ? (q - ((u64)(-(k + 1)) + 1) % q) % q : (u64)k % q; // interfaceMod
                            table.push_back(mform(r, q));
                        }
                        jobs.push_back(J);
                    }
            std::vector<char> h(table.size() * sizeof(u64) + jobs.size() * sizeof(LinJob));
            memcpy(h.data(), table.data(), table.size() * sizeof(u64));
            memcpy(h.data() + table.size() * sizeof(u64), jobs.data(), jobs.size() * sizeof(LinJob));
            char *dbuf = nullptr;
            if ((rc = stage_cached(c, h, &dbuf))) return CtV();
            const LinJob *dj = reinterpret_cast<const LinJob *>(dbuf + table.size() * sizeof(u64));
            unsigned gx = 32;
            while (gx < 256 && gx * (unsigned)jobs.size() < 592) gx *= 2;
            launch_k(c, k_lincomb, dim3(gx, (unsigned)jobs.size()), dim3(256), dj, reinterpret_cast<const u64 *>(dbuf), c->dmods);
            c->launches += 1;
            if ((rc = check_launch(c, "lincomb"))) return CtV();
        }
        if ((rc = hec_rescale_many(c, raw(res), eval_scale))) return CtV();
        return res;
    }
    static int bitlen(int x) { int n = 0; while (x) { n++; x >>= 1; } return n; }
    CtV recurse(double ts, int ls, int ld, const Poly &p) {
        if (p.degree() < (1 << ls)) {
            if (p.lead && p.max_deg > ((1 << ld) - (1 << (ls - 1))) && ls > 1) {
                ld = bitlen(p.degree());
                return recurse(ts, ld >> 1, ld, p);
            }
            return from_basis(ts, p);
        }
        int nxt = 1 << ls;
        while (nxt < (p.degree() >> 1) + 1) nxt <<= 1;
        Poly q, r;
        split(p, nxt, q, r);
        int lvl = level(C[nxt]) - 1;
        if (q.max_deg >= (1 << (ld - 1)) && q.lead) lvl++;
        double qi = (double)c->q(lvl);
        volatile double ts_q = ts * qi / scale(C[nxt]);
        CtV res = recurse(ts_q, ls, ld, q);
        if (res.empty()) return res;
        CtV tmp = recurse(ts, ls, ld, r);
        if (tmp.empty()) return tmp;
        if (level(res) > level(tmp))
            while (level(res) != level(tmp) + 1)
                for (auto &x : res)
                    if ((rc = hec_drop_level(c, x.get(), 1))) return CtV();
        if (std::min(level(res), level(C[nxt])) > level(tmp)) { // the product's level: rescale first, then add
            CtV prod = mul_relin_rescale(res, C[nxt]);
            if (prod.empty()) return prod;
            if ((rc = add_inplace_many(c, raw(prod), craw(tmp)))) return CtV();
            return prod;
        }
        {   // the product's level <= tmp's: Add first, then Rescale -- fused with the relinearisation when the Add is a plain one
            std::vector<const hec_ct *> ra = craw(res), rb = craw(C[nxt]), rt = craw(tmp);
            if (hec_relin_rescale_fusable(c, ra, rb, eval_scale, &rt)) {
                std::vector<hec_ct *> o;
                if ((rc = hec_mul_relin_rescale_many(c, ra, rb, eval_scale, o, &rt))) return CtV();
                CtV v;
                for (hec_ct *p2 : o) v.push_back(hold(p2));
                return v;
            }
        }
        CtV prod = mul_relin(res, C[nxt]);
        if (prod.empty()) return prod;
        if (level(prod) > level(tmp)) {
            if ((rc = hec_rescale_many(c, raw(prod), eval_scale))) return CtV();
            if ((rc = add_inplace_many(c, raw(prod), craw(tmp)))) return CtV();
        } else {
            if ((rc = add_inplace_many(c, raw(prod), craw(tmp)))) return CtV();
            if ((rc = hec_rescale_many(c, raw(prod), eval_scale))) return CtV();
        }
        return prod;
    }
    CtV run(const std::vector<const hec_ct *> &cts, const double *coeffs, int n, double target_scale) {
        int deg = n - 1, ld = bitlen(deg), ls = ld >> 1;
        for (const hec_ct *ct : cts) {
            if (ct->level != cts[0]->level || ct->scale != cts[0]->scale) {
                rc = c->fail(HEC_E_INVAL, "batched EvaluatePoly needs a common level and scale");
                return CtV();
            }
            if (ct->level < ld) { // checkEnoughLevels
                rc = c->fail(HEC_E_LEVEL, std::to_string(ct->level) + " levels < " + std::to_string(ld) + " log(d) -> cannot evaluate");
                return CtV();
            }
        }
        CtV c1;
        for (const hec_ct *ct : cts) {
            hec_ct *x = nullptr;
            if ((rc = hec_ct_copy_new(c, ct, &x))) return CtV();
            c1.push_back(hold(x));
        }
        C[1] = c1;
        for (int i = 2; i < (1 << ls); i++)
            if (!power(i)) return CtV();
        for (int i = ls; i < ld; i++)
            if (!power(1 << i)) return CtV();
        Poly p;
        p.co.assign(coeffs, coeffs + n);
        p.max_deg = deg;
        p.lead = true;
        return recurse(target_scale, ls, ld, p);
    }
};
// detach a result from its shared_ptr without freeing the buffer
static hec_ct *release(CtP &p) {
    hec_ct *o = new hec_ct(*p);
    p->owned = false; // the copy keeps the buffer
    return o;
}
static int evaluate_poly_many(hec_ctx *c, const std::vector<const hec_ct *> &cts, const double *coeffs, int n, double target_scale,
                              double eval_scale, std::vector<hec_ct *> &outs, bool cheby = false) {
    PolyEval E{c, eval_scale, cheby};
    CtV r = E.run(cts, coeffs, n, target_scale);
    if (r.empty()) return E.rc ? E.rc : HEC_E_INVAL;
    outs.clear();
    for (auto &p : r) outs.push_back(release(p));
    return HEC_OK;
}
} // namespace

// EvaluatePoly(ct, NewPoly(coeffs), targetScale) with real coefficients coeffs[0..n-1] (index = degree);
// eval_scale = params.Scale(), the minScale of every internal Rescale (evaluator.scale).
extern "C" int hec_evaluate_poly(hec_ctx *c, const hec_ct *ct, const double *coeffs, int n, double target_scale,
                                 double eval_scale, hec_ct **out) {
    if (!c || !ct || !coeffs || n < 2 || !out) return c ? c->fail(HEC_E_INVAL, "evaluate_poly args") : HEC_E_INVAL;
    cudaSetDevice(c->device);
    std::vector<hec_ct *> o;
    int rc = evaluate_poly_many(c, {ct}, coeffs, n, target_scale, eval_scale, o);
    if (rc) return rc;
    *out = o[0];
    return HEC_OK;
}

// EvaluateCheby(ct, cheby, targetScale) (L:ckks/polynomial_evaluation.go): coeffs are the coefficients in the Chebyshev
// basis T_0..T_(n-1) on [-1, 1] (the change of variable to [a, b] is the caller's, as in the reference's bootstrapper)
extern "C" int hec_evaluate_cheby(hec_ctx *c, const hec_ct *ct, const double *coeffs, int n, double target_scale,
                                  double eval_scale, hec_ct **out) {
    if (!c || !ct || !coeffs || n < 2 || !out) return c ? c->fail(HEC_E_INVAL, "evaluate_cheby args") : HEC_E_INVAL;
    cudaSetDevice(c->device);
    std::vector<hec_ct *> o;
    int rc = evaluate_poly_many(c, {ct}, coeffs, n, target_scale, eval_scale, o, true);
    if (rc) return rc;
    *out = o[0];
    return HEC_OK;
}

// evalReLU(params, evaluator, ct, alpha) (conv.go:435-480) on n independent ciphertexts of a common level and scale
// (the loop over ct_boots[ul] of eval.go:470-476 as one call): sign(x) by three composed minimax polynomials, then
// x * (bconst * sign(x) + aconst).  Every step is one launch sequence for the whole batch.  Consumes 10 levels
// (3 + 3 + 4); the result has scale ct.Scale * eval_scale (Mul + Relinearize, not rescaled -- the caller rescales,
// as the reference's callers do).
extern "C" int hec_eval_relu_many(hec_ctx *c, const hec_ct *const *cts, int n, double alpha, double eval_scale, hec_ct **outs) {
    if (!c || !cts || n < 1 || !outs) return c ? c->fail(HEC_E_INVAL, "eval_relu args") : HEC_E_INVAL;
    cudaSetDevice(c->device);
    static const double P1[] = {0.0, 10.8541842577442, 0.0, -62.2833925211098, 0.0, 114.369227820443, 0.0, -62.8023496973074};
    static const double P2[] = {0.0, 4.13976170985111, 0.0, -5.84997640211679, 0.0, 2.94376255659280, 0.0, -0.454530437460152};
    static const double P3[] = {0.0, 3.29956739043733, 0.0, -7.84227260291355, 0.0, 12.8907764115564, 0.0, -12.4917112584486, 0.0,
                                6.94167991428074, 0.0, -2.04298067399942, 0.0, 0.246407138926031};
    double aconst = (alpha + 1) / 2.0, bconst = (1 - alpha) / 2.0;
    double p3[14];
    for (int i = 0; i < 14; i++) { volatile double v = P3[i] * bconst; p3[i] = v; }
    std::vector<const hec_ct *> in(cts, cts + n);
    std::vector<hec_ct *> s1, s2, s3, x, res;
    auto cv = [](const std::vector<hec_ct *> &v) { return std::vector<const hec_ct *>(v.begin(), v.end()); };
    int rc = evaluate_poly_many(c, in, P1, 8, eval_scale, eval_scale, s1);
    if (!rc) rc = evaluate_poly_many(c, cv(s1), P2, 8, eval_scale, eval_scale, s2);
    if (!rc) rc = evaluate_poly_many(c, cv(s2), p3, 14, eval_scale, eval_scale, s3);
    if (!rc) rc = add_const_many(c, s3, aconst);                       // AddConstNew(ctxt_sign, aconst)
    for (int m = 0; m < n && !rc; m++) {
        hec_ct *t = nullptr;
        if (!(rc = hec_ct_copy_new(c, cts[m], &t))) {
            x.push_back(t);
            rc = hec_drop_level(c, t, t->level - s3[m]->level);        // DropLevel(ctxt_in, ...) (on a copy: the input handle is const)
        }
    }
    if (!rc) rc = hec_mul_relin_many(c, cv(s3), cv(x), res);           // Mul + Relinearize
    for (auto *v : {&s1, &s2, &s3, &x})
        for (hec_ct *t : *v) hec_ct_free(c, t);
    if (rc) return rc;
    for (int m = 0; m < n; m++) outs[m] = res[m];
    return HEC_OK;
}
extern "C" int hec_eval_relu(hec_ctx *c, const hec_ct *ct, double alpha, double eval_scale, hec_ct **out) {
    if (!ct || !out) return HEC_E_INVAL;
    return hec_eval_relu_many(c, &ct, 1, alpha, eval_scale, out);
}
