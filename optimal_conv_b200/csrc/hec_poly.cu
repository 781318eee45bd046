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

// AddConst(ct, c, ct) for a real constant (L:ckks/evaluator.go AddConst): the NTT of a constant polynomial is that
// constant in every slot, so c0[j] += scaleUpExact(c, ct.Scale, q_i) for all j; c1 is untouched.
extern "C" int hec_add_const(hec_ctx *c, hec_ct *ct, double constant) {
    if (!c || !ct) return HEC_E_INVAL;
    cudaSetDevice(c->device);
    if (constant == 0) return HEC_OK;
    std::vector<EwJob> jobs;
    for (int i = 0; i <= ct->level; i++) {
        u64 q = c->q(i);
        jobs.push_back(ewjob(ct->limb(0, i), nullptr, ct->limb(0, i), i, scale_up_exact(constant, ct->scale, q) % q));
    }
    return launch_ew<EW_CENTER>(c, jobs);
}

// MultByGaussianIntegerAndAdd(ct, cReal, 0, out) (L:ckks/evaluator.go): out += ct * cReal on both polynomials,
// levels 0..min(ct.Level, out.Level); interfaceMod maps a negative constant to q - (|c| mod q).
extern "C" int hec_mult_by_int_and_add(hec_ctx *c, const hec_ct *ct, int64_t k, hec_ct *out) {
    if (!c || !ct || !out) return HEC_E_INVAL;
    cudaSetDevice(c->device);
    int level = std::min(ct->level, out->level);
    std::vector<EwJob> jobs;
    for (int i = 0; i <= level; i++) {
        u64 q = c->q(i);
        u64 r = k < 0 ? (q - ((u64)(-(k + 1)) + 1) % q) % q : (u64)k % q;
        for (int p = 0; p < 2; p++) jobs.push_back(ewjob(ct->limb(p, i), nullptr, out->limb(p, i), i, mform(r, q)));
    }
    return launch_ew<EW_MULSCALAR_ADD>(c, jobs);
}

namespace {
typedef std::shared_ptr<hec_ct> CtP;
struct PolyEval {
    hec_ctx *c;
    double eval_scale;
    int rc = HEC_OK;
    std::map<int, CtP> C;
    struct Poly { std::vector<double> co; int max_deg; bool lead; int degree() const { return (int)co.size() - 1; } };

    CtP hold(hec_ct *p) { hec_ctx *cc = c; return CtP(p, [cc](hec_ct *x) { hec_ct_free(cc, x); }); }
    CtP zero(int level, double scale) {
        hec_ct *o = nullptr;
        if ((rc = hec_ct_alloc(c, level, scale, &o))) return nullptr;
        if (cudaMemsetAsync(o->buf, 0, (size_t)2 * o->alloc * HEC_N * sizeof(u64), c->stream) != cudaSuccess) {
            rc = c->fail(HEC_E_CUDA, "cudaMemsetAsync");
            hec_ct_free(c, o);
            return nullptr;
        }
        return hold(o);
    }
    CtP mul_relin(const CtP &a, const CtP &b) {
        hec_ct *o = nullptr;
        if ((rc = hec_mul_relin_new(c, a.get(), b.get(), &o))) return nullptr;
        return hold(o);
    }
    // computePowerBasis: C[n] = Rescale(MulRelinNew(C[ceil(n/2)], C[n/2]))
    bool power(int n) {
        if (C.count(n)) return true;
        int a = (n + 1) / 2, b = n >> 1;
        if (!power(a) || !power(b)) return false;
        CtP r = mul_relin(C[a], C[b]);
        if (!r || (rc = hec_rescale(c, r.get(), eval_scale))) return false;
        C[n] = r;
        return true;
    }
    static void split(const Poly &p, int sp, Poly &q, Poly &r) {
        r.co.assign(p.co.begin(), p.co.begin() + sp);
        r.max_deg = p.max_deg == p.degree() ? sp - 1 : p.max_deg - (p.degree() - sp + 1);
        r.lead = false;
        q.co.assign(p.co.begin() + sp, p.co.end());
        q.max_deg = p.max_deg;
        q.lead = p.lead;
    }
    CtP from_basis(double ts, const Poly &p) {
        if (p.degree() == 0) {
            CtP res = zero(C[1]->level, ts);
            if (res && fabs(p.co[0]) > 1e-14) rc = hec_add_const(c, res.get(), p.co[0]);
            return rc ? nullptr : res;
        }
        int lvl = C[p.degree()]->level;
        double qi = (double)c->q(lvl);
        CtP res = zero(lvl, ts * qi);
        if (!res) return nullptr;
        if (fabs(p.co[0]) > 1e-14 && (rc = hec_add_const(c, res.get(), p.co[0]))) return nullptr;
        for (int key = p.degree(); key > 0; key--)
            if (fabs(p.co[key]) > 1e-14) {
                volatile double const_scale = ts * qi / C[key]->scale; // (ts * qi) / scale, as the Go expression associates
                volatile double prod = p.co[key] * const_scale;
                // Go's int64(float64) on amd64 is CVTTSD2SQ: out-of-range (and NaN) give the "integer indefinite" -2^63
                int64_t k = (prod >= -9223372036854775808.0 && prod < 9223372036854775808.0) ? (int64_t)prod : INT64_MIN;
                if ((rc = hec_mult_by_int_and_add(c, C[key].get(), k, res.get()))) return nullptr;
            }
        if ((rc = hec_rescale(c, res.get(), eval_scale))) return nullptr;
        return res;
    }
    static int bitlen(int x) { int n = 0; while (x) { n++; x >>= 1; } return n; }
    CtP recurse(double ts, int ls, int ld, const Poly &p) {
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
        int level = C[nxt]->level - 1;
        if (q.max_deg >= (1 << (ld - 1)) && q.lead) level++;
        double qi = (double)c->q(level);
        volatile double ts_q = ts * qi / C[nxt]->scale;
        CtP res = recurse(ts_q, ls, ld, q);
        if (!res) return nullptr;
        CtP tmp = recurse(ts, ls, ld, r);
        if (!tmp) return nullptr;
        if (res->level > tmp->level)
            while (res->level != tmp->level + 1)
                if ((rc = hec_drop_level(c, res.get(), 1))) return nullptr;
        CtP prod = mul_relin(res, C[nxt]);
        if (!prod) return nullptr;
        if (prod->level > tmp->level) {
            if ((rc = hec_rescale(c, prod.get(), eval_scale))) return nullptr;
            if ((rc = hec_add(c, prod.get(), tmp.get(), prod.get()))) return nullptr;
        } else {
            if ((rc = hec_add(c, prod.get(), tmp.get(), prod.get()))) return nullptr;
            if ((rc = hec_rescale(c, prod.get(), eval_scale))) return nullptr;
        }
        return prod;
    }
    CtP run(const hec_ct *ct, const double *coeffs, int n, double target_scale) {
        int deg = n - 1, ld = bitlen(deg), ls = ld >> 1;
        if (ct->level < ld) { // checkEnoughLevels
            rc = c->fail(HEC_E_LEVEL, std::to_string(ct->level) + " levels < " + std::to_string(ld) + " log(d) -> cannot evaluate");
            return nullptr;
        }
        hec_ct *c1 = nullptr;
        if ((rc = hec_ct_copy_new(c, ct, &c1))) return nullptr;
        C[1] = hold(c1);
        for (int i = 2; i < (1 << ls); i++)
            if (!power(i)) return nullptr;
        for (int i = ls; i < ld; i++)
            if (!power(1 << i)) return nullptr;
        Poly p;
        p.co.assign(coeffs, coeffs + n);
        p.max_deg = deg;
        p.lead = true;
        return recurse(target_scale, ls, ld, p);
    }
};
// detach the result from its shared_ptr without freeing it
static hec_ct *release(hec_ctx *c, CtP &p) {
    hec_ct *o = new hec_ct(*p);
    p->owned = false; // the copy keeps the buffer
    (void)c;
    return o;
}
} // namespace

// EvaluatePoly(ct, NewPoly(coeffs), targetScale) with real coefficients coeffs[0..n-1] (index = degree);
// eval_scale = params.Scale(), the minScale of every internal Rescale (evaluator.scale).
extern "C" int hec_evaluate_poly(hec_ctx *c, const hec_ct *ct, const double *coeffs, int n, double target_scale,
                                 double eval_scale, hec_ct **out) {
    if (!c || !ct || !coeffs || n < 2 || !out) return c ? c->fail(HEC_E_INVAL, "evaluate_poly args") : HEC_E_INVAL;
    cudaSetDevice(c->device);
    PolyEval E{c, eval_scale};
    CtP r = E.run(ct, coeffs, n, target_scale);
    if (!r) return E.rc ? E.rc : HEC_E_INVAL;
    *out = release(c, r);
    return HEC_OK;
}

// evalReLU(params, evaluator, ct, alpha) (conv.go:435-480): sign(x) by three composed minimax polynomials,
// then x * (bconst * sign(x) + aconst).  Consumes 10 levels (3 + 3 + 4); the result has scale ct.Scale * eval_scale
// (Mul + Relinearize, not rescaled -- the caller rescales, as the reference's callers do).
extern "C" int hec_eval_relu(hec_ctx *c, const hec_ct *ct, double alpha, double eval_scale, hec_ct **out) {
    if (!c || !ct || !out) return HEC_E_INVAL;
    cudaSetDevice(c->device);
    static const double P1[] = {0.0, 10.8541842577442, 0.0, -62.2833925211098, 0.0, 114.369227820443, 0.0, -62.8023496973074};
    static const double P2[] = {0.0, 4.13976170985111, 0.0, -5.84997640211679, 0.0, 2.94376255659280, 0.0, -0.454530437460152};
    static const double P3[] = {0.0, 3.29956739043733, 0.0, -7.84227260291355, 0.0, 12.8907764115564, 0.0, -12.4917112584486, 0.0,
                                6.94167991428074, 0.0, -2.04298067399942, 0.0, 0.246407138926031};
    double aconst = (alpha + 1) / 2.0, bconst = (1 - alpha) / 2.0;
    double p3[14];
    for (int i = 0; i < 14; i++) { volatile double v = P3[i] * bconst; p3[i] = v; }
    hec_ct *s1 = nullptr, *s2 = nullptr, *s3 = nullptr, *x = nullptr, *res = nullptr;
    int rc = hec_evaluate_poly(c, ct, P1, 8, eval_scale, eval_scale, &s1);
    if (!rc) rc = hec_evaluate_poly(c, s1, P2, 8, eval_scale, eval_scale, &s2);
    if (!rc) rc = hec_evaluate_poly(c, s2, p3, 14, eval_scale, eval_scale, &s3);
    if (!rc) rc = hec_add_const(c, s3, aconst);                 // AddConstNew(ctxt_sign, aconst)
    if (!rc) rc = hec_ct_copy_new(c, ct, &x);
    if (!rc) rc = hec_drop_level(c, x, x->level - s3->level);   // DropLevel(ctxt_in, ...) (on a copy: the input handle is const)
    if (!rc) rc = hec_mul_relin_new(c, s3, x, &res);            // Mul + Relinearize
    hec_ct_free(c, s1); hec_ct_free(c, s2); hec_ct_free(c, s3); hec_ct_free(c, x);
    if (rc) return rc;
    *out = res;
    return HEC_OK;
}
