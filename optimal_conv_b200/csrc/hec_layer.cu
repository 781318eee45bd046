// hec_layer.cu -- one network layer as one entry point: evalConv_BNRelu_new (eval.go:272-575) on the device.
//
// The reference's routine strings together, on its two evaluators (context.pack_evaluator with one special prime for
// the convolution, context.evaluator with five for everything else, main.go:39-40):
//   evalConv_BN (eval.go:224-263: prep_Ker on the host, conv_then_pack, bias Add)   [once, or twice + Add for "StrConv_sparse"]
//   monomial products MulNew(ct, x^offset) for the odd / strided kinds                (eval.go:318-330, 372-397, 405-419)
//   ct_conv.Scale *= 2^pow                                                            (eval.go:437)
//   btp.BootstrappConv_CtoS                                                           (eval.go:447-459)
//   evalReLU + MulByPow2 on each half                                                 (eval.go:470-476)
//   ext_ctxt / ext_double_ctxt / keep_ctxt per half, by kind                          (eval.go:492-539)
//   btp.BootstrappConv_StoC, Rescale                                                  (eval.go:540-562)
// What stays on the host is what the Go program computes in floats or with its encoder: reshaping and encoding the
// kernels (prep_Ker -> plaintexts, or hec_encode_coeffs_many), the index maps of rot_util.go turned into slot-encoded
// masks (EncodeNTT), the bootstrapper's matrices and sine polynomial.  They arrive here as uploaded handles; the
// `kind` switch of the reference is reduced to which of them are present.
#include <cmath>
#include "hec_host.cuh"

using namespace hec;

int hec_ct_alloc(hec_ctx *c, int level, double scale, hec_ct **out);

// ext_double_ctxt (conv.go:374-414; do_rescale = 1) and bsgs_ctxt (conv.go:303-344; do_rescale = 0):
// mid = sum_i RotateNew(MulNew(input, pts_m[i]), rots_m[i]);  out = sum_i RotateNew(MulNew(mid, pts_r[i]), rots_r[i]) [; Rescale]
extern "C" int hec_ext_double_ctxt(hec_ctx *ev, const hec_ct *input, int n_m, const int *rots_m, const hec_pt *const *pts_m,
                                   int n_r, const int *rots_r, const hec_pt *const *pts_r, int do_rescale, double min_scale,
                                   hec_ct **out) {
    if (!ev || !input || !out) return ev ? ev->fail(HEC_E_INVAL, "ext_double_ctxt args") : HEC_E_INVAL;
    hec_ct *mid = nullptr;
    int rc = hec_ext_ctxt(ev, input, n_m, rots_m, pts_m, 0, 0.0, &mid);
    if (rc) return rc;
    rc = hec_ext_ctxt(ev, mid, n_r, rots_r, pts_r, do_rescale, min_scale, out);
    hec_ct_free(ev, mid);
    return rc;
}

extern "C" int hec_conv_bn_relu(hec_ctx *pack_ev, hec_ctx *ev, const hec_ct *ct_input, const hec_layer_args *a, hec_ct **out) {
    if (!pack_ev || !ev || !ct_input || !a || !out) return ev ? ev->fail(HEC_E_INVAL, "conv_bn_relu args") : HEC_E_INVAL;
    if (a->n_conv < 1 || a->n_conv > 2 || a->iter < 1 || a->iter > 2 || !a->btp || !a->pt_idx)
        return ev->fail(HEC_E_INVAL, "conv_bn_relu: n_conv and iter must be 1 or 2; bootstrapper and monomials are required");
    if (pack_ev->device != ev->device) return ev->fail(HEC_E_INVAL, "conv_bn_relu: the two evaluators must live on one device");
    cudaSetDevice(ev->device);
    int rc = HEC_OK;
    hec_ct *in = nullptr, *conv[2] = {nullptr, nullptr}, *ct_conv = nullptr, *boots[2] = {nullptr, nullptr};
    hec_ct *relu[2] = {nullptr, nullptr}, *keep[2] = {nullptr, nullptr}, *res = nullptr, *tmp = nullptr;
    auto done = [&](int code, hec_ctx *who = nullptr) {
        if (code && who && who != ev) ev->err = who->err; // report the failing evaluator's message through `ev`
        if (in) hec_ct_free(pack_ev, in);
        for (int i = 0; i < 2; i++) {
            if (conv[i]) hec_ct_free(pack_ev, conv[i]);
            if (boots[i]) hec_ct_free(ev, boots[i]);
            if (relu[i]) hec_ct_free(ev, relu[i]);
            if (keep[i]) hec_ct_free(ev, keep[i]);
        }
        if (ct_conv) hec_ct_free(pack_ev, ct_conv);
        if (tmp) hec_ct_free(ev, tmp);
        if (code && res) { hec_ct_free(ev, res); res = nullptr; }
        return code;
    };
    // ---- "StrConv_odd": x^offset onto the input first (eval.go:318-330; MulNew on the main evaluator in the reference --
    //      a plaintext product involves no key, so either evaluator gives the same residues)
    const hec_ct *cin = ct_input;
    if (a->pt_pre) {
        if ((rc = hec_mul_pt_new(pack_ev, ct_input, a->pt_pre, &in))) return done(rc, pack_ev);
        cin = in;
    }
    // ---- evalConv_BN: conv_then_pack + the bias Add (eval.go:250-258), once or twice
    for (int k = 0; k < a->n_conv; k++) {
        if (!a->pt_ker[k]) return done(ev->fail(HEC_E_INVAL, "conv_bn_relu: kernel plaintexts missing"));
        if ((rc = hec_conv_then_pack(pack_ev, cin, a->pt_ker[k], a->max_ob, a->norm[k], a->out_scale, a->pt_idx, a->pt_bias[k],
                                     a->conv_flags, &conv[k])))
            return done(rc, pack_ev);
    }
    if (a->n_conv == 2) { // "StrConv_sparse": second result shifted by x^(norm/4), the two added (eval.go:372-381)
        if (a->pt_shift2) {
            hec_ct *sh = nullptr;
            if ((rc = hec_mul_pt_new(pack_ev, conv[1], a->pt_shift2, &sh))) return done(rc, pack_ev);
            hec_ct_free(pack_ev, conv[1]);
            conv[1] = sh;
        }
        if ((rc = hec_add_new(pack_ev, conv[0], conv[1], &ct_conv))) return done(rc, pack_ev);
    } else {
        ct_conv = conv[0];
        conv[0] = nullptr;
    }
    if (a->pt_post) { // the closing x^0 / -x^offset product of the strided kinds (eval.go:388-397, 405-419)
        hec_ct *sh = nullptr;
        if ((rc = hec_mul_pt_new(pack_ev, ct_conv, a->pt_post, &sh))) return done(rc, pack_ev);
        hec_ct_free(pack_ev, ct_conv);
        ct_conv = sh;
    }
    ct_conv->scale = ct_conv->scale * pow(2.0, a->pow); // eval.go:437
    // the level-0 ciphertext changes evaluator as it is (limb q0 is the same in both chains); the main evaluator's
    // stream must see the pack evaluator's work
    if ((rc = hec_sync(pack_ev))) return done(rc, pack_ev);
    // ---- BootstrappConv_CtoS (eval.go:447-459)
    double cst = 0;
    if ((rc = hec_bootstrap_ctos(ev, ct_conv, a->btp, a->ctos_mats, a->n_ctos, &boots[0], &boots[1], &cst))) return done(rc);
    // ---- evalReLU + MulByPow2 on the halves that exist (eval.go:470-476)
    {
        const hec_ct *ins[2];
        int n = 0, idx[2];
        for (int ul = 0; ul < a->iter; ul++)
            if (boots[ul]) { ins[n] = boots[ul]; idx[n++] = ul; }
        hec_ct *outs[2] = {nullptr, nullptr};
        if (n && (rc = hec_eval_relu_many(ev, ins, n, a->alpha, a->min_scale, outs))) return done(rc);
        for (int k = 0; k < n; k++) {
            relu[idx[k]] = outs[k];
            if ((rc = hec_mul_by_pow2(ev, relu[idx[k]], (int)a->pow))) return done(rc);
        }
    }
    // ---- between-layer data movement per half (eval.go:492-539)
    for (int ul = 0; ul < a->iter; ul++) {
        if (!relu[ul]) continue;
        switch (a->move_kind) {
        case HEC_MOVE_KEEP:
            if (!a->keep_mask[ul]) return done(ev->fail(HEC_E_INVAL, "conv_bn_relu: keep mask missing"));
            rc = hec_keep_ctxt(ev, relu[ul], a->keep_mask[ul], a->min_scale, &keep[ul]);
            break;
        case HEC_MOVE_EXT:
            rc = hec_ext_ctxt(ev, relu[ul], a->n_r[ul], a->rots_r[ul], a->pts_r[ul], 1, a->min_scale, &keep[ul]);
            break;
        case HEC_MOVE_EXT_DOUBLE:
            rc = hec_ext_double_ctxt(ev, relu[ul], a->n_m[ul], a->rots_m[ul], a->pts_m[ul], a->n_r[ul], a->rots_r[ul], a->pts_r[ul],
                                     1, a->min_scale, &keep[ul]);
            break;
        default:
            rc = ev->fail(HEC_E_INVAL, "conv_bn_relu: unknown move_kind");
        }
        if (rc) return done(rc);
    }
    // ---- BootstrappConv_StoC + Rescale (eval.go:540-562); with iter == 1 the second half is nil
    if (!keep[0]) return done(ev->fail(HEC_E_INVAL, "conv_bn_relu: the first half is empty"));
    if ((rc = hec_bootstrap_stoc(ev, keep[0], a->iter == 2 ? keep[1] : nullptr, a->stoc_mats, a->n_stoc, &res))) return done(rc);
    if ((rc = hec_rescale(ev, res, a->min_scale))) return done(rc);
    *out = res;
    return done(HEC_OK);
}
