/*
 * hec.h -- C ABI of libhec.so: a B200-native (sm_100a) CKKS evaluator for the
 * homomorphic-convolution hot path of dwkim606/optimal_conv.
 *
 * This is the drop-in boundary.  The reference's Go code programs
 * against the interface `ckks.Evaluator` held in `context.evaluator` /
 * `context.pack_evaluator` (main.go:39-40); every entry point below names the
 * reference call it replaces.  A cgo shim binds exactly these.
 *
 * Conventions
 *  - plain pointers and sizes only; polynomials are passed as arrays of per-limb
 *    `const uint64_t*` (N contiguous words each, canonical residues in [0,q_i)),
 *    exactly `ring.Poly.Coeffs[limb]` of the Lattigo fork.  NTT domain unless stated.
 *  - the caller owns host buffers and may free/move them when a call returns
 *    (uploads copy; no host pointer is retained -- the cgo pointer rule).
 *  - ciphertexts / plaintexts / keys live on the device behind opaque handles.
 *  - every function returns 0 or a negative HEC_E_* code; hec_last_error(ctx) gives
 *    text.  The Go shim panics on non-zero, which reproduces the reference's
 *    panic-on-inconsistency behaviour (conv.go:541-543, eval.go:252-257).
 *  - a context is NOT re-entrant (like a Lattigo evaluator with shared pools); one
 *    CUDA stream per context; calls are asynchronous on it; download / sync block.
 *  - there is no CPU fallback: every op is a CUDA kernel; if no device is present
 *    hec_ctx_create fails with HEC_E_CUDA.
 */
#ifndef HEC_H
#define HEC_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define HEC_OK 0
#define HEC_E_INVAL (-1)       /* bad argument */
#define HEC_E_CUDA (-2)        /* CUDA runtime error / no device */
#define HEC_E_LEVEL (-3)       /* level or degree mismatch; Rescale at level 0 */
#define HEC_E_NOKEY (-4)       /* rotation key missing (L:ckks/evaluator.go:1575-1597 panic) */
#define HEC_E_SCALE (-5)       /* "LV or scale after conv then pack, inconsistent" (conv.go:541) */
#define HEC_E_UNSUPPORTED (-6) /* shape outside what this build implements */
#define HEC_E_NOMEM (-7)

typedef struct hec_ctx hec_ctx;
typedef struct hec_ct hec_ct;     /* ckks.Ciphertext (degree 1) on device */
typedef struct hec_pt hec_pt;     /* ckks.Plaintext (NTT form) on device */
typedef struct hec_plan hec_plan; /* a prepared evalConv_BN (kernels resident, CUDA graph) */

const char *hec_version(void);

/* ---- context: replaces ckks.NewEvaluator(params, evk) (main.go:430,435,441; conv.go:258).
 * Q/P are params.Q() / params.P(); logN must be 16 in this build. */
int hec_ctx_create(hec_ctx **ctx, int logN, const uint64_t *Q, int nQ, const uint64_t *P, int nP, int device);
void hec_ctx_destroy(hec_ctx *ctx);
const char *hec_last_error(const hec_ctx *ctx);
int hec_sync(hec_ctx *ctx);
/* device-side elapsed time helpers for benchmarks (CUDA events on the context stream) */
int hec_timer_start(hec_ctx *ctx);
int hec_timer_stop_ms(hec_ctx *ctx, float *ms); /* synchronises */
/* number of kernels this context has launched so far (for bench.py's gpu_launches) */
uint64_t hec_launch_count(const hec_ctx *ctx);
/* page-lock / unlock caller memory (e.g. the backing arrays of ring.Poly.Coeffs; Go's heap does not move
 * objects) so that copies to and from it are asynchronous DMA at full PCIe rate */
int hec_host_register(hec_ctx *ctx, void *ptr, size_t bytes);
int hec_host_unregister(hec_ctx *ctx, void *ptr);

/* ---- plaintexts: what Encoder.EncodeCoeffs+ToNTT / EncodeNTT produce (conv.go:513-514,
 * eval.go:242-243).  limbs[0..level]. */
int hec_pt_upload(hec_ctx *ctx, int level, const uint64_t *const *limbs, double scale, hec_pt **out);
void hec_pt_free(hec_ctx *ctx, hec_pt *pt);
/* Encoder.EncodeCoeffs(values, pt) followed by Encoder.ToNTT(pt) on the device (conv.go:513-514 in prep_Ker,
 * eval.go:242-243 for the bias; L:ckks/encoder.go EncodeCoeffs -> L:ckks/utils.go scaleUpVecExact, then
 * ring.NTTLvl): values[0..n_values) are coefficients of X^i, rounded to integers at `scale` exactly as the
 * reference rounds them (half up on the magnitude; exact big-integer path above 2^64), reduced modulo
 * q_0..q_level, coefficients past n_values cleared.  n_values > N is HEC_E_INVAL (the reference panics),
 * and so is a NaN or infinite value.  _many encodes `count` vectors stored back to back (the B kernel
 * plaintexts of one layer) with one copy and batched launches. */
int hec_encode_coeffs(hec_ctx *ctx, const double *values, int n_values, int level, double scale, hec_pt **out);
int hec_encode_coeffs_many(hec_ctx *ctx, const double *values, int count, int n_values, int level, double scale,
                           hec_pt **out /* [count] */);
/* limbs[0..level] <- the plaintext as canonical residues, NTT domain (Plaintext.Value.Coeffs after ToNTT) */
int hec_pt_download(hec_ctx *ctx, const hec_pt *pt, uint64_t *const *limbs);

/* ---- ciphertexts: rlwe.Ciphertext{Value [2]*ring.Poly} + Scale. */
int hec_ct_upload(hec_ctx *ctx, int level, const uint64_t *const *c0, const uint64_t *const *c1,
                  double scale, hec_ct **out);
int hec_ct_download(hec_ctx *ctx, const hec_ct *ct, uint64_t *const *c0, uint64_t *const *c1);
int hec_ct_copy_new(hec_ctx *ctx, const hec_ct *ct, hec_ct **out); /* ct.CopyNew() (conv.go:273) */
int hec_ct_level(const hec_ct *ct);                                /* ct.Level() */
double hec_ct_scale(const hec_ct *ct);                             /* ct.Scale */
void hec_ct_set_scale(hec_ct *ct, double scale);                   /* ct.SetScalingFactor (conv.go:274) */
void hec_ct_free(hec_ctx *ctx, hec_ct *ct);

/* ---- rotation keys: rlwe.SwitchingKey{Value [][2]*ring.Poly} for Galois element galEl,
 * as Lattigo stores them (NTT + Montgomery form over Q||P).
 * limbs[(d*2 + k)*(nQ+nP) + t] = Value[d][k].Coeffs[t].  Only digits / Q-limbs needed up to
 * max_level are read and kept (the level-0 slice of a pack key is 2 MiB of 812 MiB). */
int hec_swk_upload(hec_ctx *ctx, uint64_t galEl, int max_level, const uint64_t *const *limbs);
int hec_swk_drop(hec_ctx *ctx, uint64_t galEl);

/* ---- evaluator ops (the subset of ckks.Evaluator the conv path calls) ------------------ */
/* MulNew(ct, pt)  (conv.go:168,288,527) */
int hec_mul_pt_new(hec_ctx *ctx, const hec_ct *ct, const hec_pt *pt, hec_ct **out);
/* MultByConst(ct, c, ct) */
int hec_mult_by_const(hec_ctx *ctx, hec_ct *ct, double c);
/* Rescale(ct, minScale, ct)  -- error HEC_E_LEVEL at level 0 like the reference */
int hec_rescale(hec_ctx *ctx, hec_ct *ct, double min_scale);
/* SetScale(ct, scale)  (conv.go:528) */
int hec_set_scale(hec_ctx *ctx, hec_ct *ct, double scale);
/* Add(a,b,out) / AddNew / SubNew (conv.go:289-292).  out may alias a or b. */
int hec_add(hec_ctx *ctx, const hec_ct *a, const hec_ct *b, hec_ct *out);
int hec_add_new(hec_ctx *ctx, const hec_ct *a, const hec_ct *b, hec_ct **out);
int hec_sub_new(hec_ctx *ctx, const hec_ct *a, const hec_ct *b, hec_ct **out);
/* Add(ct, pt, ct)  (eval.go:258) */
int hec_add_pt(hec_ctx *ctx, hec_ct *ct, const hec_pt *pt);
/* ---- ct x ct (the multiply of evalReLU's polynomial evaluation, conv.go:435-480)
 * hec_rlk_upload: rlwe.RelinearizationKey{Keys[0]} = one SwitchingKey, same layout as hec_swk_upload
 *   (replaces the Rlk field of rlwe.EvaluationKey passed to ckks.NewEvaluator, main.go:430).
 * hec_mul_relin_new: MulRelinNew(a, b) (mulRelin ciphertext branch, L:ckks/evaluator.go:1398-1444);
 *   a == b is the squaring case.  Error HEC_E_NOKEY without a relinearisation key. */
#define HEC_RLK_ID 0ull /* key-table id of the relinearisation key (never a Galois element: those are odd) */
int hec_rlk_upload(hec_ctx *ctx, int max_level, const uint64_t *const *limbs);
int hec_mul_relin_new(hec_ctx *ctx, const hec_ct *a, const hec_ct *b, hec_ct **out);
/* ---- polynomial evaluation: evalReLU (conv.go:435-480) and the evaluator ops it is made of
 * hec_sub: Sub(a, b, out).  hec_add / hec_sub / *_new follow evaluateInPlace (L:ckks/evaluator.go:365-473):
 *   level = min, scale = max, and when the scales differ by a factor whose floor is > 1 the smaller-scale
 *   operand is first multiplied by that integer.
 * hec_drop_level: DropLevel(ct, levels).  hec_add_const: AddConst(ct, c, ct), real c.
 * hec_mult_by_int_and_add: MultByGaussianIntegerAndAdd(ct, c, 0, out).
 * hec_evaluate_poly: EvaluatePoly(ct, NewPoly(coeffs), target_scale) (L:ckks/polynomial_evaluation.go), real
 *   coefficients, index = degree; eval_scale = params.Scale() (the evaluator's rescale threshold).
 * hec_eval_relu: evalReLU(params, evaluator, ct, alpha); needs the relinearisation key and 10 levels; the result is the un-rescaled product. */
int hec_sub(hec_ctx *ctx, const hec_ct *a, const hec_ct *b, hec_ct *out);
int hec_drop_level(hec_ctx *ctx, hec_ct *ct, int levels);
int hec_mul_by_pow2(hec_ctx *ctx, hec_ct *ct, int pow2); /* MulByPow2(ct, pow2, ct) (eval.go:476) */
/* MultByi(ct, ct) (divide = 0) / DivByi(ct, ct) (divide = 1); Conjugate(ct, out): the key for galEl 2N-1
 * (GaloisElementForRowRotation) must have been uploaded.  Building blocks of CoeffsToSlots / SlotsToCoeffs. */
int hec_mult_by_i(hec_ctx *ctx, hec_ct *ct, int divide);
int hec_conjugate(hec_ctx *ctx, const hec_ct *ct, hec_ct *out);
int hec_add_const(hec_ctx *ctx, hec_ct *ct, double c);
int hec_mult_by_int_and_add(hec_ctx *ctx, const hec_ct *ct, int64_t c, hec_ct *out);
int hec_evaluate_poly(hec_ctx *ctx, const hec_ct *ct, const double *coeffs, int n, double target_scale,
                      double eval_scale, hec_ct **out);
/* EvaluateCheby(ct, cheby, target_scale): coefficients in the Chebyshev basis on [-1, 1] (the sine evaluation of the
 * bootstrapper, L:ckks/bootstrap.go evaluateCheby, calls it after its own change of variable) */
int hec_evaluate_cheby(hec_ctx *ctx, const hec_ct *ct, const double *coeffs, int n, double target_scale,
                       double eval_scale, hec_ct **out);
int hec_eval_relu(hec_ctx *ctx, const hec_ct *ct, double alpha, double eval_scale, hec_ct **out);
/* the loop `for ul: ct_boots[ul] = evalReLU(...)` of eval.go:470-476 as one call: n ciphertexts of a common level and
 * scale, every step one launch sequence for the whole batch */
int hec_eval_relu_many(hec_ctx *ctx, const hec_ct *const *cts, int n, double alpha, double eval_scale, hec_ct **outs);
/* ---- hoisted linear transform: LinearTransform(ct, *PtDiagMatrix) -> MultiplyByDiagMatrixBSGS
 * (L:ckks/linear_transform.go), the core of the bootstrapper's CoeffsToSlots / SlotsToCoeffs.
 * hec_ptdiag_upload takes PtDiagMatrix{LogSlots, N1, Level, Scale, Vec} as the reference's encoder builds it:
 *   limbs[d*(level+1+nP) + t] = Vec[keys[d]][0].Coeffs[t] (t <= level), then Vec[keys[d]][1].Coeffs[..] (the P part),
 *   NTT + Montgomery form; keys[d] in [0, 2^log_slots); n1 a power of two.  Rotation keys for the baby steps
 *   k mod n1 and the giant steps n1*(k / n1) must have been uploaded (HEC_E_NOKEY otherwise). */
typedef struct hec_ptdiag hec_ptdiag;
int hec_ptdiag_upload(hec_ctx *ctx, int log_slots, int n1, int level, double scale, int ndiag, const int *keys,
                      const uint64_t *const *limbs, hec_ptdiag **out);
void hec_ptdiag_free(hec_ctx *ctx, hec_ptdiag *m);
int hec_linear_transform(hec_ctx *ctx, const hec_ct *ct, const hec_ptdiag *m, hec_ct **out);
/* CoeffsToSlots(vec, pDFTInv, eval) / SlotsToCoeffs(ct0, ct1, pDFT, eval) (L:ckks/bootstrap.go; the two halves of the
 * split bootstrapping, BootstrappConv_CtoS / _StoC): chains of LinearTransform + Rescale over the factor matrices,
 * then real / imaginary extraction by conjugation (needs the key of galEl 2N-1).  Under sparse packing
 * (the matrices' LogSlots < LogN-1) CoeffsToSlots returns one ciphertext (*ct1 = NULL; needs the key of rotation 2^LogSlots).
 * hec_sub_sum: Bootstrapper.subSum (rotations by 2^i, i = log_slots .. LogN-2). */
int hec_sub_sum(hec_ctx *ctx, hec_ct *ct, int log_slots);
int hec_coeffs_to_slots(hec_ctx *ctx, const hec_ct *ct, const hec_ptdiag *const *mats, int n, hec_ct **ct0, hec_ct **ct1);
int hec_slots_to_coeffs(hec_ctx *ctx, const hec_ct *ct0, const hec_ct *ct1, const hec_ptdiag *const *mats, int n, hec_ct **out);
/* ---- split bootstrapping, first half: btp.BootstrappConv_CtoS(ct) (eval.go:447-459; the fork's ckks/bootstrap.go).
 * hec_btp_params carries the fields of ckks.Bootstrapper the routine reads (names as in the reference's struct):
 * the Go host builds its Bootstrapper as before and passes them, with the pDFTInv factor matrices uploaded through
 * hec_ptdiag_upload.  Needs the rotation keys of the matrices, the conjugation key and the relinearisation key.
 * Returns the two halves (real / imaginary parts after the sine evaluation) and the constant the reference returns.
 * hec_mod_up: Bootstrapper.modUp (level 0 -> top level, centred lift). */
typedef struct {
    double prescale, postscale, sinescale, sqrt2pi, sc_fac; /* Bootstrapper.prescale .. scFac */
    double message_ratio;                                   /* BootstrappingParameters.MessageRatio */
    double params_scale;                                    /* params.Scale() */
    int sin_type, sin_rescal, arcsine_deg;                  /* SinType (1 = Cos1, 2 = Cos2), SinRescal, ArcSineDeg (must be 0) */
    const uint64_t *sine_qi; int n_sine_qi;                 /* SineEvalModuli.Qi */
    const double *cheby; int n_cheby; double cheby_a, cheby_b; /* sineEvalPoly: Chebyshev coefficients (real), interval */
} hec_btp_params;
int hec_mod_up(hec_ctx *ctx, const hec_ct *ct, hec_ct **out);
int hec_bootstrap_ctos(hec_ctx *ctx, const hec_ct *ct, const hec_btp_params *bp, const hec_ptdiag *const *pdftinv, int nmat,
                       hec_ct **ct0, hec_ct **ct1, double *constant);
/* btp.Bootstrapp(ct) (test_BL.go:133), the un-split bootstrapping of the baseline network: the same head, then
 * SlotsToCoeffs with the pDFT factors */
int hec_bootstrapp(hec_ctx *ctx, const hec_ct *ct, const hec_btp_params *bp, const hec_ptdiag *const *pdftinv, int ninv,
                   const hec_ptdiag *const *pdft, int nfwd, hec_ct **out);
/* btp.BootstrappConv_StoC(ct0, ct1) (eval.go:540-560) = SlotsToCoeffs(ct0, ct1, btp.pDFT, evaluator); ct1 may be NULL.
 * The caller rescales afterwards (eval.go:562: hec_rescale). */
int hec_bootstrap_stoc(hec_ctx *ctx, const hec_ct *ct0, const hec_ct *ct1, const hec_ptdiag *const *pdft, int nmat, hec_ct **out);
/* RotateGal(ct, galEl, out)  (conv.go:291); out may alias ct */
int hec_rotate_gal(hec_ctx *ctx, const hec_ct *ct, uint64_t galEl, hec_ct *out);
/* RotateNew(ct, k)  (eval.go:123) */
int hec_rotate_new(hec_ctx *ctx, const hec_ct *ct, int k, hec_ct **out);
/* RotateHoisted(ct, rotations) (conv.go:133): outs[i] <- rotation by rots[i] */
int hec_rotate_hoisted(hec_ctx *ctx, const hec_ct *ct, const int *rots, int n, hec_ct **outs);
/* GaloisElementForColumnRotationBy(k) */
uint64_t hec_galois_for_rotation(const hec_ctx *ctx, int k);

/* ---- ring-level entry points (parity tests of the kernels themselves) ------------------ */
/* ring: 0 = Q chain, 1 = P chain.  Host in / host out, one limb. ring.NTT / ring.InvNTT */
int hec_ntt(hec_ctx *ctx, int ring, int limb, const uint64_t *in, uint64_t *out, int inverse);
/* KeySwitcher.SwitchKeysInPlace on a host poly c1[0..level] with the key of galEl */
int hec_keyswitch(hec_ctx *ctx, int level, const uint64_t *const *c1, uint64_t galEl,
                  uint64_t *const *d0, uint64_t *const *d1);
/* FastBasisExtender.ModDownSplitNTTPQ: accQ[0..level], accP[0..nP) -> out[0..level] */
int hec_moddown(hec_ctx *ctx, int level, const uint64_t *const *accQ, const uint64_t *const *accP,
                uint64_t *const *out);

/* ---- the conv path --------------------------------------------------------------------- */
#define HEC_CONV_FUSED 0   /* fused kernels (the measured path) */
#define HEC_CONV_OPLEVEL 1 /* replay conv.go:522-546 / 266-300 op by op through the evaluator ops */
/* conv_then_pack (conv.go:522-546) [+ Add(pl_bn_b) of evalConv_BN, eval.go:258, when pt_bias != NULL].
 * pt_ker[max_ob] at level ECD_LV, pt_idx[logN] (gen_idxNlogs, conv.go:241-261), keys for galEl
 * 2^j+1 must have been uploaded.  Returns HEC_E_SCALE on the reference's consistency panic.  The fused path takes
 * max_ob <= 4096 (all packings the reference uses, up to in_wid 4) at a level-1 input; HEC_E_UNSUPPORTED beyond. */
int hec_conv_then_pack(hec_ctx *ctx, const hec_ct *ct_in, const hec_pt *const *pt_ker, int max_ob, int norm,
                       double out_scale, const hec_pt *const *pt_idx, const hec_pt *pt_bias, int flags,
                       hec_ct **out);

/* Baseline (rotation-per-tap) convolution: the timed interval of evalConv_BN_BL_test
 * (eval.go:108-131) = preConv_BL (conv.go:120-143: RotateHoisted by i*in_wid+j, |i|,|j| <= ker_wid/2)
 * + rot_iters x [postConv_BL (conv.go:146-178: sum over taps of MulNew) + RotateNew(i*rot_step) + Add]
 * + Add(pl_bn_b).  pt_taps[i*ker_wid^2 + tap] are the plaintexts postConv_BL encodes on the host
 * (conv.go:165-166), passed pre-encoded.  Keys for all rotations must be uploaded. */
int hec_conv_bl(hec_ctx *ctx, const hec_ct *ct_in, int in_wid, int ker_wid, int rot_iters, int rot_step,
                const hec_pt *const *pt_taps, const hec_pt *pt_bias, hec_ct **out);

/* ---- between-layer helpers, any level / alpha; masks are host-encoded plaintexts.
 * ext_ctxt (conv.go:347-371), each half of bsgs_ctxt (conv.go:303-344) and of ext_double_ctxt
 * (conv.go:374-414): out = sum_i RotateNew(MulNew(input, pts[i]), rots[i]); Rescale(out, min_scale) if
 * do_rescale.  keep_ctxt (conv.go:417-431): MulNew(input, mask); Rescale. */
int hec_ext_ctxt(hec_ctx *ctx, const hec_ct *input, int n, const int *rots, const hec_pt *const *pts,
                 int do_rescale, double min_scale, hec_ct **out);
int hec_keep_ctxt(hec_ctx *ctx, const hec_ct *input, const hec_pt *mask, double min_scale, hec_ct **out);

/* ext_double_ctxt (conv.go:374-414; do_rescale = 1) and bsgs_ctxt (conv.go:303-344; do_rescale = 0): the masked
 * rotate-and-sum twice, mid = sum over (rots_m, pts_m) of the input, out = sum over (rots_r, pts_r) of mid. */
int hec_ext_double_ctxt(hec_ctx *ctx, const hec_ct *input, int n_m, const int *rots_m, const hec_pt *const *pts_m,
                        int n_r, const int *rots_r, const hec_pt *const *pts_r, int do_rescale, double min_scale,
                        hec_ct **out);

/* ---- one layer: evalConv_BNRelu_new (eval.go:272-575) --------------------------------------------------------
 * conv (+BN) on the pack evaluator -> 2^pow relabel -> BootstrappConv_CtoS -> evalReLU + MulByPow2 per half ->
 * ext_ctxt / ext_double_ctxt / keep_ctxt per half -> BootstrappConv_StoC -> Rescale, on the main evaluator.
 * The reference's `kind` string selects which pieces run; here that is expressed by which operands are present:
 *   "Conv", "Conv_sparse", "Conv_inside", "StrConv_inside":  n_conv = 1, move_kind = HEC_MOVE_KEEP
 *      (for the *_inside kinds the caller passes the dilated kernel, eval.go:421-433, as it passes any kernel)
 *   "TransConv":                                        n_conv = 1, move_kind = HEC_MOVE_EXT (r_idx[in_wid][ul])
 *   "StrConv", "StrConv_odd" (+ pt_pre), "StrConv_fast": n_conv = 1, HEC_MOVE_EXT or HEC_MOVE_EXT_DOUBLE (fast_pack)
 *   "StrConv_sparse":       n_conv = 2 (kernels split by output parity, norm/2), pt_shift2 = x^(norm/4), pt_post, EXT_DOUBLE
 *   "StrConv_sparse_full":  n_conv = 1, pt_post, HEC_MOVE_EXT_DOUBLE
 * Everything the Go host computes in floats or with its encoder arrives as handles: kernel / bias plaintexts
 * (prep_Ker, or hec_encode_coeffs_many), the monomials x^offset (EncodeCoeffs of a unit vector), the slot-encoded
 * masks of rot_util.go's index maps (EncodeNTT at the scales conv.go:347-431 prescribes), the bootstrapper. */
#define HEC_MOVE_KEEP 0
#define HEC_MOVE_EXT 1
#define HEC_MOVE_EXT_DOUBLE 2
typedef struct {
    int n_conv;                        /* 1, or 2 for "StrConv_sparse" */
    const hec_pt *const *pt_ker[2];    /* [max_ob] kernel plaintexts of each convolution (level ECD_LV) */
    const hec_pt *pt_bias[2];          /* pl_bn_b of each (level 0, scale out_scale), or NULL */
    int max_ob, norm[2];               /* max_batch; norm of each convolution (norm/2 for the split pair) */
    double out_scale;                  /* 2^round(log2 q0 - (pow + 8))  (eval.go:369) */
    const hec_pt *const *pt_idx;       /* cont.pl_idx */
    int conv_flags;                    /* HEC_CONV_FUSED / HEC_CONV_OPLEVEL */
    const hec_pt *pt_pre, *pt_shift2, *pt_post; /* monomial plaintexts (scale 1), each may be NULL */
    double pow, alpha;                 /* 2^pow scale relabel and MulByPow2; leaky-ReLU slope */
    int iter;                          /* halves to process: 2 for full packing, 1 otherwise */
    const hec_btp_params *btp;         /* the bootstrapper chosen by log_sparse */
    const hec_ptdiag *const *ctos_mats; int n_ctos;
    const hec_ptdiag *const *stoc_mats; int n_stoc;
    int move_kind;                     /* HEC_MOVE_* */
    const hec_pt *keep_mask[2];        /* HEC_MOVE_KEEP: ext_idx mask per half */
    int n_r[2]; const int *rots_r[2]; const hec_pt *const *pts_r[2]; /* r_idx / r_idx_l per half */
    int n_m[2]; const int *rots_m[2]; const hec_pt *const *pts_m[2]; /* m_idx / m_idx_l per half (EXT_DOUBLE) */
    double min_scale;                  /* params.Scale() */
} hec_layer_args;
int hec_conv_bn_relu(hec_ctx *pack_ctx, hec_ctx *ctx, const hec_ct *ct_input, const hec_layer_args *args, hec_ct **out);

/* A prepared evalConv_BN for `batch` independent input ciphertexts per run: kernel
 * plaintexts, monomials, bias and keys stay resident; the kernel sequence is captured in a
 * CUDA graph.  in_level must be 1 (ECD_LV) in this build.  The plan keeps its own copies of the plaintexts
 * (kernels with the SetScale constant folded in, monomials, bias) and references the uploaded keys:
 * hec_swk_drop / hec_swk_upload on a key a live plan reads only release the handle; its device memory goes
 * with the last plan that reads it.  The context must outlive the plan.
 * HEC_E_SCALE when pt_bias is not at out_scale (eval.go:252-257). */
int hec_plan_create(hec_ctx *ctx, const hec_pt *const *pt_ker, int max_ob, int norm, double in_scale,
                    double out_scale, const hec_pt *const *pt_idx, const hec_pt *pt_bias, int batch,
                    hec_plan **plan);
/* device-resident run: outs[i] are allocated on first use if *outs[i] == NULL */
int hec_plan_run(hec_plan *plan, const hec_ct *const *ins, hec_ct **outs);
/* host-buffer run (the end-to-end call): for each ciphertext i, in_c0/in_c1 point to
 * 2 limb pointers (level 1), out_c0/out_c1 to 1 limb pointer (level 0).  Copies run inside. */
int hec_plan_run_host(hec_plan *plan, const uint64_t *const *in_c0, const uint64_t *const *in_c1,
                      uint64_t *const *out_c0, uint64_t *const *out_c1);
/* Pipelined host-buffer runs: submit enqueues H2D copies, the kernel graph and D2H copies on three
 * streams and returns; wait blocks until that batch's outputs are in the caller's buffers.  Two
 * batches may be in flight, so the copies of one batch overlap the kernels of its neighbours.
 * Host buffers should be pinned and must stay valid until the matching wait returns. */
int hec_plan_submit_host(hec_plan *plan, const uint64_t *const *in_c0, const uint64_t *const *in_c1,
                         uint64_t *const *out_c0, uint64_t *const *out_c1, int *ticket);
int hec_plan_wait(hec_plan *plan, int ticket);
/* device-side elapsed time of a pipelined sequence (first H2D enqueued .. last D2H finished) */
int hec_plan_span_begin(hec_plan *plan);
int hec_plan_span_end_ms(hec_plan *plan, float *ms);
/* per-launch device times (ms) of one run with the kernels launched one by one, in launch
 * order A1,A2,A3,(B1..B5) per pack level -- measurement aid for the roofline report */
int hec_plan_profile(hec_plan *plan, const hec_ct *const *ins, float *ms, int cap, int *n);
/* names of those launches, comma separated, in order ("A1,A2,A3,B1,...,B5" per pack level; a plan whose forward
 * transforms are deferred -- see below -- has "A1,A2,B1..B5,...,F1,F2"); returns their number */
int hec_plan_kernel_names(const hec_plan *plan, char *buf, int cap);
/* 1 if the plan carries its level-0 polynomials as pairs (U, e), value = U - NTT(e), and runs one forward transform per
 * output polynomial at the end instead of one per rescale / mod-down (same results bit for bit; chosen at creation when
 * max_ob <= 256, batch * max_ob / norm >= 64 and the pl_idx plaintexts are the monomials of conv.go:241-254; HEC_DEFER=0
 * in the environment turns it off, 2 forces it for small batches too; the single-ciphertext plans hec_conv_then_pack
 * builds and caches never defer: they may run once, and a deferred plan costs more to set up than one run saves) */
int hec_plan_is_deferred(const hec_plan *plan);
void hec_plan_destroy(hec_plan *plan);
/* hec_conv_then_pack (fused) keeps the plans it builds, keyed by its arguments' identities (plaintext and key
 * uploads, max_ob, norm, scales): a per-convolution caller (conv.go:522, one call per image with the layer's
 * kernels) pays for a plan once.  Number of plans currently cached (at most 8, least recently used goes first). */
int hec_plan_cache_size(const hec_ctx *ctx);
/* Host arithmetic, no device: smallest y < p with uint64(float64(y)/float64(p)) == 1, or ~0 if there is none --
 * the float overflow count of the single-prime exact basis extension (L:ring/ring_basis_extension.go:670-713)
 * as the step function the fused mod-down kernel evaluates. */
uint64_t hec_float_quotient_threshold(uint64_t p);

#ifdef __cplusplus
}
#endif
#endif
