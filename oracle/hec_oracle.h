/*
 * hec_oracle.h -- CPU oracle for the CKKS homomorphic-convolution hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (optimal_conv_b200/,
 * libhec.so) may include, link or call this.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs use it, as the checker /
 * the reported CPU baseline.
 *
 * PARITY: PINNED BY OUTPUTS OF THE REFERENCE'S OWN COMPILED CODE.  The reference
 * (dwkim606/optimal_conv) keeps its arithmetic in the un-vendored Go module
 * github.com/dwkim606/test_lattigo (v0.0.0-20220812213541-eb33b0555aaa, pinned only by
 * the build info of the prebuilt /root/reference/test_run); no Go toolchain exists in
 * the build image and the reference holds no ciphertext-level golden vectors
 * (SURVEY.md 8c).  This file therefore restates the Lattigo v2.2/2.3 algorithms as
 * recovered in SURVEY.md Appendix B -- and is pinned against the reference itself:
 * tests/golden/{x86emu,refmachine}.py interpret the compiled routines of test_run from
 * their disassembly (the binary is never executed) on seeded inputs, and the oracle must
 * reproduce every output bit for bit (tests/test_ref_vectors.py,
 * tests/test_ref_eval_vectors.py):
 *   ring   primitiveRoot (35 moduli), NewRing tables, NTT/InvNTT(+Lazy), PermuteNTTIndex,
 *          reconstructRNS+multSum, divRoundByLastModulusNTT, ModDownSplitNTTPQ
 *   rlwe   NewKeySwitcher, SwitchKeysInPlace (alpha = 1, 2, 5; every level)
 *   ckks   NewEvaluator, MulNew, MulRelinNew, Add/Sub (scale matching, every aliasing), Add(ct,pt), Rescale,
 *          MulByPow2, MultByi/DivByi, Conjugate, RotateNew, RotateHoisted, EvaluatePoly, EvaluateCheby,
 *          LinearTransform (MultiplyByDiagMatrixBSGS), CoeffsToSlots, SlotsToCoeffs,
 *          Bootstrapper.modUp / BootstrappConv_CtoS / Bootstrapp
 *   main   conv_then_pack (+ pack_ctxts, + the bias Add of evalConv_BN) incl. its panic, at N = 2^8..2^10 and
 *          for all four golden configs at N = 2^16; preConv_BL; evalReLU
 * (the compositions above the ring / key-switch level live in orc.py, on top of the C routines of this file)
 * Go-runtime services the interpreted code calls (allocation, maps, math/big, prime
 * factorisation of q-1) are supplied in Python and listed in refmachine.py.  In addition:
 * algebraic self-tests, a semantic encrypt -> conv -> decrypt test against a float
 * convolution, and an independent big-integer model.
 *
 * Citation convention: "L:pkg/file.go:a-b" = the Lattigo fork's source location
 * (SURVEY.md citation convention); "conv.go:N" = /root/reference/conv.go.
 */
#ifndef HEC_ORACLE_H
#define HEC_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_ctx orc_ctx;

/* A switching key as Lattigo stores it: [digit][2][nQ+nP][N] words, NTT +
 * Montgomery form (L:rlwe/keys.go SwitchingKey{Value [][2]*ring.Poly}). */

/* ---- context -------------------------------------------------------------- */
orc_ctx *orc_ctx_new(int logN, const uint64_t *Q, int nQ, const uint64_t *P, int nP);
void orc_ctx_free(orc_ctx *c);
int orc_N(const orc_ctx *c);
int orc_beta_full(const orc_ctx *c);                   /* ceil(nQ/alpha) */
/* table export so tests can cross-check constants: which=0 psi, 1 psi_inv; ring 0=Q,1=P */
void orc_get_table(const orc_ctx *c, int ring, int limb, int which, uint64_t *out);
uint64_t orc_get_const(const orc_ctx *c, int ring, int limb, int which); /* 0 q,1 qinv,2 bred_hi,3 bred_lo,4 ninv_mont,5 generator */

/* ---- scalar primitives (for known-answer tests) ---------------------------- */
uint64_t orc_s_mred(uint64_t x, uint64_t y, uint64_t q, uint64_t qinv);
uint64_t orc_s_mform(uint64_t a, uint64_t q, uint64_t bhi, uint64_t blo);
uint64_t orc_s_bred_add(uint64_t a, uint64_t q, uint64_t bhi);

/* ---- ring level (ring 0 = Q chain, 1 = P chain) ---------------------------- */
void orc_ntt(const orc_ctx *c, int ring, int limb, const uint64_t *in, uint64_t *out);   /* canonical out */
void orc_intt(const orc_ctx *c, int ring, int limb, const uint64_t *in, uint64_t *out);  /* canonical out */
void orc_permute_index(int logN, uint64_t galEl, uint32_t *index);
uint64_t orc_galois_for_rotation(int logN, int k);

/* ---- evaluator level.  Polys are contiguous [(level+1)][N] ------------------ */
/* MulNew(ct, pt): L:ckks/evaluator.go:1360-1444 (pt branch) */
void orc_mul_pt(const orc_ctx *c, int level, const uint64_t *ct0, const uint64_t *ct1,
                const uint64_t *pt, uint64_t *o0, uint64_t *o1);
/* MultByConst limb constants: L:ckks/evaluator.go:508-561,782-863; utils.go:31-57.
 * Returns the scale multiplier (1 or q_level) and writes k[i], i<=level. */
double orc_const_limbs(const orc_ctx *c, int level, double constant, uint64_t *k);
void orc_mul_const(const orc_ctx *c, int level, const uint64_t *in, const uint64_t *k, uint64_t *out);
/* divRoundByLastModulusNTT: L:ring/ring_scaling.go:442-513.  in [(level+1)][N] -> out [level][N] */
void orc_div_round_last_ntt(const orc_ctx *c, int level, const uint64_t *in, uint64_t *out);
/* Rescale loop count: L:ckks/evaluator.go:1291-1325 */
int orc_rescale_count(const orc_ctx *c, int level, double scale, double min_scale, double *scale_out);
void orc_add(const orc_ctx *c, int level, const uint64_t *a, const uint64_t *b, uint64_t *o);
void orc_sub(const orc_ctx *c, int level, const uint64_t *a, const uint64_t *b, uint64_t *o);

/* SwitchKeysInPlace: L:rlwe/keyswitch.go:72-225.  c1 [(level+1)][N] NTT; swk as above;
 * d0,d1 [(level+1)][N]. */
void orc_keyswitch(const orc_ctx *c, int level, const uint64_t *c1, const uint64_t *swk,
                   uint64_t *d0, uint64_t *d1);
/* SwitchKeysInPlaceNoModDown: L:rlwe/keyswitch.go:149-225.  a0Q,a1Q [(level+1)][N]; a0P,a1P [nP][N] */
void orc_keyswitch_nomoddown(const orc_ctx *c, int level, const uint64_t *c1, const uint64_t *swk,
                             uint64_t *a0Q, uint64_t *a0P, uint64_t *a1Q, uint64_t *a1P);
/* MulCoeffsMontgomery[AndAdd] / Add over nlimbs limbs of ring 0 (Q) or 1 (P): L:ring/ring_operations.go */
void orc_poly_mulmont(const orc_ctx *c, int ring, int nlimbs, const uint64_t *a, const uint64_t *b, uint64_t *out, int accumulate);
void orc_poly_add(const orc_ctx *c, int ring, int nlimbs, const uint64_t *a, const uint64_t *b, uint64_t *out);
/* Encoder.EncodeCoeffs (ckks.scaleUpVecExact); out [level+1][N], coefficient domain, with the reference's q-for-0 word */
void orc_scale_up_vec_exact(const orc_ctx *c, const double *values, int n, double scale, int level, uint64_t *out);
/* test hook: digit d of DecomposeSingleNTT (L:rlwe/keyswitch.go:121-141) */
void orc_decompose_digit(const orc_ctx *c, int level, int d, const uint64_t *c1ntt, uint64_t *dQ, uint64_t *dP);
/* ModDownSplitNTTPQ: L:ring/ring_basis_extension.go:247-291.  accQ [(level+1)][N],
 * accP [nP][N] (both NTT, canonical) -> out [(level+1)][N] */
void orc_moddown(const orc_ctx *c, int level, const uint64_t *accQ, const uint64_t *accP, uint64_t *out);
/* RotateGal / permuteNTT: L:ckks/evaluator.go:1539-1597 */
void orc_rotate_gal(const orc_ctx *c, int level, const uint64_t *ct0, const uint64_t *ct1,
                    uint64_t galEl, const uint64_t *swk, uint64_t *o0, uint64_t *o1);
/* RotateHoisted for one rotation given the shared decomposition is recomputed inside
 * (bit-identical to the hoisted path: same residues, canonical outputs). */

/* ---- compositions ----------------------------------------------------------- */
/* evalConv_BN hot interval (eval.go:250-260): conv_then_pack (conv.go:522-546) incl.
 * pack_ctxts (conv.go:266-300) + bias add.
 *   ct_in  : level-1 ciphertext, c0/c1 [2][N]
 *   pt_ker : B plaintexts [B][2][N] (NTT, canonical)
 *   pt_idx : logN monomial plaintexts at level 0, [logN][N]
 *   swk[j] : key for galEl 2^(j+1)+1, j=0..logN-1 (pointer may be NULL if unused)
 *   pt_bias: [N] level-0 or NULL
 *   out    : level-0 ciphertext c0/c1 [N]
 * returns 0, or -1 on the reference's "LV or scale ... inconsistent" panic. */
int orc_conv_then_pack(const orc_ctx *c, const uint64_t *ct0, const uint64_t *ct1, double ct_scale,
                       const uint64_t *pt_ker, double pt_scale, int B, int norm, double out_scale,
                       const uint64_t *pt_idx, const uint64_t *const *swk, const uint64_t *pt_bias,
                       uint64_t *o0, uint64_t *o1, double *o_scale, int nthreads,
                       double *t_mult, double *t_pack);

/* ---- semantic helpers (keygen/encrypt/decrypt) -------------- */
/* secret: ternary, Hamming weight h; out sQ [nQ][N], sP [nP][N] NTT+Montgomery */
void orc_gen_secret(const orc_ctx *c, uint64_t seed, int h, uint64_t *sQ, uint64_t *sP);
/* switching key for galEl (s_in = s, s_out = sigma_{galEl^-1}(s)); swk [beta][2][nQ+nP][N] */
void orc_gen_rotkey(const orc_ctx *c, uint64_t seed, uint64_t galEl, const uint64_t *sQ,
                    const uint64_t *sP, uint64_t *swk);
/* encrypt coefficient-domain message m [(level+1)][N] (already reduced per limb) */
void orc_encrypt(const orc_ctx *c, uint64_t seed, int level, const uint64_t *m,
                 const uint64_t *sQ, uint64_t *ct0, uint64_t *ct1);
/* decrypt -> coefficient-domain residues [(level+1)][N] */
void orc_decrypt(const orc_ctx *c, int level, const uint64_t *ct0, const uint64_t *ct1,
                 const uint64_t *sQ, uint64_t *m);

#ifdef __cplusplus
}
#endif
#endif
