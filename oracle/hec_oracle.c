/*
 * hec_oracle.c -- CPU oracle (TEST INFRASTRUCTURE ONLY; see hec_oracle.h).
 *
 * Scalar, single-threaded restatement of the Lattigo-fork arithmetic that the
 * reference's conv path executes.  Pinned against outputs of
 * the reference's own compiled routines, interpreted from the disassembly of its
 * prebuilt binary (hec_oracle.h; tests/test_ref_vectors.py, test_ref_eval_vectors.py).
 * OpenMP is used only in orc_conv_then_pack when nthreads > 1 (the "generous"
 * all-cores CPU figure); nthreads == 1 is the reference-faithful single thread.
 */
#include "hec_oracle.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef uint64_t u64;
typedef unsigned __int128 u128;

/* ========================================================================= *
 * B.1 scalar primitives -- L:ring/modular_reduction.go:11-184, R = 2^64
 * ========================================================================= */
static inline u64 mulhi(u64 a, u64 b) { return (u64)(((u128)a * b) >> 64); }

/* MRed: x*y*2^-64 mod q in [0,q)  (L:ring/modular_reduction.go:61-64) */
static inline u64 mred(u64 x, u64 y, u64 q, u64 qinv) {
    u128 p = (u128)x * y;
    u64 H = mulhi((u64)p * qinv, q);
    u64 r = (u64)(p >> 64) - H + q;
    if (r >= q) r -= q;
    return r;
}
/* MRedConstant: same, result in [0,2q)  (:73-75) */
static inline u64 mred_lazy(u64 x, u64 y, u64 q, u64 qinv) {
    u128 p = (u128)x * y;
    u64 H = mulhi((u64)p * qinv, q);
    return (u64)(p >> 64) - H + q;
}
/* BRedAdd: a mod q for a below a few q  (:93-96) */
static inline u64 bred_add(u64 a, u64 q, u64 bhi) {
    u64 r = a - mulhi(a, bhi) * q;
    if (r >= q) r -= q;
    return r;
}
/* MForm: a*2^64 mod q  (:11-14) */
static inline u64 mform(u64 a, u64 q, u64 bhi, u64 blo) {
    u64 r = (u64)(-(a * bhi + mulhi(a, blo)) * q);
    if (r >= q) r -= q;
    return r;
}
static inline u64 cred(u64 a, u64 q) { return a >= q ? a - q : a; }

static u64 mulmod(u64 a, u64 b, u64 q) { return (u64)(((u128)a * b) % q); }
static u64 modexp(u64 b, u64 e, u64 q) {
    u64 r = 1;
    b %= q;
    while (e) {
        if (e & 1) r = mulmod(r, b, q);
        b = mulmod(b, b, q);
        e >>= 1;
    }
    return r;
}
static u64 bitrev(u64 x, int bits) {
    u64 r = 0;
    for (int i = 0; i < bits; i++) r |= ((x >> i) & 1) << (bits - 1 - i);
    return r;
}

/* ========================================================================= *
 * B.2 ring tables -- L:ring/ring.go:118-200, L:ring/utils.go:69-90
 * ========================================================================= */
typedef struct {
    u64 q, qinv, bhi, blo;
    u64 ninv;      /* MForm(N^-1) */
    u64 gen;       /* primitive root used */
    u64 *psi;      /* NttPsi[brev(j)] = psi^j, Montgomery form */
    u64 *psi_inv;  /* NttPsiInv[brev(j)] = psi^-j */
} ring_mod;

/* smallest g >= 3 that is a primitive root (the loop increments before testing,
 * starting from 2)  -- L:ring/utils.go:69-90 */
static u64 primitive_root(u64 q) {
    u64 fac[64];
    int nf = 0;
    u64 m = q - 1;
    for (u64 p = 2; p * p <= m; p += (p == 2 ? 1 : 2)) {
        if (m % p == 0) {
            fac[nf++] = p;
            while (m % p == 0) m /= p;
        }
    }
    if (m > 1) fac[nf++] = m;
    u64 g = 2;
    for (;;) {
        g++;
        int ok = 1;
        for (int i = 0; i < nf; i++)
            if (modexp(g, (q - 1) / fac[i], q) == 1) { ok = 0; break; }
        if (ok) return g;
    }
}

static void ring_mod_init(ring_mod *r, u64 q, int logN) {
    u64 N = 1ull << logN;
    r->q = q;
    u64 inv = 1; /* Newton: q*inv = 1 mod 2^64 */
    for (int i = 0; i < 7; i++) inv *= 2 - q * inv;
    r->qinv = inv;
    u128 b = (~(u128)0) / q; /* floor(2^128/q) (q odd > 1 never divides 2^128) */
    r->bhi = (u64)(b >> 64);
    r->blo = (u64)b;
    r->ninv = mform(modexp(N, q - 2, q), q, r->bhi, r->blo);
    u64 g = primitive_root(q);
    r->gen = g;
    u64 power = (q - 1) / (N << 1);
    u64 psi_m = mform(modexp(g, power, q), q, r->bhi, r->blo);
    u64 psi_inv_m = mform(modexp(g, (q - 1) - power, q), q, r->bhi, r->blo);
    r->psi = (u64 *)malloc(N * 8);
    r->psi_inv = (u64 *)malloc(N * 8);
    r->psi[0] = mform(1, q, r->bhi, r->blo);
    r->psi_inv[0] = r->psi[0];
    for (u64 j = 1; j < N; j++) {
        u64 prev = bitrev(j - 1, logN), next = bitrev(j, logN);
        r->psi[next] = mred(r->psi[prev], psi_m, q, r->qinv);
        r->psi_inv[next] = mred(r->psi_inv[prev], psi_inv_m, q, r->qinv);
    }
}

/* ========================================================================= *
 * B.3 NTT -- L:ring/ring_ntt.go:74-626
 * ========================================================================= */
/* NTTLazy: Cooley-Tukey, natural in, bit-reversed out; transient values < 8q, output
 * < 6q for canonical input (no final reduction: [A] test_run 0x4e97c0-0x4eaf80). */
static void ntt_lazy(const ring_mod *r, int N, const u64 *in, u64 *out) {
    const u64 q = r->q, qinv = r->qinv, twoq = 2 * q, fourq = 4 * q;
    int t = N >> 1;
    u64 F = r->psi[1];
    for (int j = 0; j < t; j++) {
        u64 V = mred_lazy(in[j + t], F, q, qinv);
        u64 U = in[j];
        out[j] = U + V;
        out[j + t] = U + twoq - V;
    }
    for (int m = 2; m < N; m <<= 1) {
        /* reduce on every other stage: bits.Len(m) odd */
        int len = 0;
        for (int mm = m; mm; mm >>= 1) len++;
        t >>= 1;
        int reduce = (len & 1) || t == 1; /* [A] test_run 0x4eabee: the t==1 stage always reduces */
        const u64 *psi = r->psi + m;
        if (t >= 8) { /* the reference unrolls these loops x8 (ring_ntt.go) */
            for (int i = 0; i < m; i++) {
                u64 *x = out + ((size_t)i * t << 1), *y = x + t;
                F = psi[i];
                if (reduce) {
                    for (int j = 0; j < t; j++) {
                        u64 U = x[j];
                        if (U >= fourq) U -= fourq;
                        u64 V = mred_lazy(y[j], F, q, qinv);
                        x[j] = U + V; y[j] = U + twoq - V;
                    }
                } else {
                    for (int j = 0; j < t; j++) {
                        u64 U = x[j];
                        u64 V = mred_lazy(y[j], F, q, qinv);
                        x[j] = U + V; y[j] = U + twoq - V;
                    }
                }
            }
        } else {
            /* t = 4, 2, 1: walk the array in strides of 2t with the inner loop fully unrolled */
#define BF(a, b, w) do { u64 U = out[a]; if (reduce && U >= fourq) U -= fourq; \
                         u64 V = mred_lazy(out[b], w, q, qinv); out[a] = U + V; out[b] = U + twoq - V; } while (0)
            if (t == 4) for (int i = 0; i < m; i++) { int b = i << 3; u64 w = psi[i]; BF(b, b + 4, w); BF(b + 1, b + 5, w); BF(b + 2, b + 6, w); BF(b + 3, b + 7, w); }
            else if (t == 2) for (int i = 0; i < m; i++) { int b = i << 2; u64 w = psi[i]; BF(b, b + 2, w); BF(b + 1, b + 3, w); }
            else for (int i = 0; i < m; i++) { int b = i << 1; BF(b, b + 1, psi[i]); }
#undef BF
        }
    }
}
/* the lazy output is < 8q; NTT (non lazy) finishes with BRedAdd (ring_ntt.go:82-97).
 * NTTLazy callers in Lattigo consume the lazy value through MRed; we hand them the
 * same residue class reduced to [0,2q) like the original (bound asserted). */
static void ntt_full(const ring_mod *r, int N, const u64 *in, u64 *out) {
    ntt_lazy(r, N, in, out);
    for (int j = 0; j < N; j++) out[j] = bred_add(out[j], r->q, r->bhi);
}

/* InvNTT: Gentleman-Sande; final pass MRed(x, NInv) (canonical) or lazy [0,2q). */
static void intt_core(const ring_mod *r, int N, const u64 *in, u64 *out, int lazy) {
    const u64 q = r->q, qinv = r->qinv, twoq = 2 * q, fourq = 4 * q;
    int t = 1;
    int h = N >> 1;
    for (int i = 0; i < h; i++) {
        u64 F = r->psi_inv[h + i];
        u64 U = in[2 * i], V = in[2 * i + 1];
        u64 X = U + V;
        if (X >= twoq) X -= twoq;
        out[2 * i] = X;
        out[2 * i + 1] = mred_lazy(U + fourq - V, F, q, qinv);
    }
    t <<= 1;
    for (int m = N >> 1; m > 1; m >>= 1) {
        h = m >> 1;
        const u64 *psi = r->psi_inv + h;
#define IBF(a, b, w) do { u64 U = out[a], V = out[b]; u64 X = U + V; if (X >= twoq) X -= twoq; \
                          out[a] = X; out[b] = mred_lazy(U + fourq - V, w, q, qinv); } while (0)
        if (t == 2) for (int i = 0; i < h; i++) { int b = i << 2; u64 w = psi[i]; IBF(b, b + 2, w); IBF(b + 1, b + 3, w); }
        else if (t == 4) for (int i = 0; i < h; i++) { int b = i << 3; u64 w = psi[i]; IBF(b, b + 4, w); IBF(b + 1, b + 5, w); IBF(b + 2, b + 6, w); IBF(b + 3, b + 7, w); }
        else {
            for (int i = 0; i < h; i++) {
                u64 *x = out + ((size_t)i * t << 1), *y = x + t;
                u64 F = psi[i];
                for (int j = 0; j < t; j++) {
                    u64 U = x[j], V = y[j];
                    u64 X = U + V;
                    if (X >= twoq) X -= twoq;
                    x[j] = X;
                    y[j] = mred_lazy(U + fourq - V, F, q, qinv);
                }
            }
        }
#undef IBF
        t <<= 1;
    }
    if (lazy)
        for (int j = 0; j < N; j++) out[j] = mred_lazy(out[j], r->ninv, q, qinv);
    else
        for (int j = 0; j < N; j++) out[j] = mred(out[j], r->ninv, q, qinv);
}

/* ========================================================================= *
 * context: Q chain, P chain, rescale / mod-down / mod-up tables
 * ========================================================================= */
typedef struct {
    int n;          /* source limbs */
    int nt;         /* targets = nQ + nP (Q first) */
    u64 *qib_mont;  /* [n]    MForm((Qd/q_i)^-1 mod q_i) */
    u64 *qisp_mont; /* [nt][n] MForm(Qd/q_i mod p_t) */
    u64 *qpj_inv;   /* [nt][n+1] -v*Qd mod p_t */
} modup_tab;

struct orc_ctx {
    int logN, N, nQ, nP;
    ring_mod *Q, *P;
    u64 **resc;       /* resc[L][i], i<L: MForm(q_i - q_L^-1 mod q_i)  (RescaleParams) */
    u64 *pinv_mont;   /* [nQ] MForm(P^-1 mod q_i)  (modDownParamsPQ) */
    modup_tab pq;     /* basis P (all nP) -> Q targets  (paramsPQ) */
    int alpha, beta_full;
    int *xalpha;      /* full-level digit sizes */
    modup_tab **dec;  /* dec[d][n_d] for n_d = 2..alpha (index n_d), NULL otherwise */
};

static const ring_mod *target_mod(const orc_ctx *c, int t) { return t < c->nQ ? &c->Q[t] : &c->P[t - c->nQ]; }

/* genModUpParams (L:ring/ring_basis_extension.go:43-145): source basis S (ring_mod*[n]),
 * targets all Q then all P. */
static void modup_tab_init(const orc_ctx *c, modup_tab *T, const ring_mod **S, int n) {
    T->n = n;
    T->nt = c->nQ + c->nP;
    T->qib_mont = (u64 *)calloc(n, 8);
    T->qisp_mont = (u64 *)calloc((size_t)T->nt * n, 8);
    T->qpj_inv = (u64 *)calloc((size_t)T->nt * (n + 1), 8);
    for (int i = 0; i < n; i++) {
        u64 qi = S[i]->q;
        u64 star = 1; /* Qd/q_i mod q_i */
        for (int k = 0; k < n; k++)
            if (k != i) star = mulmod(star, S[k]->q % qi, qi);
        u64 barre = modexp(star, qi - 2, qi);
        T->qib_mont[i] = mform(barre, qi, S[i]->bhi, S[i]->blo);
        for (int t = 0; t < T->nt; t++) {
            const ring_mod *pm = target_mod(c, t);
            u64 s = 1;
            for (int k = 0; k < n; k++)
                if (k != i) s = mulmod(s, S[k]->q % pm->q, pm->q);
            T->qisp_mont[(size_t)t * n + i] = mform(s, pm->q, pm->bhi, pm->blo);
        }
    }
    for (int t = 0; t < T->nt; t++) {
        const ring_mod *pm = target_mod(c, t);
        u64 Qmod = 1;
        for (int k = 0; k < n; k++) Qmod = mulmod(Qmod, S[k]->q % pm->q, pm->q);
        u64 v = pm->q - Qmod; /* Qmod != 0 for distinct primes; if target is in S, Qmod==0 -> v=q -> cred */
        u64 *row = T->qpj_inv + (size_t)t * (n + 1);
        row[0] = 0;
        for (int i = 1; i <= n; i++) row[i] = cred(row[i - 1] + v, pm->q);
    }
}
static void modup_tab_free(modup_tab *T) { free(T->qib_mont); free(T->qisp_mont); free(T->qpj_inv); }

orc_ctx *orc_ctx_new(int logN, const uint64_t *Q, int nQ, const uint64_t *P, int nP) {
    orc_ctx *c = (orc_ctx *)calloc(1, sizeof(orc_ctx));
    c->logN = logN; c->N = 1 << logN; c->nQ = nQ; c->nP = nP;
    c->Q = (ring_mod *)calloc(nQ, sizeof(ring_mod));
    c->P = (ring_mod *)calloc(nP > 0 ? nP : 1, sizeof(ring_mod));
    for (int i = 0; i < nQ; i++) ring_mod_init(&c->Q[i], Q[i], logN);
    for (int i = 0; i < nP; i++) ring_mod_init(&c->P[i], P[i], logN);
    /* RescaleParams[L-1][i] = MForm(q_i - (q_L^-1 mod q_i))  (L:ring/ring.go:63-117) */
    c->resc = (u64 **)calloc(nQ, sizeof(u64 *));
    for (int L = 1; L < nQ; L++) {
        c->resc[L] = (u64 *)calloc(L, 8);
        for (int i = 0; i < L; i++) {
            u64 qi = c->Q[i].q;
            u64 inv = modexp(c->Q[L].q % qi, qi - 2, qi);
            c->resc[L][i] = mform(qi - inv, qi, c->Q[i].bhi, c->Q[i].blo);
        }
    }
    if (nP > 0) {
        c->pinv_mont = (u64 *)calloc(nQ, 8);
        for (int i = 0; i < nQ; i++) {
            u64 qi = c->Q[i].q, Pm = 1;
            for (int j = 0; j < nP; j++) Pm = mulmod(Pm, c->P[j].q % qi, qi);
            c->pinv_mont[i] = mform(modexp(Pm, qi - 2, qi), qi, c->Q[i].bhi, c->Q[i].blo);
        }
        const ring_mod *S[64];
        for (int j = 0; j < nP; j++) S[j] = &c->P[j];
        modup_tab_init(c, &c->pq, S, nP);
        /* Decomposer (L:ring/ring_basis_extension.go:482-538) */
        c->alpha = nP;
        c->beta_full = (nQ + nP - 1) / nP;
        c->xalpha = (int *)calloc(c->beta_full, sizeof(int));
        for (int d = 0; d < c->beta_full; d++) c->xalpha[d] = nP;
        if (nQ % nP) c->xalpha[c->beta_full - 1] = nQ % nP;
        c->dec = (modup_tab **)calloc(c->beta_full, sizeof(modup_tab *));
        for (int d = 0; d < c->beta_full; d++) {
            c->dec[d] = (modup_tab *)calloc(nP + 1, sizeof(modup_tab));
            for (int nd = 2; nd <= c->xalpha[d]; nd++) {
                for (int k = 0; k < nd; k++) S[k] = &c->Q[d * nP + k];
                modup_tab_init(c, &c->dec[d][nd], S, nd);
            }
        }
    }
    return c;
}

void orc_ctx_free(orc_ctx *c) {
    if (!c) return;
    for (int i = 0; i < c->nQ; i++) { free(c->Q[i].psi); free(c->Q[i].psi_inv); }
    for (int i = 0; i < c->nP; i++) { free(c->P[i].psi); free(c->P[i].psi_inv); }
    for (int L = 1; L < c->nQ; L++) free(c->resc[L]);
    free(c->resc);
    if (c->nP > 0) {
        free(c->pinv_mont);
        modup_tab_free(&c->pq);
        for (int d = 0; d < c->beta_full; d++) {
            for (int nd = 2; nd <= c->xalpha[d]; nd++) modup_tab_free(&c->dec[d][nd]);
            free(c->dec[d]);
        }
        free(c->dec); free(c->xalpha);
    }
    free(c->Q); free(c->P); free(c);
}
int orc_N(const orc_ctx *c) { return c->N; }
int orc_beta_full(const orc_ctx *c) { return c->beta_full; }
static const ring_mod *rm(const orc_ctx *c, int ring, int limb) { return ring == 0 ? &c->Q[limb] : &c->P[limb]; }
void orc_get_table(const orc_ctx *c, int ring, int limb, int which, uint64_t *out) {
    const ring_mod *r = rm(c, ring, limb);
    memcpy(out, which == 0 ? r->psi : r->psi_inv, (size_t)c->N * 8);
}
uint64_t orc_get_const(const orc_ctx *c, int ring, int limb, int which) {
    const ring_mod *r = rm(c, ring, limb);
    switch (which) {
    case 0: return r->q; case 1: return r->qinv; case 2: return r->bhi;
    case 3: return r->blo; case 4: return r->ninv; default: return r->gen;
    }
}
uint64_t orc_s_mred(uint64_t x, uint64_t y, uint64_t q, uint64_t qinv) { return mred(x, y, q, qinv); }
uint64_t orc_s_mform(uint64_t a, uint64_t q, uint64_t bhi, uint64_t blo) { return mform(a, q, bhi, blo); }
uint64_t orc_s_bred_add(uint64_t a, uint64_t q, uint64_t bhi) { return bred_add(a, q, bhi); }

void orc_ntt(const orc_ctx *c, int ring, int limb, const uint64_t *in, uint64_t *out) {
    ntt_full(rm(c, ring, limb), c->N, in, out);
}
void orc_intt(const orc_ctx *c, int ring, int limb, const uint64_t *in, uint64_t *out) {
    intt_core(rm(c, ring, limb), c->N, in, out, 0);
}

/* ========================================================================= *
 * B.4 automorphism -- L:ring/ring_automorphism.go:31-82, L:rlwe/params.go:308-312
 * ========================================================================= */
void orc_permute_index(int logN, uint64_t galEl, uint32_t *index) {
    u64 N = 1ull << logN, mask = (N << 1) - 1;
    for (u64 i = 0; i < N; i++) {
        u64 t1 = 2 * bitrev(i, logN) + 1;
        u64 t2 = (((galEl * t1) & mask) - 1) >> 1;
        index[i] = (uint32_t)bitrev(t2, logN);
    }
}
uint64_t orc_galois_for_rotation(int logN, int k) {
    u64 twoN = 2ull << logN, mask = twoN - 1;
    u64 e = (u64)(int64_t)k & mask, g = 1, b = 5;
    while (e) { if (e & 1) g = (g * b) & mask; b = (b * b) & mask; e >>= 1; }
    return g;
}
static void permute(const u64 *in, const uint32_t *idx, u64 *out, int N) {
    for (int j = 0; j < N; j++) out[j] = in[idx[j]];
}

/* ========================================================================= *
 * B.5-B.7 evaluator elementwise ops
 * ========================================================================= */
void orc_mul_pt(const orc_ctx *c, int level, const uint64_t *ct0, const uint64_t *ct1,
                const uint64_t *pt, uint64_t *o0, uint64_t *o1) {
    int N = c->N;
    for (int i = 0; i <= level; i++) {
        const ring_mod *r = &c->Q[i];
        for (int j = 0; j < N; j++) {
            size_t k = (size_t)i * N + j;
            u64 m = mform(pt[k], r->q, r->bhi, r->blo);      /* MFormLvl */
            o0[k] = mred(m, ct0[k], r->q, r->qinv);          /* MulCoeffsMontgomeryLvl */
            o1[k] = mred(m, ct1[k], r->q, r->qinv);
        }
    }
}

/* scaleUpExact (L:ckks/utils.go:31-57): floor(|c*n| + 0.5) mod q with sign; the
 * big.Float arithmetic there is prec-53 round-to-nearest-even == IEEE double. */
static u64 scale_up_exact(double value, double n, u64 q) {
    int neg = value < 0;
    volatile double x = neg ? -n * value : n * value;
    volatile double y = x + 0.5;
    double fl = floor(y);
    u64 res;
    if (fl < 18446744073709551616.0) res = (u64)fl % q;
    else { /* > 2^64: exact integer in a double; reduce via 128-bit */
        int e; double m = frexp(fl, &e); /* fl = m*2^e, m in [0.5,1) */
        u64 mant = (u64)ldexp(m, 53); int sh = e - 53;
        u64 rr = mant % q;
        for (int i = 0; i < sh; i++) rr = (u64)(((u128)rr * 2) % q);
        res = rr;
    }
    if (neg) res = q - res;
    return res;
}
double orc_const_limbs(const orc_ctx *c, int level, double constant, uint64_t *k) {
    /* getConstAndScale (L:ckks/evaluator.go:508-561), float64 case */
    double scale = 1.0;
    if (constant != 0) {
        double vi = (double)(int64_t)constant;
        if (constant - vi != 0) scale = (double)c->Q[level].q;
    }
    for (int i = 0; i <= level; i++) k[i] = constant != 0 ? scale_up_exact(constant, scale, c->Q[i].q) : 0;
    return scale;
}
void orc_mul_const(const orc_ctx *c, int level, const uint64_t *in, const uint64_t *k, uint64_t *out) {
    int N = c->N;
    for (int i = 0; i <= level; i++) {
        const ring_mod *r = &c->Q[i];
        u64 km = mform(k[i], r->q, r->bhi, r->blo);
        for (int j = 0; j < N; j++) out[(size_t)i * N + j] = mred(in[(size_t)i * N + j], km, r->q, r->qinv);
    }
}
void orc_div_round_last_ntt(const orc_ctx *c, int level, const uint64_t *in, uint64_t *out) {
    int N = c->N;
    const ring_mod *rl = &c->Q[level];
    u64 *t = (u64 *)malloc((size_t)N * 8), *u = (u64 *)malloc((size_t)N * 8);
    intt_core(rl, N, in + (size_t)level * N, t, 0);
    u64 half = (rl->q - 1) >> 1;
    for (int j = 0; j < N; j++) t[j] = cred(t[j] + half, rl->q);
    for (int i = 0; i < level; i++) {
        const ring_mod *r = &c->Q[i];
        u64 hneg = r->q - bred_add(half, r->q, r->bhi);
        for (int j = 0; j < N; j++) u[j] = t[j] + hneg;
        ntt_lazy(r, N, u, u);
        u64 rp = c->resc[level][i];
        for (int j = 0; j < N; j++) {
            /* NTTLazy output is < 8q here; Lattigo feeds it to MRed as u + 2q - p: the
             * product bound x*y < q*2^64 holds (x < 10q < 2^64 for q < 2^61). */
            out[(size_t)i * N + j] = mred(u[j] + 2 * r->q - in[(size_t)i * N + j], rp, r->q, r->qinv);
        }
    }
    free(t); free(u);
}
int orc_rescale_count(const orc_ctx *c, int level, double scale, double min_scale, double *scale_out) {
    int nb = 0;
    while (level - nb > 0 && scale / (double)c->Q[level - nb].q >= min_scale / 2) {
        scale /= (double)c->Q[level - nb].q;
        nb++;
    }
    *scale_out = scale;
    return nb;
}
void orc_add(const orc_ctx *c, int level, const uint64_t *a, const uint64_t *b, uint64_t *o) {
    int N = c->N;
    for (int i = 0; i <= level; i++)
        for (int j = 0; j < N; j++) { size_t k = (size_t)i * N + j; o[k] = cred(a[k] + b[k], c->Q[i].q); }
}
void orc_sub(const orc_ctx *c, int level, const uint64_t *a, const uint64_t *b, uint64_t *o) {
    int N = c->N;
    for (int i = 0; i <= level; i++)
        for (int j = 0; j < N; j++) { size_t k = (size_t)i * N + j; o[k] = cred(a[k] + c->Q[i].q - b[k], c->Q[i].q); }
}

/* ========================================================================= *
 * B.8 basis extension + key switch
 * ========================================================================= */
/* modUpExact / reconstructRNS / multSum (L:ring/ring_basis_extension.go:438-457,670-779).
 * src[i] : n coefficient-domain limbs (any representative < 2^64 with src*qib < q*2^64)
 * dst[t] : written for every t with want[t] != 0; non-canonical like the original
 *          (acc_hi - H + p + qpjInv[v], no final subtraction). */
static void modup_exact(const orc_ctx *c, const modup_tab *T, const ring_mod **S, const u64 *const *src,
                        u64 *const *dst, const char *want, int N) {
    int n = T->n;
    u64 y[64];
    for (int x = 0; x < N; x++) {
        double vi = 0.0;
        for (int i = 0; i < n; i++) {
            y[i] = mred(src[i][x], T->qib_mont[i], S[i]->q, S[i]->qinv);
            volatile double f = (double)y[i] / (double)S[i]->q; /* no FMA / excess precision */
            volatile double s = vi + f;
            vi = s;
        }
        u64 v = (u64)vi;
        for (int t = 0; t < T->nt; t++) {
            if (!want[t]) continue;
            const ring_mod *pm = target_mod(c, t);
            const u64 *w = T->qisp_mont + (size_t)t * n;
            u64 rlo = 0, rhi = 0;
            for (int i = 0; i < n; i++) {
                u128 m = (u128)y[i] * w[i];
                u64 mlo = (u64)m, mhi = (u64)(m >> 64);
                u64 nlo = rlo + mlo;
                rhi += mhi + (nlo < rlo);
                rlo = nlo;
            }
            u64 hhi = mulhi(rlo * pm->qinv, pm->q);
            dst[t][x] = rhi - hhi + pm->q + T->qpj_inv[(size_t)t * (n + 1) + v];
        }
    }
}

/* ModDownSplitNTTPQ (L:ring/ring_basis_extension.go:247-291) */
void orc_moddown(const orc_ctx *c, int level, const uint64_t *accQ, const uint64_t *accP, uint64_t *out) {
    int N = c->N, nP = c->nP;
    u64 *e = (u64 *)malloc((size_t)nP * N * 8);
    u64 *x = (u64 *)malloc((size_t)(level + 1) * N * 8);
    const ring_mod *S[64];
    const u64 *src[64];
    u64 *dst[128] = {0};
    char want[128];
    memset(want, 0, sizeof want);
    for (int j = 0; j < nP; j++) {
        intt_core(&c->P[j], N, accP + (size_t)j * N, e + (size_t)j * N, 1); /* InvNTTLazy */
        S[j] = &c->P[j]; src[j] = e + (size_t)j * N;
    }
    for (int i = 0; i <= level; i++) { dst[i] = x + (size_t)i * N; want[i] = 1; }
    modup_exact(c, &c->pq, S, src, dst, want, N);
    for (int i = 0; i <= level; i++) {
        const ring_mod *r = &c->Q[i];
        u64 *xi = x + (size_t)i * N;
        ntt_lazy(r, N, xi, xi);
        /* [A] test_run 0x4e5049-0x4e5055 and 0x4e4d77-0x4e4da7: params := q_i - modDownParams[i];
         * p2 = MRed(p3 + 2q_i - p1, params) with p3 the NTTLazy output (< 8q, so the sum
         * never wraps for q < 2^61) and p1 canonical. */
        u64 negpinv = r->q - c->pinv_mont[i];
        for (int j = 0; j < N; j++)
            out[(size_t)i * N + j] = mred(xi[j] + 2 * r->q - accQ[(size_t)i * N + j], negpinv, r->q, r->qinv);
    }
    free(e); free(x);
}

/* DecomposeSingleNTT / DecomposeAndSplit (L:rlwe/keyswitch.go:121-141,
 * L:ring/ring_basis_extension.go:543-664): digit d of c (coefficient domain, cinv) ->
 * NTT-domain limbs dQ[0..level], dP[0..nP).  Out-of-digit limbs are NTTLazy outputs
 * (any 64-bit representative); in-digit limbs are the canonical NTT form of c1. */
static void decompose_digit(const orc_ctx *c, int level, int d, const u64 *c1ntt, const u64 *cinv,
                            u64 *dQ, u64 *dP) {
    int N = c->N, nP = c->nP, alpha = c->alpha;
    int st = d * alpha;
    int nd = c->xalpha[d];
    if (st + nd > level + 1) nd = level + 1 - st;
    if (nd == 1) {
        /* plain copy of the single digit limb into every target limb (no centring) */
        for (int i = 0; i <= level; i++) memcpy(dQ + (size_t)i * N, cinv + (size_t)st * N, (size_t)N * 8);
        for (int j = 0; j < nP; j++) memcpy(dP + (size_t)j * N, cinv + (size_t)st * N, (size_t)N * 8);
    } else {
        const modup_tab *T = &c->dec[d][nd];
        const ring_mod *S[64]; const u64 *src[64]; u64 *dst[128] = {0}; char want[128];
        memset(want, 0, sizeof want);
        for (int k = 0; k < nd; k++) { S[k] = &c->Q[st + k]; src[k] = cinv + (size_t)(st + k) * N; }
        for (int i = 0; i <= level; i++) { dst[i] = dQ + (size_t)i * N; want[i] = !(i >= st && i < st + nd); }
        for (int j = 0; j < nP; j++) { dst[c->nQ + j] = dP + (size_t)j * N; want[c->nQ + j] = 1; }
        modup_exact(c, T, S, src, dst, want, N);
    }
    for (int i = 0; i <= level; i++) {
        u64 *p = dQ + (size_t)i * N;
        if (i >= st && i < st + nd) memcpy(p, c1ntt + (size_t)i * N, (size_t)N * 8); /* reuse NTT form */
        else ntt_lazy(&c->Q[i], N, p, p); /* lazy (< 2^64); consumed by MRed only */
    }
    for (int j = 0; j < nP; j++) {
        u64 *p = dP + (size_t)j * N;
        ntt_lazy(&c->P[j], N, p, p);
    }
}
/* Note on decompose_digit's ntt_lazy inputs: the copy path feeds residues of a larger
 * modulus into a smaller one (values < 2^61) and the general path feeds multSum outputs
 * (< 2^64 only if rhi small).  ntt_lazy's first stage computes U + V with U the raw
 * input, so inputs must stay below 2^64 - 2q; both sources satisfy it (rhi < n*q_src). */

/* test hook: DecomposeAndSplit + NTT of digit d, given c1 in the NTT domain ([level+1][N]);
 * dQ [(level+1)][N], dP [nP][N] (NTT domain, any representative) */
void orc_decompose_digit(const orc_ctx *c, int level, int d, const uint64_t *c1ntt, uint64_t *dQ, uint64_t *dP) {
    int N = c->N, L = level + 1;
    u64 *cinv = (u64 *)malloc((size_t)L * N * 8);
    for (int i = 0; i < L; i++) intt_core(&c->Q[i], N, c1ntt + (size_t)i * N, cinv + (size_t)i * N, 0);
    decompose_digit(c, level, d, c1ntt, cinv, dQ, dP);
    free(cinv);
}

/* SwitchKeysInPlaceNoModDown (L:rlwe/keyswitch.go:149-225): the inner product of the decomposition of c1 with
 * the key, left in the basis Q||P -- a0Q, a1Q [(level+1)][N], a0P, a1P [nP][N], canonical residues (the
 * reference's lazy accumulators end with Reduce).  Building block of the hoisted linear transforms. */
void orc_keyswitch_nomoddown(const orc_ctx *c, int level, const uint64_t *c1, const uint64_t *swk,
                             uint64_t *a0Q, uint64_t *a0P, uint64_t *a1Q, uint64_t *a1P) {
    int N = c->N, nQ = c->nQ, nP = c->nP, L = level + 1;
    int beta = (L + c->alpha - 1) / c->alpha;
    size_t polyw = (size_t)(nQ + nP) * N;
    u64 *cinv = (u64 *)malloc((size_t)L * N * 8);
    u64 *dQ = (u64 *)malloc((size_t)L * N * 8), *dP = (u64 *)malloc((size_t)nP * N * 8);
    memset(a0Q, 0, (size_t)L * N * 8); memset(a1Q, 0, (size_t)L * N * 8);
    memset(a0P, 0, (size_t)nP * N * 8); memset(a1P, 0, (size_t)nP * N * 8);
    for (int i = 0; i < L; i++) intt_core(&c->Q[i], N, c1 + (size_t)i * N, cinv + (size_t)i * N, 0);
    for (int d = 0; d < beta; d++) {
        decompose_digit(c, level, d, c1, cinv, dQ, dP);
        const u64 *k0 = swk + (size_t)(2 * d) * polyw, *k1 = swk + (size_t)(2 * d + 1) * polyw;
        /* MulCoeffsMontgomeryConstant[AndAddNoMod] + Reduce: canonical accumulators.
         * We reduce every digit (same residues as the lazy schedule of the original). */
        for (int i = 0; i < L; i++) {
            const ring_mod *r = &c->Q[i];
            for (int j = 0; j < N; j++) {
                size_t k = (size_t)i * N + j;
                a0Q[k] = cred(a0Q[k] + mred(dQ[k], k0[k], r->q, r->qinv), r->q);
                a1Q[k] = cred(a1Q[k] + mred(dQ[k], k1[k], r->q, r->qinv), r->q);
            }
        }
        for (int p = 0; p < nP; p++) {
            const ring_mod *r = &c->P[p];
            for (int j = 0; j < N; j++) {
                size_t k = (size_t)p * N + j, kk = (size_t)(nQ + p) * N + j;
                a0P[k] = cred(a0P[k] + mred(dP[k], k0[kk], r->q, r->qinv), r->q);
                a1P[k] = cred(a1P[k] + mred(dP[k], k1[kk], r->q, r->qinv), r->q);
            }
        }
    }
    free(cinv); free(dQ); free(dP);
}

void orc_keyswitch(const orc_ctx *c, int level, const uint64_t *c1, const uint64_t *swk,
                   uint64_t *d0, uint64_t *d1) {
    int N = c->N, nP = c->nP, L = level + 1;
    u64 *a0Q = (u64 *)malloc((size_t)L * N * 8), *a1Q = (u64 *)malloc((size_t)L * N * 8);
    u64 *a0P = (u64 *)malloc((size_t)nP * N * 8), *a1P = (u64 *)malloc((size_t)nP * N * 8);
    orc_keyswitch_nomoddown(c, level, c1, swk, a0Q, a0P, a1Q, a1P);
    orc_moddown(c, level, a0Q, a0P, d0);
    orc_moddown(c, level, a1Q, a1P, d1);
    free(a0Q); free(a1Q); free(a0P); free(a1P);
}

/* out = (accumulate ? out : 0) + a * b * 2^-64 limb-wise over `nlimbs` limbs of ring 0 (Q) or 1 (P), canonical:
 * MulCoeffsMontgomery[AndAdd] with b in Montgomery form (the diagonals of a PtDiagMatrix are stored that way) */
void orc_poly_mulmont(const orc_ctx *c, int ring, int nlimbs, const uint64_t *a, const uint64_t *b, uint64_t *out, int accumulate) {
    int N = c->N;
    for (int i = 0; i < nlimbs; i++) {
        const ring_mod *r = ring ? &c->P[i] : &c->Q[i];
        for (int j = 0; j < N; j++) {
            size_t k = (size_t)i * N + j;
            u64 v = mred(a[k], b[k], r->q, r->qinv);
            out[k] = accumulate ? cred(out[k] + v, r->q) : v;
        }
    }
}
/* out = a + b limb-wise over ring 0 / 1 */
void orc_poly_add(const orc_ctx *c, int ring, int nlimbs, const uint64_t *a, const uint64_t *b, uint64_t *out) {
    int N = c->N;
    for (int i = 0; i < nlimbs; i++) {
        const ring_mod *r = ring ? &c->P[i] : &c->Q[i];
        for (int j = 0; j < N; j++) { size_t k = (size_t)i * N + j; out[k] = cred(a[k] + b[k], r->q); }
    }
}

/* Encoder.EncodeCoeffs = ckks.scaleUpVecExact (L:ckks/utils.go:59-123) as the pinned binary computes it (disassembly at
 * 0x532380): n*|v| > 2^64 -> X = trunc(double(n*|v|) + 0.5) as an exact integer (big.Float at 53 bits is an IEEE
 * double; Int truncates), else X = uint64(n*|v| + 0.5) with Go's amd64 conversion (CVTTSD2SI, split at 2^63);
 * out = X mod q for v >= 0 (and -0.0), q - (X mod q) for v < 0 -- q itself when the residue is 0; coefficients
 * past n are cleared.  out: [level+1][N], coefficient domain. */
static u64 go_cvttsd2si(double y) {
    if (!(y < 9223372036854775808.0) || y < -9223372036854775808.0) return 1ull << 63;
    return (u64)(int64_t)y;
}
static u64 go_f2u(double x) {
    return x < 9223372036854775808.0 ? go_cvttsd2si(x) : (go_cvttsd2si(x - 9223372036854775808.0) | (1ull << 63));
}
void orc_scale_up_vec_exact(const orc_ctx *c, const double *values, int n, double scale, int level, uint64_t *out) {
    int N = c->N;
    for (int j = 0; j <= level; j++) {
        u64 q = c->Q[j].q;
        u64 *o = out + (size_t)j * N;
        for (int i = 0; i < N; i++) {
            if (i >= n) { o[i] = 0; continue; }
            double v = values[i];
            int neg = v < 0.0;
            volatile double ax = scale * fabs(v); /* volatile: one rounded product, no fused multiply-add */
            volatile double y = ax + 0.5;
            u64 r;
            if (ax > 18446744073709551616.0) {
                u64 bits;
                double yy = y;
                memcpy(&bits, &yy, 8);
                int e = (int)((bits >> 52) & 0x7ff) - 1075; /* y = mant * 2^e, e >= 12 */
                u128 t = ((bits & ((1ull << 52) - 1)) | (1ull << 52)) % q;
                for (int k = 0; k < e; k++) t = (t << 1) % q;
                r = (u64)t;
            } else {
                r = go_f2u(y) % q;
            }
            o[i] = neg ? q - r : r;
        }
    }
}

void orc_rotate_gal(const orc_ctx *c, int level, const uint64_t *ct0, const uint64_t *ct1,
                    uint64_t galEl, const uint64_t *swk, uint64_t *o0, uint64_t *o1) {
    int N = c->N, L = level + 1;
    u64 *d0 = (u64 *)malloc((size_t)L * N * 8), *d1 = (u64 *)malloc((size_t)L * N * 8);
    uint32_t *idx = (uint32_t *)malloc((size_t)N * 4);
    orc_keyswitch(c, level, ct1, swk, d0, d1);
    orc_add(c, level, d0, ct0, d0);
    orc_permute_index(c->logN, galEl, idx);
    for (int i = 0; i < L; i++) {
        permute(d0 + (size_t)i * N, idx, o0 + (size_t)i * N, N);
        permute(d1 + (size_t)i * N, idx, o1 + (size_t)i * N, N);
    }
    free(d0); free(d1); free(idx);
}

/* ========================================================================= *
 * B.10 compositions -- conv.go:266-300, 522-546; eval.go:250-260
 * ========================================================================= */
static double now_s(void) {
    struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

int orc_conv_then_pack(const orc_ctx *c, const uint64_t *ct0, const uint64_t *ct1, double ct_scale,
                       const uint64_t *pt_ker, double pt_scale, int B, int norm, double out_scale,
                       const uint64_t *pt_idx, const uint64_t *const *swk, const uint64_t *pt_bias,
                       uint64_t *o0, uint64_t *o1, double *o_scale, int nthreads,
                       double *t_mult, double *t_pack) {
    int N = c->N;
    size_t N_ = (size_t)N;
    double t0 = now_s();
    u64 **x0 = (u64 **)calloc(B, sizeof(u64 *)), **x1 = (u64 **)calloc(B, sizeof(u64 *));
    double *sc = (double *)calloc(B, sizeof(double));
    int rc = 0;
    (void)nthreads;
    /* Stage A (conv.go:525-531): MulNew + SetScale(out_scale / (B/norm)) */
#pragma omp parallel for schedule(dynamic) num_threads(nthreads) if (nthreads > 1)
    for (int i = 0; i < B; i++) {
        if (i % norm) continue;
        u64 *m0 = (u64 *)malloc(2 * N_ * 8), *m1 = (u64 *)malloc(2 * N_ * 8);
        orc_mul_pt(c, 1, ct0, ct1, pt_ker + (size_t)i * 2 * N_, m0, m1);
        double s = ct_scale * pt_scale;
        double target = out_scale / (double)(B / norm);
        /* SetScale (L:ckks/evaluator.go:1194-1209) */
        u64 k[2];
        double up = orc_const_limbs(c, 1, target / s, k);
        orc_mul_const(c, 1, m0, k, m0);
        orc_mul_const(c, 1, m1, k, m1);
        s *= up;
        double s2;
        int nb = orc_rescale_count(c, 1, s, target, &s2);
        x0[i] = (u64 *)malloc(N_ * 8); x1[i] = (u64 *)malloc(N_ * 8);
        if (nb == 1) {
            orc_div_round_last_ntt(c, 1, m0, x0[i]);
            orc_div_round_last_ntt(c, 1, m1, x1[i]);
        } else {
#pragma omp critical
            rc = -1; /* would leave the ciphertext at level 1: conv.go:541 panics */
            memcpy(x0[i], m0, N_ * 8); memcpy(x1[i], m1, N_ * 8);
        }
        sc[i] = target;
        free(m0); free(m1);
    }
    double t1 = now_s();
    /* pack_ctxts (conv.go:266-300) */
    int real_cnum = B / norm;
    for (int i = 0; i < B; i++) if (i % norm == 0) sc[i] *= (double)real_cnum;
    int step = B / 2, logStep = 0;
    for (int i = step; i > 1; i /= 2) logStep++;
    int j = c->logN - logStep;
    while (step >= norm && step >= 1) {
        u64 g = (1ull << j) + 1;
        const u64 *key = swk[j - 1]; /* key index i <-> galEl 2^(i+1)+1 (conv.go:255) */
        const u64 *mono = pt_idx + (size_t)logStep * N_;
#pragma omp parallel for schedule(dynamic) num_threads(nthreads) if (nthreads > 1)
        for (int i = 0; i < step; i += norm) {
            u64 *t10 = (u64 *)malloc(N_ * 8), *t11 = (u64 *)malloc(N_ * 8);
            u64 *t20 = (u64 *)malloc(N_ * 8), *t21 = (u64 *)malloc(N_ * 8);
            orc_mul_pt(c, 0, x0[i + step], x1[i + step], mono, t10, t11);   /* tmp1 = ct[i+step]*X^step */
            orc_sub(c, 0, x0[i], t10, t20); orc_sub(c, 0, x1[i], t11, t21); /* tmp2 = ct[i]-tmp1 */
            orc_add(c, 0, x0[i], t10, t10); orc_add(c, 0, x1[i], t11, t11); /* tmp1 = ct[i]+tmp1 */
            orc_rotate_gal(c, 0, t20, t21, g, key, t20, t21);               /* tmp2 = sigma_g(tmp2) */
            orc_add(c, 0, t10, t20, x0[i]); orc_add(c, 0, t11, t21, x1[i]);
            free(t10); free(t11); free(t20); free(t21);
        }
        step /= 2; logStep--; j++;
    }
    double res_scale = sc[0];
    if (res_scale != out_scale) rc = -1; /* conv.go:541-543 */
    if (pt_bias) orc_add(c, 0, x0[0], pt_bias, x0[0]); /* eval.go:258 */
    memcpy(o0, x0[0], N_ * 8); memcpy(o1, x1[0], N_ * 8);
    *o_scale = res_scale;
    double t2 = now_s();
    if (t_mult) *t_mult = t1 - t0;
    if (t_pack) *t_pack = t2 - t1;
    for (int i = 0; i < B; i++) { free(x0[i]); free(x1[i]); }
    free(x0); free(x1); free(sc);
    return rc;
}

/* ========================================================================= *
 * B.9 semantic helpers (not on the evaluation path; distributions immaterial)
 * ========================================================================= */
static u64 sm64(u64 *s) {
    u64 z = (*s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static u64 uniform_below(u64 *s, u64 q) {
    u64 mask = ~0ull >> __builtin_clzll(q);
    for (;;) { u64 v = sm64(s) & mask; if (v < q) return v; }
}
static int64_t gauss(u64 *s, double sigma) {
    for (;;) {
        double u1 = ((double)(sm64(s) >> 11) + 1.0) / 9007199254740993.0;
        double u2 = (double)(sm64(s) >> 11) / 9007199254740992.0;
        double g = sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2) * sigma;
        if (fabs(g) <= 6 * sigma) return (int64_t)llround(g);
    }
}
static void small_to_ntt_mont(const ring_mod *r, int N, const int64_t *v, u64 *out, int to_mont) {
    for (int j = 0; j < N; j++) out[j] = v[j] >= 0 ? (u64)v[j] % r->q : r->q - ((u64)(-v[j]) % r->q);
    ntt_full(r, N, out, out);
    if (to_mont) for (int j = 0; j < N; j++) out[j] = mform(out[j], r->q, r->bhi, r->blo);
}
void orc_gen_secret(const orc_ctx *c, uint64_t seed, int h, uint64_t *sQ, uint64_t *sP) {
    int N = c->N;
    int64_t *s = (int64_t *)calloc(N, sizeof(int64_t));
    u64 st = seed;
    for (int k = 0; k < h;) {
        int pos = (int)(sm64(&st) % (u64)N);
        if (s[pos]) continue;
        s[pos] = (sm64(&st) & 1) ? 1 : -1;
        k++;
    }
    for (int i = 0; i < c->nQ; i++) small_to_ntt_mont(&c->Q[i], N, s, sQ + (size_t)i * N, 1);
    for (int i = 0; i < c->nP; i++) small_to_ntt_mont(&c->P[i], N, s, sP + (size_t)i * N, 1);
    free(s);
}
void orc_gen_rotkey(const orc_ctx *c, uint64_t seed, uint64_t galEl, const uint64_t *sQ,
                    const uint64_t *sP, uint64_t *swk) {
    int N = c->N, nQ = c->nQ, nP = c->nP, nT = nQ + nP;
    size_t polyw = (size_t)nT * N;
    u64 st = seed ^ (galEl * 0x9E3779B97F4A7C15ull);
    /* s_out = sigma_{g^-1}(s): permute the NTT-form secret by the index of g^-1 mod 2N */
    u64 twoN = 2ull * N, ginv = 1, b = galEl, e = N - 1; /* ord of unit group divides N; g^(N-1)... */
    /* inverse via exponentiation: the unit group of Z_2N has exponent N/2; g^-1 = g^(N/2-1) */
    e = (u64)N / 2 - 1;
    while (e) { if (e & 1) ginv = (ginv * b) & (twoN - 1); b = (b * b) & (twoN - 1); e >>= 1; }
    uint32_t *idx = (uint32_t *)malloc((size_t)N * 4);
    orc_permute_index(c->logN, ginv, idx);
    u64 *sout = (u64 *)malloc(polyw * 8);
    for (int t = 0; t < nT; t++) {
        const u64 *src = t < nQ ? sQ + (size_t)t * N : sP + (size_t)(t - nQ) * N;
        permute(src, idx, sout + (size_t)t * N, N);
    }
    int64_t *ev = (int64_t *)malloc((size_t)N * sizeof(int64_t));
    u64 *en = (u64 *)malloc((size_t)N * 8);
    for (int d = 0; d < c->beta_full; d++) {
        u64 *k0 = swk + (size_t)(2 * d) * polyw, *k1 = swk + (size_t)(2 * d + 1) * polyw;
        for (int j = 0; j < N; j++) ev[j] = gauss(&st, 3.2);
        for (int t = 0; t < nT; t++) {
            const ring_mod *r = target_mod(c, t);
            const u64 *sin = t < nQ ? sQ + (size_t)t * N : sP + (size_t)(t - nQ) * N;
            small_to_ntt_mont(r, N, ev, en, 1);
            u64 Pm = 0;
            if (t >= d * c->alpha && t < d * c->alpha + c->xalpha[d]) {
                Pm = 1;
                for (int p = 0; p < nP; p++) Pm = mulmod(Pm, c->P[p].q % r->q, r->q);
            }
            for (int j = 0; j < N; j++) {
                size_t k = (size_t)t * N + j;
                u64 a = uniform_below(&st, r->q);
                k1[k] = a;
                u64 as = mred(a, sout[k], r->q, r->qinv);          /* a*s_out in Montgomery form */
                u64 v = cred(en[j] + r->q - as, r->q);
                if (Pm) v = cred(v + mulmod(Pm, sin[j], r->q), r->q); /* + P*s_in (Montgomery) */
                k0[k] = v;
            }
        }
    }
    free(idx); free(sout); free(ev); free(en);
}
void orc_encrypt(const orc_ctx *c, uint64_t seed, int level, const uint64_t *m,
                 const uint64_t *sQ, uint64_t *ct0, uint64_t *ct1) {
    int N = c->N;
    u64 st = seed;
    int64_t *ev = (int64_t *)malloc((size_t)N * sizeof(int64_t));
    u64 *tmp = (u64 *)malloc((size_t)N * 8);
    for (int j = 0; j < N; j++) ev[j] = gauss(&st, 3.2);
    for (int i = 0; i <= level; i++) {
        const ring_mod *r = &c->Q[i];
        for (int j = 0; j < N; j++) {
            u64 e = ev[j] >= 0 ? (u64)ev[j] : r->q - (u64)(-ev[j]);
            tmp[j] = cred(m[(size_t)i * N + j] + e, r->q);
        }
        ntt_full(r, N, tmp, tmp);
        for (int j = 0; j < N; j++) {
            size_t k = (size_t)i * N + j;
            u64 a = uniform_below(&st, r->q);
            ct1[k] = a;
            ct0[k] = cred(tmp[j] + r->q - mred(a, sQ[k], r->q, r->qinv), r->q);
        }
    }
    free(ev); free(tmp);
}
void orc_decrypt(const orc_ctx *c, int level, const uint64_t *ct0, const uint64_t *ct1,
                 const uint64_t *sQ, uint64_t *m) {
    int N = c->N;
    for (int i = 0; i <= level; i++) {
        const ring_mod *r = &c->Q[i];
        u64 *o = m + (size_t)i * N;
        for (int j = 0; j < N; j++) {
            size_t k = (size_t)i * N + j;
            o[j] = cred(ct0[k] + mred(ct1[k], sQ[k], r->q, r->qinv), r->q);
        }
        intt_core(r, N, o, o, 0);
    }
}
