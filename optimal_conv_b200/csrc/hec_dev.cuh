// hec_dev.cuh -- device-side modular arithmetic and the two NTT tile shapes (sm_100a).
//
// A limb of N = 2^16 words is viewed as a 256 x 256 matrix (index j = r*256 + c).
//  * "column" kernels own 256 rows x 16 columns and run the 8 stages whose butterfly
//    distance is a multiple of 256 (forward stages m = 1..128, inverse stages t >= 256);
//  * "row" kernels own 16 contiguous 256-word blocks and run the 8 stages inside a block
//    (forward m = 256..32768, inverse t = 1..128).
// A thread keeps 16 coefficients in registers and does 4 stages between exchanges, so a
// 16-stage transform is 2 kernels x (4 + exchange + 4).  The forward transform always ends
// on a row kernel and the inverse always starts on one, which is what lets the fused
// kernels chain  inverse -> pointwise -> forward  without leaving the tile.
//
// Conventions follow the Lattigo fork the reference relies on: psi is the
// (q-1)/2N-th power of the smallest primitive root >= 3, tables are indexed NttPsi[brev(j)] =
// psi^j, natural-order input and bit-reversed output; point-wise products use Montgomery form
// with R = 2^64 exactly like MulCoeffsMontgomery.  How a residue class is represented
// *inside* a transform is ours: twiddles are stored as Shoup pairs (w, floor(w*2^64/q)) --
// the kernels are bound by the integer-multiply pipe (profiles/r01a), and a Shoup product is
// 5 wide + 4 narrow multiplies (approximate quotient) against 10 + 3 for a Montgomery product -- and ranges are
// lazy (forward: no correction at all for q < 2^57, a 4q correction per stage otherwise; inverse: [0,4q)).  Every value that leaves for the caller is canonical, so results are
// bit-identical to the reference's canonical residues.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

typedef unsigned long long u64;
typedef unsigned int u32;

#define HEC_LOGN 16
#define HEC_N 65536
#define HEC_TILE 4096           // words per CTA tile (both shapes)
#define HEC_TILES_PER_LIMB 16
#define HEC_THREADS 256
#define HEC_ROW_PITCH 272       // 256 + one pad word per 16 (bank-conflict-free 16p+k reads)
#ifndef HEC_MINB
#define HEC_MINB 3               // resident CTAs per SM the fused kernels are compiled for (<= 85 registers)
#endif

struct ModC {
    u64 q, qinv;      // q * qinv = 1 mod 2^64
    u64 q2;           // 2q
    u64 rmod;         // R mod q  (mred(x, rmod) = x mod q, canonical)
    u64 ninv_w, ninv_s; // N^-1 mod q and its Shoup companion
    u64 wn_w, wn_s;     // psi_inv[1] * N^-1 (the twiddle of the last inverse stage with the N^-1 pass folded in) and companion
    const ulonglong2 *psi;     // Shoup pairs (w, floor(w * 2^64 / q)) of NttPsi, in consumption order (see fwd4)
    const ulonglong2 *psi_inv; // same for NttPsiInv
    int tight;        // q >= 2^57: the forward transform needs range corrections (see fwd4)
    int small;        // q < 2^31: the transforms run on 32-bit words (fwd4_32 / inv4_32)
    u32 mu;           // floor(2^64 / q) for q > 2^40 (else 0): quotient estimate of canon()
    u32 pad;
};

// ---- scalar primitives -----------------------------------------------------------------
// Cost model of the integer pipe, measured (profiles/r02_ubench_arith.txt, profiles/r01e_pipe_mix.csv): IMAD.WIDE.U32
// occupies the fmaheavy pipe for 4 cycles per warp and sub-partition whatever its addend, a 32-bit IMAD (and the
// IMAD.MOV / IMAD.X / IMAD.IADD forms ptxas likes to emit) for 2, ALU instructions 2 on their own pipe, one
// instruction issues per cycle.  A 64-bit modular product is therefore priced by its number of wide products:
//   exact Shoup     y*w - floor(y*ws/2^64)*q        6 wide + 4 narrow   37.6 clk   (__umul64hi + two low products)
//   lazy Montgomery                                10 wide + 3 narrow   49.0 clk
//   approximate Shoup (below)                       5 wide + 4 narrow   31.9 clk
// HEC_ARITH = 2 (default) uses the approximate form in every butterfly; HEC_ARITH = 1 keeps the exact product there
// (round-1 arithmetic, for A/B runs).  Spelling the exact products without 64-bit addends was measured too and does
// not pay (38.0 / 54.4 clk: the carries cost more ALU instructions than the addend forms cost multiplier cycles).
#ifndef HEC_ARITH
#define HEC_ARITH 2
#endif
// 32 x 32 -> 64 without an addend
__device__ __forceinline__ void mulw(u32 &lo, u32 &hi, u32 a, u32 b) {
    asm("{\n\t.reg .u64 t;\n\tmul.wide.u32 t, %2, %3;\n\tmov.b64 {%0,%1}, t;\n\t}" : "=r"(lo), "=r"(hi) : "r"(a), "r"(b));
}
__device__ __forceinline__ u64 pack64(u32 lo, u32 hi) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "r"(lo), "r"(hi)); return r; }
__device__ __forceinline__ void unpack64(u64 x, u32 &lo, u32 &hi) { asm("mov.b64 {%0,%1}, %2;" : "=r"(lo), "=r"(hi) : "l"(x)); }
// floor(x*y / 2^64) - e, e in {0,1,2}: the low partial product and the carry out of the middle column are dropped
// (3 wide products instead of 4, none with an addend)
__device__ __forceinline__ u64 mulhi64_approx(u64 x, u64 y) {
    u32 x0, x1, y0, y1, l1, h1, a, b, d;
    unpack64(x, x0, x1); unpack64(y, y0, y1);
    mulw(l1, h1, x1, y1); mulw(d, a, x0, y1); mulw(d, b, x1, y0);
    const u64 s = (u64)l1 + a + b;
    return s + ((u64)h1 << 32);
}
// x*y mod 2^64: one wide product, two 32-bit IMADs chained through their addend
__device__ __forceinline__ u64 mullo64(u64 x, u64 y) {
    u32 x0, x1, y0, y1, l, h;
    unpack64(x, x0, x1); unpack64(y, y0, y1);
    mulw(l, h, x0, y0);
    h = x0 * y1 + h;
    h = x1 * y0 + h;
    return pack64(l, h);
}
// x*y*R^-1 mod q, result in (0, 2q).  Needs x*y < q*2^64 (always true for y < q).
__device__ __forceinline__ u64 mred_lazy(u64 x, u64 y, u64 q, u64 qinv) {
    u64 lo = x * y;
    u64 hi = __umul64hi(x, y);
    u64 h = __umul64hi(lo * qinv, q);
    return hi - h + q;
}
// canonical [0,q)
__device__ __forceinline__ u64 mred(u64 x, u64 y, u64 q, u64 qinv) {
    u64 r = mred_lazy(x, y, q, qinv);
    return r >= q ? r - q : r;
}
// conditional subtraction a >= q ? a - q : a as a predicated 64-bit subtract (2 ISETP + 2 predicated
// IADD3 instead of compare + subtract + 2 SEL)
__device__ __forceinline__ u64 cred(u64 a, u64 q) {
    asm("{\n\t.reg .pred p;\n\tsetp.ge.u64 p, %0, %1;\n\t@p sub.u64 %0, %0, %1;\n\t}" : "+l"(a) : "l"(q));
    return a;
}
__device__ __forceinline__ u64 addmod(u64 a, u64 b, u64 q) { return cred(a + b, q); }
__device__ __forceinline__ u64 submod(u64 a, u64 b, u64 q) { return cred(a + q - b, q); }

// Shoup product: y * w mod q in [0,2q) for ANY y < 2^64, given ws = floor(w * 2^64 / q).
__device__ __forceinline__ u64 shoup(u64 y, ulonglong2 w, u64 q) {
    u64 qe = __umul64hi(y, w.y);
    return y * w.x - qe * q;
}
// the same with the approximate quotient: [0, 3q + y*q/2^64), below 4q for any y
__device__ __forceinline__ u64 shoup4(u64 y, ulonglong2 w, u64 q) {
#if HEC_ARITH == 2
    u64 qe = mulhi64_approx(y, w.y);
    return mullo64(y, w.x) - mullo64(qe, q);
#else
    return shoup(y, w, q);
#endif
}
// Cooley-Tukey butterfly: T = Y*w < 4q, X' = X + T, Y' = X - T + 4q.  FIX: X >= 4q is pulled back by 4q first
// (needs X < 8q; with it X', Y' < 8q, which still fits 64 bits for q < 2^61).
template <bool FIX>
__device__ __forceinline__ void ct_bfly(u64 &X, u64 &Y, ulonglong2 w, u64 q, u64 q2) {
    u64 x = X;
    if (FIX) x = cred(x, 2 * q2);
    u64 t = shoup4(Y, w, q);
    X = x + t;
    Y = x - t + 2 * q2;
}
// Gentleman-Sande butterfly, [0,4q) -> [0,4q): X' = X + Y pulled back below 4q, Y' = (X - Y + 4q)*w < 4q
// (the sum stays below 8q < 2^64 for q < 2^61)
__device__ __forceinline__ void gs_bfly(u64 &X, u64 &Y, ulonglong2 w, u64 q, u64 q2) {
    u64 u = X, v = Y;
    X = cred(u + v, 2 * q2);
    Y = shoup4(u - v + 2 * q2, w, q);
}

// Four forward stages on 16 register-resident coefficients.  The coefficient at slot k
// pairs with slot k+d for d = 8,4,2,1; stage d uses twiddles NttPsi[ng*base + gi], ng = 8/d,
// gi = k / (2d), where `base` encodes where the 16 coefficients sit in the limb:
//   column kernel, rows p+16k : base = 1            column kernel, rows 16g+k : base = 16+g
//   row kernel, words p+16k   : base = 256+b        row kernel, words 16p+k   : base = 4096+16b+p
// The device tables hold exactly these 15 twiddles per base, in the order they are consumed
// (slot j = ng-1+gi), so a thread reads t[j*stride]: contiguous per thread for the first three
// layouts, and interleaved over the 16 lanes of a block for the last one (coalesced 256-byte
// reads instead of 16 different cache lines -- profiles/r01_ubench_int_pipe.txt).  Table layout
// (one array of N pairs per modulus and direction, a permutation of NttPsi[1..N-1]):
//   [0,16)       column layout A              t = T + 0,                 stride 1
//   [16,256)     column layout B, [g][15]     t = T + 16 + 15 g,         stride 1
//   [256,4096)   row layout A', [b][15]       t = T + 256 + 15 b,        stride 1
//   [4096,65536) row layout B', [b][15][16]   t = T + 4096 + 240 b + p,  stride 16
// Ranges.  Every twiddle product is < 4q, so a stage adds at most 4q to a value.  TIGHT = false (q < 2^57): no
// correction at all; a 16-stage transform of inputs < 4q stays < 68q < 2^64.  TIGHT = true (q < 2^61, 8q < 2^64):
// every stage first pulls X back below 4q, so values entering and leaving are < 8q (the Y operand never needs it:
// the Shoup product accepts any 64-bit value).
#define HEC_TW_COLA 0
#define HEC_TW_COLB 16
#define HEC_TW_ROWA 256
#define HEC_TW_ROWB 4096
template <bool TIGHT, int STRIDE>
__device__ __forceinline__ void fwd4(u64 (&x)[16], const ulonglong2 *__restrict__ t, u64 q, u64 q2) {
#pragma unroll
    for (int lg = 0; lg < 4; lg++) {
        const int d = 8 >> lg, ng = 1 << lg;
#pragma unroll
        for (int gi = 0; gi < ng; gi++) {
            ulonglong2 w = __ldg(t + (ng - 1 + gi) * STRIDE);
#pragma unroll
            for (int k = 0; k < d; k++) ct_bfly<TIGHT>(x[gi * 2 * d + k], x[gi * 2 * d + k + d], w, q, q2);
        }
    }
}
// Four inverse stages (d = 1,2,4,8), same slots from the psi^-1 table.  In: values < 4q, out: values < 4q.
// DEFER (q < 2^57): the sums X' = X + Y are not pulled back after every stage.  A slot that was last a Y (a product,
// < 4q) s stages ago holds a value < 4q * 2^s -- a compile-time function of slot and stage (gs_bound) -- so the
// difference is offset by that multiple of q and the 16 slots are brought back below 4q once, after the fourth stage
// (15 conditional subtractions instead of 32; none at all where the final N^-1 products follow, inv4_final).
#ifndef HEC_GS_DEFER
#define HEC_GS_DEFER 1
#endif
__host__ __device__ constexpr int gs_bound(int slot, int stages_done) { // in units of q
    int low = slot & ((1 << stages_done) - 1), b = 4 << stages_done;
    for (int h = 0; h < stages_done; h++)
        if (low >> h & 1) b = 4 << (stages_done - 1 - h);
    return b;
}
template <int NSTAGES, bool DEFER, int STRIDE>
__device__ __forceinline__ void gs_stages(u64 (&x)[16], const ulonglong2 *__restrict__ t, u64 q) {
#pragma unroll
    for (int lg = 3; lg > 3 - NSTAGES; lg--) {
        const int d = 8 >> lg, ng = 1 << lg, done = 3 - lg;
#pragma unroll
        for (int gi = 0; gi < ng; gi++) {
            ulonglong2 w = __ldg(t + (ng - 1 + gi) * STRIDE);
#pragma unroll
            for (int k = 0; k < d; k++) {
                const int i = gi * 2 * d + k, j = i + d;
                const u64 u = x[i], v = x[j];
                if (DEFER) {
                    x[i] = u + v;
                    x[j] = shoup4(u - v + (u64)gs_bound(j, done) * q, w, q);
                } else {
                    x[i] = cred(u + v, 4 * q);
                    x[j] = shoup4(u - v + 4 * q, w, q);
                }
            }
        }
    }
}
template <bool DEFER, int STRIDE>
__device__ __forceinline__ void inv4(u64 (&x)[16], const ulonglong2 *__restrict__ t, u64 q, u64 q2) {
    gs_stages<4, DEFER, STRIDE>(x, t, q);
    if (DEFER) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
#pragma unroll
            for (int b = gs_bound(i, 4) / 2; b >= 4; b /= 2) x[i] = cred(x[i], (u64)b * q);
        }
    }
}
// final pass of InvNTT: x * N^-1, canonical
__device__ __forceinline__ u64 inv_final(u64 x, const ModC &M) {
    return cred(shoup(x, make_ulonglong2(M.ninv_w, M.ninv_s), M.q), M.q);
}
// The last four inverse stages of a transform with that final pass folded into the last stage (distance N/2, one
// twiddle w): X' = (X + Y) * N^-1, Y' = (X - Y) * (w * N^-1), both canonical -- two exact products per butterfly
// instead of an approximate one, a range correction and two exact ones.  t: the psi^-1 table in column layout A.
template <bool DEFER>
__device__ __forceinline__ void inv4_final(u64 (&x)[16], const ulonglong2 *__restrict__ t, const ModC &M) {
    const u64 q = M.q;
    gs_stages<3, DEFER, 1>(x, t, q);
    const ulonglong2 n1 = make_ulonglong2(M.ninv_w, M.ninv_s), n2 = make_ulonglong2(M.wn_w, M.wn_s);
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const u64 u = x[k], v = x[k + 8];
        x[k] = cred(shoup(u + v, n1, q), q);
        x[k + 8] = cred(shoup(u - v + (u64)(DEFER ? gs_bound(k + 8, 3) : 4) * q, n2, q), q);
    }
}
// x < 16q -> canonical
__device__ __forceinline__ u64 canon16(u64 x, u64 q) { return cred(cred(cred(cred(x, 8 * q), 4 * q), 2 * q), q); }
// x < 8q -> canonical
__device__ __forceinline__ u64 canon8(u64 x, u64 q) { return cred(cred(cred(x, 4 * q), 2 * q), q); }
// y < 2^61 -> y mod q in [0,2q) for 2^40 < q, mu = floor(2^64/q) < 2^24: the quotient estimate
// floor((y >> 32) * mu / 2^32) is floor(y/q) or one less (one wide product, then t*q)
__device__ __forceinline__ u64 reduce_lazy(u64 y, u64 q, u32 mu) {
    u32 lo, t;
    mulw(lo, t, (u32)(y >> 32), mu);
    return y - mullo64((u64)t, q);
}
// any 64-bit representative -> canonical residue.  For q > 2^40 the quotient is estimated from the high word
// (floor((x >> 32) * mu / 2^32) is floor(x/q), one or two less: two wide products and two conditional subtractions, about
// a third of the Montgomery product by R mod q that serves the small moduli)
__device__ __forceinline__ u64 canon(u64 x, const ModC &M) {
    if (M.mu) {
        u32 lo, t;
        mulw(lo, t, (u32)(x >> 32), M.mu);
        return cred(cred(x - mullo64((u64)t, M.q), M.q2), M.q);
    }
    return mred(x, M.rmod, M.q, M.qinv);
}

// ---- row-kernel tile geometry ------------------------------------------------------------
// 256 threads; thread (p = tid&15, bb = tid>>4) works on block b = 16*tile + bb.
// layout A' : slot k <-> word p + 16k   (global accesses coalesced: 16 lanes x 8 B)
// layout B' : slot k <-> word 16p + k
// The exchange only moves data inside a half-warp's own block, so __syncwarp() suffices.
struct RowGeom {
    u32 p, bb, b;
    u32 gbase;   // word offset of layout-A' slot 0 in the limb: b*256 + p
    u32 sbase;   // smem offset of this block
    __device__ __forceinline__ RowGeom(u32 tile) {
        u32 tid = threadIdx.x;
        p = tid & 15; bb = tid >> 4; b = tile * 16 + bb;
        gbase = b * 256 + p;
        sbase = bb * HEC_ROW_PITCH;
    }
    __device__ __forceinline__ u32 twA() const { return HEC_TW_ROWA + 15 * b; }
    __device__ __forceinline__ u32 twB() const { return HEC_TW_ROWB + 240 * b + p; }
};
__device__ __forceinline__ void row_loadA(u64 (&x)[16], const u64 *__restrict__ g, const RowGeom &G) {
#pragma unroll
    for (int k = 0; k < 16; k++) x[k] = g[G.gbase + 16 * k];
}
__device__ __forceinline__ void row_storeA(const u64 (&x)[16], u64 *__restrict__ g, const RowGeom &G) {
#pragma unroll
    for (int k = 0; k < 16; k++) g[G.gbase + 16 * k] = x[k];
}
__device__ __forceinline__ void row_AtoB(u64 (&x)[16], u64 *sm, const RowGeom &G) {
#pragma unroll
    for (int k = 0; k < 16; k++) sm[G.sbase + G.p + 17 * k] = x[k];
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 16; k++) x[k] = sm[G.sbase + 17 * G.p + k];
    __syncwarp();
}
__device__ __forceinline__ void row_BtoA(u64 (&x)[16], u64 *sm, const RowGeom &G) {
#pragma unroll
    for (int k = 0; k < 16; k++) sm[G.sbase + 17 * G.p + k] = x[k];
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 16; k++) x[k] = sm[G.sbase + G.p + 17 * k];
    __syncwarp();
}
// layout-B' global access: 16 contiguous words per thread, as 8 x 128-bit
__device__ __forceinline__ void row_loadB(u64 (&x)[16], const u64 *__restrict__ g, const RowGeom &G) {
    const ulonglong2 *v = reinterpret_cast<const ulonglong2 *>(g + G.b * 256 + 16 * G.p);
#pragma unroll
    for (int k = 0; k < 8; k++) { ulonglong2 t = __ldg(v + k); x[2 * k] = t.x; x[2 * k + 1] = t.y; }
}
// in: layout A' (values < 4q, or < 8q if tight); out: layout B' (lazy: < 68q, or < 8q if tight)
__device__ __forceinline__ void row_fwd8(u64 (&x)[16], u64 *sm, const RowGeom &G, const ModC &M) {
    if (M.tight) {
        fwd4<true, 1>(x, M.psi + G.twA(), M.q, M.q2);
        row_AtoB(x, sm, G);
        fwd4<true, 16>(x, M.psi + G.twB(), M.q, M.q2);
    } else {
        fwd4<false, 1>(x, M.psi + G.twA(), M.q, M.q2);
        row_AtoB(x, sm, G);
        fwd4<false, 16>(x, M.psi + G.twB(), M.q, M.q2);
    }
}
// in: layout B' (values < 4q); out: layout A' (< 4q)
__device__ __forceinline__ void row_inv8(u64 (&x)[16], u64 *sm, const RowGeom &G, const ModC &M) {
    if (HEC_GS_DEFER && !M.tight) {
        inv4<true, 16>(x, M.psi_inv + G.twB(), M.q, M.q2);
        row_BtoA(x, sm, G);
        inv4<true, 1>(x, M.psi_inv + G.twA(), M.q, M.q2);
    } else {
        inv4<false, 16>(x, M.psi_inv + G.twB(), M.q, M.q2);
        row_BtoA(x, sm, G);
        inv4<false, 1>(x, M.psi_inv + G.twA(), M.q, M.q2);
    }
}

// ---- column-kernel tile geometry ---------------------------------------------------------
// 256 threads; thread (cc = tid&15, pg = tid>>4) works on column c = 16*tile + cc.
// layout A : slot k <-> row pg + 16k ;  layout B : slot k <-> row 16pg + k.
// Both layouts are coalesced in global memory (16 lanes x 8 B along a row).
struct ColGeom {
    u32 cc, pg, c;
    __device__ __forceinline__ ColGeom(u32 tile) {
        u32 tid = threadIdx.x;
        cc = tid & 15; pg = tid >> 4; c = tile * 16 + cc;
    }
    __device__ __forceinline__ u32 gA(int k) const { return (pg + 16 * k) * 256 + c; }
    __device__ __forceinline__ u32 gB(int k) const { return (16 * pg + k) * 256 + c; }
    __device__ __forceinline__ u32 sA(int k) const { return (pg + 16 * k) * 16 + cc; }
    __device__ __forceinline__ u32 sB(int k) const { return (16 * pg + k) * 16 + cc; }
};
__device__ __forceinline__ void col_AtoB(u64 (&x)[16], u64 *sm, const ColGeom &G) {
#pragma unroll
    for (int k = 0; k < 16; k++) sm[G.sA(k)] = x[k];
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; k++) x[k] = sm[G.sB(k)];
    __syncthreads();
}
__device__ __forceinline__ void col_BtoA(u64 (&x)[16], u64 *sm, const ColGeom &G) {
#pragma unroll
    for (int k = 0; k < 16; k++) sm[G.sB(k)] = x[k];
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; k++) x[k] = sm[G.sA(k)];
    __syncthreads();
}
// in: layout A, out: layout B
__device__ __forceinline__ void col_fwd8(u64 (&x)[16], u64 *sm, const ColGeom &G, const ModC &M) {
    if (M.tight) {
        fwd4<true, 1>(x, M.psi + HEC_TW_COLA, M.q, M.q2);
        col_AtoB(x, sm, G);
        fwd4<true, 1>(x, M.psi + HEC_TW_COLB + 15 * G.pg, M.q, M.q2);
    } else {
        fwd4<false, 1>(x, M.psi + HEC_TW_COLA, M.q, M.q2);
        col_AtoB(x, sm, G);
        fwd4<false, 1>(x, M.psi + HEC_TW_COLB + 15 * G.pg, M.q, M.q2);
    }
}
// in: layout B (values < 4q), out: layout A (< 4q)
__device__ __forceinline__ void col_inv8(u64 (&x)[16], u64 *sm, const ColGeom &G, const ModC &M) {
    if (HEC_GS_DEFER && !M.tight) {
        inv4<true, 1>(x, M.psi_inv + HEC_TW_COLB + 15 * G.pg, M.q, M.q2);
        col_BtoA(x, sm, G);
        inv4<true, 1>(x, M.psi_inv + HEC_TW_COLA, M.q, M.q2);
    } else {
        inv4<false, 1>(x, M.psi_inv + HEC_TW_COLB + 15 * G.pg, M.q, M.q2);
        col_BtoA(x, sm, G);
        inv4<false, 1>(x, M.psi_inv + HEC_TW_COLA, M.q, M.q2);
    }
}
// the same, ending the inverse transform: out = canonical coefficients (N^-1 included)
__device__ __forceinline__ void col_inv8_final(u64 (&x)[16], u64 *sm, const ColGeom &G, const ModC &M) {
    if (HEC_GS_DEFER && !M.tight) {
        inv4<true, 1>(x, M.psi_inv + HEC_TW_COLB + 15 * G.pg, M.q, M.q2);
        col_BtoA(x, sm, G);
        inv4_final<true>(x, M.psi_inv + HEC_TW_COLA, M);
    } else {
        inv4<false, 1>(x, M.psi_inv + HEC_TW_COLB + 15 * G.pg, M.q, M.q2);
        col_BtoA(x, sm, G);
        inv4_final<false>(x, M.psi_inv + HEC_TW_COLA, M);
    }
}

// ---- moduli below 2^31 (the eleven ReLU primes of the reference's sets, 30-31 bits) --------------------------------
// The same transforms on 32-bit words: a twiddle product is one high product and two 32-bit IMADs
// (qe = hi32(y * ws32), r = y*w - qe*q in [0,2q) for any y < 2^32, ws32 = floor(w * 2^32 / q) = the high word of the
// 64-bit companion already in the tables) -- 8 multiplier-pipe cycles against 28 for the 64-bit butterfly -- and a
// thread's 16 coefficients take 16 registers.  Values stay in [0,2q) (2q < 2^32, but 4q is not: the sum is decided by
// comparing one operand with 2q minus the other, never formed beyond 2q).
// Limbs are still 64-bit words in memory (the reference's layout); only the arithmetic is narrow.
__device__ __forceinline__ u32 shoup32(u32 y, ulonglong2 w, u32 q) {
    const u32 qe = __umulhi(y, (u32)(w.y >> 32));
    return y * (u32)w.x - qe * q;
}
__device__ __forceinline__ u32 add2q(u32 a, u32 b, u32 q2) { const u32 c = q2 - b; return a >= c ? a - c : a + b; } // [0,2q)^2 -> [0,2q)
__device__ __forceinline__ u32 sub2q(u32 a, u32 b, u32 q2) { const u32 d = a - b; return a >= b ? d : d + q2; }
template <int STRIDE>
__device__ __forceinline__ void fwd4_32(u32 (&x)[16], const ulonglong2 *__restrict__ t, u32 q, u32 q2) {
#pragma unroll
    for (int lg = 0; lg < 4; lg++) {
        const int d = 8 >> lg, ng = 1 << lg;
#pragma unroll
        for (int gi = 0; gi < ng; gi++) {
            const ulonglong2 w = __ldg(t + (ng - 1 + gi) * STRIDE);
#pragma unroll
            for (int k = 0; k < d; k++) {
                const int i = gi * 2 * d + k, j = i + d;
                const u32 X = x[i], T = shoup32(x[j], w, q);
                x[i] = add2q(X, T, q2);
                x[j] = sub2q(X, T, q2);
            }
        }
    }
}
template <int STRIDE>
__device__ __forceinline__ void inv4_32(u32 (&x)[16], const ulonglong2 *__restrict__ t, u32 q, u32 q2) {
#pragma unroll
    for (int lg = 3; lg >= 0; lg--) {
        const int d = 8 >> lg, ng = 1 << lg;
#pragma unroll
        for (int gi = 0; gi < ng; gi++) {
            const ulonglong2 w = __ldg(t + (ng - 1 + gi) * STRIDE);
#pragma unroll
            for (int k = 0; k < d; k++) {
                const int i = gi * 2 * d + k, j = i + d;
                const u32 u = x[i], v = x[j];
                x[i] = add2q(u, v, q2);
                x[j] = shoup32(sub2q(u, v, q2), w, q);
            }
        }
    }
}
// exchanges on a 32-bit view of the same shared-memory buffers.  Row kernels: the 17-word skew of the 64-bit layout
// is conflict-free for 32-bit words as it stands (the two half-warps of a warp sit 272 = 16 mod 32 words apart).
__device__ __forceinline__ void row_AtoB32(u32 (&x)[16], u32 *sm, const RowGeom &G) {
#pragma unroll
    for (int k = 0; k < 16; k++) sm[G.sbase + G.p + 17 * k] = x[k];
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 16; k++) x[k] = sm[G.sbase + 17 * G.p + k];
    __syncwarp();
}
__device__ __forceinline__ void row_BtoA32(u32 (&x)[16], u32 *sm, const RowGeom &G) {
#pragma unroll
    for (int k = 0; k < 16; k++) sm[G.sbase + 17 * G.p + k] = x[k];
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 16; k++) x[k] = sm[G.sbase + G.p + 17 * k];
    __syncwarp();
}
// Column kernels: word index (row*16 + cc) with bit 4 flipped by bit 8, which separates the two half-warps of a warp
// (rows 16*pg + k of an even and an odd pg) that would otherwise meet in the same 16 banks.
__device__ __forceinline__ u32 col_sw32(u32 idx) { return idx ^ ((idx >> 4) & 16u); }
__device__ __forceinline__ void col_AtoB32(u32 (&x)[16], u32 *sm, const ColGeom &G) {
#pragma unroll
    for (int k = 0; k < 16; k++) sm[col_sw32(G.sA(k))] = x[k];
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; k++) x[k] = sm[col_sw32(G.sB(k))];
    __syncthreads();
}
__device__ __forceinline__ void col_BtoA32(u32 (&x)[16], u32 *sm, const ColGeom &G) {
#pragma unroll
    for (int k = 0; k < 16; k++) sm[col_sw32(G.sB(k))] = x[k];
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; k++) x[k] = sm[col_sw32(G.sA(k))];
    __syncthreads();
}
// in: [0,2q), out: [0,2q)
__device__ __forceinline__ void row_fwd8_32(u32 (&x)[16], u32 *sm, const RowGeom &G, const ModC &M) {
    fwd4_32<1>(x, M.psi + G.twA(), (u32)M.q, (u32)M.q2);
    row_AtoB32(x, sm, G);
    fwd4_32<16>(x, M.psi + G.twB(), (u32)M.q, (u32)M.q2);
}
__device__ __forceinline__ void row_inv8_32(u32 (&x)[16], u32 *sm, const RowGeom &G, const ModC &M) {
    inv4_32<16>(x, M.psi_inv + G.twB(), (u32)M.q, (u32)M.q2);
    row_BtoA32(x, sm, G);
    inv4_32<1>(x, M.psi_inv + G.twA(), (u32)M.q, (u32)M.q2);
}
__device__ __forceinline__ void col_fwd8_32(u32 (&x)[16], u32 *sm, const ColGeom &G, const ModC &M) {
    fwd4_32<1>(x, M.psi + HEC_TW_COLA, (u32)M.q, (u32)M.q2);
    col_AtoB32(x, sm, G);
    fwd4_32<1>(x, M.psi + HEC_TW_COLB + 15 * G.pg, (u32)M.q, (u32)M.q2);
}
__device__ __forceinline__ void col_inv8_32(u32 (&x)[16], u32 *sm, const ColGeom &G, const ModC &M) {
    inv4_32<1>(x, M.psi_inv + HEC_TW_COLB + 15 * G.pg, (u32)M.q, (u32)M.q2);
    col_BtoA32(x, sm, G);
    inv4_32<1>(x, M.psi_inv + HEC_TW_COLA, (u32)M.q, (u32)M.q2);
}
__device__ __forceinline__ u32 cred32(u32 a, u32 q) { return a >= q ? a - q : a; }

// PermuteNTTIndex computed on the fly (L:ring/ring_automorphism.go:31-44):
// index_g[i] = brev(((g * (2*brev(i)+1)) mod 2N - 1) / 2)
__device__ __forceinline__ u32 brev16(u32 x) { return __brev(x) >> 16; }
__device__ __forceinline__ u32 perm_index(u32 i, u32 g) {
    u32 t = (g * (2u * brev16(i) + 1u)) & (2u * HEC_N - 1u);
    return brev16((t - 1u) >> 1);
}
