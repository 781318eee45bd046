// hec_kernels.cuh -- all CUDA kernels of libhec (sm_100a).
//
// Three families:
//  (1) generic per-limb kernels (NTT passes, element-wise ops, basis extension) from which
//      the op-level evaluator (any level / alpha / beta) is composed;
//  (2) fused kernels for the reference's conv_then_pack path (conv.go:522-546, 266-300) at
//      level 1 -> 0 with one special prime: 3 kernels for Stage A, 5 per pack-tree level;
//  (3) the same path with the forward transforms deferred (k_def*): level-0 polynomials carried as
//      pairs (U, e), value = U - NTT(e); 2 kernels for Stage A, 5 per level, 2 at the end.  The
//      default for batched plans; shares k_convA1 / k_convB1 / k_convB3 with (2).
// Reference routines each kernel covers are cited at the kernel.
#pragma once
#include <cooperative_groups.h>
#include "hec_dev.cuh"

// =========================================================================================
// (1) generic kernels
// =========================================================================================
// one limb through a transform.  Forward transforms can absorb the point-wise step in front of them and the one behind:
//   flags & 1: the input is a residue of ANOTHER modulus (< 2^64): x = (in mod q) + pro_s0     (EW_REDUCE_ADD)
//   flags & 2: out = (x + 2q - ep_b) * ep_s0 * R^-1 instead of x                              (EW_SUBMUL)
//   flags & 4: ... + ep_add on top of that (the "+ c0" that follows the mod-down of a rotation)  (EW_ADD)
// mid: where the first pass leaves its result (null: out) -- needed when ep_b aliases out
//   scatter_g != 0: the result is written through the automorphism, out[index_g^-1[i]] = x[i] with scatter_g = g^-1 mod 2N
//                   (PermuteNTTWithIndexLvl as scattered stores of the last pass instead of a gather pass of its own)
//   HEC_LJ_PRO2: x = in + ((pro_b mod q) + pro_s0) * pro_s1 * R^-1 (in canonical) -- a second coefficient-domain operand, scaled
//                (the rescale's centred remainder times P, folded into the mod-down's transform: hec_mul_relin_rescale_many)
//   HEC_LJ_ADDS: like ADD with the addend scaled: x += ep_add * ep_s1 * R^-1
struct LimbJob { const u64 *in; u64 *out; int mod; int flags; u64 *mid; const u64 *ep_b; u64 ep_s0; u64 pro_s0; const u64 *ep_add; u32 scatter_g;
                 const u64 *pro_b; u64 pro_s1; u64 ep_s1; };
#define HEC_LJ_PRO 1
#define HEC_LJ_EPI 2
#define HEC_LJ_ADD 4
#define HEC_LJ_PRO2 8
#define HEC_LJ_ADDS 16
__device__ __forceinline__ u64 lj_pro2(const LimbJob &job, u32 n, const ModC &M) {
    // `in` is canonical already (the basis extension's output); pro_b is a residue of ANOTHER modulus
    const u64 r = mred(addmod(canon(job.pro_b[n], M), job.pro_s0, M.q), job.pro_s1, M.q, M.qinv);
    return addmod(job.in[n], r, M.q);
}
// programmatic dependent launch: let the next kernel of the stream be scheduled while this grid drains, and do not touch
// memory before the previous grid has completed (both are no-ops for a kernel launched without the attribute)
#define HEC_PDL_SYNC() asm volatile("griddepcontrol.launch_dependents;\n\tgriddepcontrol.wait;" ::: "memory")
// the two halves apart: everything between them may only read kernel parameters, the static modulus table and job tables
// that were staged before the previous kernel was launched
#define HEC_PDL_TRIGGER() asm volatile("griddepcontrol.launch_dependents;" ::: "memory")
#define HEC_PDL_WAIT() asm volatile("griddepcontrol.wait;" ::: "memory")
#ifndef HEC_GEN_MINB
#define HEC_GEN_MINB 3 // resident CTAs per SM the generic transforms are compiled for (register cap): measured
                       // 3 -> -1.6 % key switch, -2.3 % CtoS against no cap; 4 (64 registers, small spills) -> +0.5 %
#endif

// forward NTT = k_col_fwd then k_row_fwd;  inverse = k_row_inv then k_col_inv
// (ring.NTT / ring.InvNTT, L:ring/ring_ntt.go:74-626).  grid = (16, njobs); the job table lives in device memory
// (content-addressed, hec.cu stage_cached), so one launch takes every limb of a batched operation.
// EXT: the job list uses HEC_LJ_PRO2 / HEC_LJ_ADDS (a separate instantiation: the common one stays as lean as it was)
template <bool EXT>
__global__ void __launch_bounds__(HEC_THREADS, HEC_GEN_MINB) k_col_fwd(const LimbJob *__restrict__ jobs, const ModC *__restrict__ mods) {
    HEC_PDL_TRIGGER();
    __shared__ u64 sm[HEC_TILE];
    HEC_PDL_WAIT();
    const LimbJob job = jobs[blockIdx.y];
    const ModC M = mods[job.mod];
    ColGeom G(blockIdx.x);
    if (M.small) {
        u32 y[16];
        if (EXT && (job.flags & HEC_LJ_PRO2)) {
#pragma unroll
            for (int k = 0; k < 16; k++) y[k] = (u32)lj_pro2(job, G.gA(k), M);
        } else if (job.flags & HEC_LJ_PRO) {
#pragma unroll
            for (int k = 0; k < 16; k++) y[k] = (u32)addmod(canon(job.in[G.gA(k)], M), job.pro_s0, M.q);
        } else {
#pragma unroll
            for (int k = 0; k < 16; k++) y[k] = (u32)job.in[G.gA(k)];
        }
        col_fwd8_32(y, reinterpret_cast<u32 *>(sm), G, M);
#pragma unroll
        for (int k = 0; k < 16; k++) job.out[G.gB(k)] = y[k];
        return;
    }
    u64 x[16];
    if (EXT && (job.flags & HEC_LJ_PRO2)) {
#pragma unroll
        for (int k = 0; k < 16; k++) x[k] = lj_pro2(job, G.gA(k), M);
    } else {
#pragma unroll
        for (int k = 0; k < 16; k++) x[k] = job.in[G.gA(k)];
        if (job.flags & HEC_LJ_PRO) {
#pragma unroll
            for (int k = 0; k < 16; k++) x[k] = addmod(canon(x[k], M), job.pro_s0, M.q);
        }
    }
    col_fwd8(x, sm, G, M);
#pragma unroll
    for (int k = 0; k < 16; k++) job.out[G.gB(k)] = x[k];
}
template <bool EXT>
__global__ void __launch_bounds__(HEC_THREADS, HEC_GEN_MINB) k_row_fwd(const LimbJob *__restrict__ jobs, const ModC *__restrict__ mods) {
    HEC_PDL_TRIGGER();
    __shared__ u64 sm[16 * HEC_ROW_PITCH];
    HEC_PDL_WAIT();
    const LimbJob job = jobs[blockIdx.y];
    const ModC M = mods[job.mod];
    RowGeom G(blockIdx.x);
    u64 x[16];
    if (M.small) {
        u32 y[16];
#pragma unroll
        for (int k = 0; k < 16; k++) y[k] = (u32)job.in[G.gbase + 16 * k];
        row_fwd8_32(y, reinterpret_cast<u32 *>(sm), G, M);
        row_BtoA32(y, reinterpret_cast<u32 *>(sm), G);
#pragma unroll
        for (int k = 0; k < 16; k++) x[k] = cred32(y[k], (u32)M.q);
    } else {
        row_loadA(x, job.in, G);
        row_fwd8(x, sm, G, M);
        row_BtoA(x, sm, G);
#pragma unroll
        for (int k = 0; k < 16; k++) x[k] = canon(x[k], M);
    }
    if (job.flags & HEC_LJ_EPI) {
#pragma unroll
        for (int k = 0; k < 16; k++) x[k] = mred(x[k] + M.q2 - job.ep_b[G.gbase + 16 * k], job.ep_s0, M.q, M.qinv);
        if (job.flags & HEC_LJ_ADD) {
#pragma unroll
            for (int k = 0; k < 16; k++) x[k] = addmod(x[k], job.ep_add[G.gbase + 16 * k], M.q);
        } else if (EXT && (job.flags & HEC_LJ_ADDS)) {
#pragma unroll
            for (int k = 0; k < 16; k++) x[k] = addmod(x[k], mred(job.ep_add[G.gbase + 16 * k], job.ep_s1, M.q, M.qinv), M.q);
        }
    }
    if (job.scatter_g) {
#pragma unroll
        for (int k = 0; k < 16; k++) job.out[perm_index(G.gbase + 16 * k, job.scatter_g)] = x[k];
    } else {
        row_storeA(x, job.out, G);
    }
}
__global__ void __launch_bounds__(HEC_THREADS, HEC_GEN_MINB) k_row_inv(const LimbJob *__restrict__ jobs, const ModC *__restrict__ mods) {
    HEC_PDL_TRIGGER();
    __shared__ u64 sm[16 * HEC_ROW_PITCH];
    HEC_PDL_WAIT();
    const LimbJob job = jobs[blockIdx.y];
    const ModC M = mods[job.mod];
    RowGeom G(blockIdx.x);
    if (M.small) {
        u32 y[16];
#pragma unroll
        for (int k = 0; k < 16; k++) y[k] = (u32)job.in[G.gbase + 16 * k];
        row_AtoB32(y, reinterpret_cast<u32 *>(sm), G);
        row_inv8_32(y, reinterpret_cast<u32 *>(sm), G, M);
#pragma unroll
        for (int k = 0; k < 16; k++) job.out[G.gbase + 16 * k] = y[k];
        return;
    }
    u64 x[16];
    row_loadA(x, job.in, G);
    row_AtoB(x, sm, G);
    row_inv8(x, sm, G, M);
    row_storeA(x, job.out, G);
}
__global__ void __launch_bounds__(HEC_THREADS, HEC_GEN_MINB) k_col_inv(const LimbJob *__restrict__ jobs, const ModC *__restrict__ mods) {
    HEC_PDL_TRIGGER();
    __shared__ u64 sm[HEC_TILE];
    HEC_PDL_WAIT();
    const LimbJob job = jobs[blockIdx.y];
    const ModC M = mods[job.mod];
    ColGeom G(blockIdx.x);
    if (M.small) {
        u32 y[16];
#pragma unroll
        for (int k = 0; k < 16; k++) y[k] = (u32)job.in[G.gB(k)];
        col_inv8_32(y, reinterpret_cast<u32 *>(sm), G, M);
        const ulonglong2 ninv = make_ulonglong2(M.ninv_w, M.ninv_s);
#pragma unroll
        for (int k = 0; k < 16; k++) job.out[G.gA(k)] = cred32(shoup32(y[k], ninv, (u32)M.q), (u32)M.q);
        return;
    }
    u64 x[16];
#pragma unroll
    for (int k = 0; k < 16; k++) x[k] = job.in[G.gB(k)];
    col_inv8_final(x, sm, G, M);
#pragma unroll
    for (int k = 0; k < 16; k++) job.out[G.gA(k)] = x[k];
}

// ---- the forward transform as ONE launch: a limb per thread-block cluster --------------------------------------
// 16 CTAs form a cluster and own one limb.  CTA r runs the column pass on column tile r (stages with distance >= 256), then
// every thread hands its 16 coefficients -- rows 16 pg .. 16 pg + 15 of one column, i.e. one column of row tile pg -- to
// CTA pg by writing them into that CTA's shared memory (distributed shared memory, st.shared::cluster), in the layout the
// row pass loads.  After a cluster barrier CTA r runs the row pass on row tile r out of its own shared memory and stores
// the result.  The limb crosses global memory once per 16 stages instead of once per 8 (k_col_fwd + k_row_fwd: write
// 512 KiB, read it back), and the second launch goes.  Same prologue / epilogue options as the two-launch form.
// The column pass writes straight into a buffer the row pass of ANOTHER CTA reads, so the receive buffer cannot share
// storage with the column pass's exchange buffer: 32 KiB + 34 KiB of shared memory per CTA.
#define HEC_F16_SMEM ((HEC_TILE + 16 * HEC_ROW_PITCH) * sizeof(u64))
__global__ void __launch_bounds__(HEC_THREADS, HEC_GEN_MINB) k_fwd16(const LimbJob *__restrict__ jobs, const ModC *__restrict__ mods) {
    namespace cg = cooperative_groups;
    HEC_PDL_TRIGGER();
    extern __shared__ __align__(128) u64 dsm[];
    u64 *sm = dsm;               // exchange buffer of the column pass
    u64 *rx = dsm + HEC_TILE;    // this CTA's row tile, filled by the 16 CTAs of the cluster; then the row pass's exchange buffer
    cg::cluster_group cl = cg::this_cluster();
    const unsigned r = cl.block_rank();
    HEC_PDL_WAIT();
    const LimbJob job = jobs[blockIdx.y];
    const ModC M = mods[job.mod];
    ColGeom G(r);
    RowGeom R(r);
    cl.sync();                   // every CTA of the cluster is resident: its shared memory may be written from now on
    u64 x[16];
    if (M.small) {
        u32 y[16];
        if (job.flags & HEC_LJ_PRO) {
#pragma unroll
            for (int k = 0; k < 16; k++) y[k] = (u32)addmod(canon(job.in[G.gA(k)], M), job.pro_s0, M.q);
        } else {
#pragma unroll
            for (int k = 0; k < 16; k++) y[k] = (u32)job.in[G.gA(k)];
        }
        col_fwd8_32(y, reinterpret_cast<u32 *>(sm), G, M);
        u32 *dst = reinterpret_cast<u32 *>(cl.map_shared_rank(rx, G.pg));
        const u32 e = 16 * r + G.cc;
#pragma unroll
        for (int k = 0; k < 16; k++) dst[k * HEC_ROW_PITCH + e + (e >> 4)] = y[k];
        cl.sync();
        u32 *rx32 = reinterpret_cast<u32 *>(rx);
#pragma unroll
        for (int k = 0; k < 16; k++) y[k] = rx32[R.sbase + R.p + 17 * k];
        __syncwarp();
        row_fwd8_32(y, rx32, R, M);
        row_BtoA32(y, rx32, R);
#pragma unroll
        for (int k = 0; k < 16; k++) x[k] = cred32(y[k], (u32)M.q);
    } else {
#pragma unroll
        for (int k = 0; k < 16; k++) x[k] = job.in[G.gA(k)];
        if (job.flags & HEC_LJ_PRO) {
#pragma unroll
            for (int k = 0; k < 16; k++) x[k] = addmod(canon(x[k], M), job.pro_s0, M.q);
        }
        col_fwd8(x, sm, G, M);   // layout B: slot k <-> row 16 pg + k of column 16 r + cc
        u64 *dst = cl.map_shared_rank(rx, G.pg);
        const u32 e = 16 * r + G.cc;
#pragma unroll
        for (int k = 0; k < 16; k++) dst[k * HEC_ROW_PITCH + e + (e >> 4)] = x[k];
        cl.sync();
#pragma unroll
        for (int k = 0; k < 16; k++) x[k] = rx[R.sbase + R.p + 17 * k]; // layout A': slot k <-> word p + 16 k of block bb
        __syncwarp();
        row_fwd8(x, rx, R, M);
        row_BtoA(x, rx, R);
#pragma unroll
        for (int k = 0; k < 16; k++) x[k] = canon(x[k], M);
    }
    if (job.flags & HEC_LJ_EPI) {
#pragma unroll
        for (int k = 0; k < 16; k++) x[k] = mred(x[k] + M.q2 - job.ep_b[R.gbase + 16 * k], job.ep_s0, M.q, M.qinv);
        if (job.flags & HEC_LJ_ADD) {
#pragma unroll
            for (int k = 0; k < 16; k++) x[k] = addmod(x[k], job.ep_add[R.gbase + 16 * k], M.q);
        }
    }
    if (job.scatter_g) {
#pragma unroll
        for (int k = 0; k < 16; k++) job.out[perm_index(R.gbase + 16 * k, job.scatter_g)] = x[k];
    } else {
        row_storeA(x, job.out, R);
    }
}

// ---- element-wise ------------------------------------------------------------------------
#define HEC_EWJOBS 48
struct EwJob { const u64 *a; const u64 *b; u64 *out; int mod; u32 g; u64 s0; u64 s1; };
struct EwJobs { EwJob j[HEC_EWJOBS]; };
enum {
    EW_MULMONT = 0, // out = a * b * R^-1        (MulCoeffsMontgomery; b in Montgomery form)
    EW_MULSCALAR,   // out = a * s0 * R^-1        (MultByConst; s0 = MForm(k))
    EW_ADD,         // out = a + b                (AddLvl)
    EW_SUB,         // out = a - b                (SubLvl)
    EW_ADD_MONT,    // out = a + b*R^-1           (Add(ct, pt) with pt stored in Montgomery form)
    EW_TOMONT,      // out = a * s0 * R^-1, s0 = R^2 mod q   (MFormLvl)
    EW_REDUCE_ADD,  // out = (a mod q) + s0       (a < 2^64 from another modulus; rescale/digit lift)
    EW_CENTER,      // out = cred(a + s0)         (rescale: + (q_L-1)/2 mod q_L)
    EW_SUBMUL,      // out = (a + 2q - b) * s0 * R^-1, a < 4q lazy, b canonical  (rescale / mod-down combine)
    EW_MAC,         // out = out + a * b * R^-1   (key inner product, canonical accumulator)
    EW_PERMUTE,     // out[i] = a[index_g[i]]     (PermuteNTTWithIndexLvl)
    EW_COPY,
    EW_MULSCALAR_ADD, // out = out + a * s0 * R^-1  (MultByGaussianIntegerAndAdd with a real integer; s0 = MForm(c mod q))
    EW_MULSCALAR_HALVES, // out = a * (i < N/2 ? s0 : s1) * R^-1  (MultByi / DivByi: +-psi^(N/2) on the two halves)
    EW_CENTER_LIFT // out = (a > s1 ? a - q0 : a) mod q with s0 = q0 mod q, s1 = q0 >> 1  (Bootstrapper.modUp)
};
template <int OP>
__device__ __forceinline__ u64 ew_apply(u64 a, u64 b, u64 o, u32 i, const EwJob &job, u64 q, u64 qinv, u64 rmod) {
    if (OP == EW_MULMONT) return mred(a, b, q, qinv);
    else if (OP == EW_MULSCALAR || OP == EW_TOMONT) return mred(a, job.s0, q, qinv);
    else if (OP == EW_ADD) return addmod(a, b, q);
    else if (OP == EW_SUB) return submod(a, b, q);
    else if (OP == EW_ADD_MONT) return addmod(a, mred(b, 1ull, q, qinv), q);
    else if (OP == EW_REDUCE_ADD) return addmod(mred(a, rmod, q, qinv), job.s0, q);
    else if (OP == EW_CENTER) return cred(a + job.s0, q);
    else if (OP == EW_SUBMUL) return mred(a + 2 * q - b, job.s0, q, qinv);
    else if (OP == EW_MAC) return addmod(o, mred(a, b, q, qinv), q);
    else if (OP == EW_MULSCALAR_ADD) return addmod(o, mred(a, job.s0, q, qinv), q);
    else if (OP == EW_MULSCALAR_HALVES) return mred(a, i < HEC_N / 2 ? job.s0 : job.s1, q, qinv);
    else if (OP == EW_CENTER_LIFT) { u64 r = mred(a, rmod, q, qinv); if (a > job.s1) r = submod(r, job.s0, q); return r; }
    else return a; // EW_PERMUTE (a was gathered), EW_COPY
}
template <int OP>
__global__ void __launch_bounds__(256) k_ew(EwJobs J, const ModC *__restrict__ mods) {
    HEC_PDL_SYNC();
    const EwJob job = J.j[blockIdx.y];
    const u64 q = mods[job.mod].q, qinv = mods[job.mod].qinv, rmod = mods[job.mod].rmod;
    constexpr bool NEEDS_B = OP == EW_MULMONT || OP == EW_ADD || OP == EW_SUB || OP == EW_ADD_MONT || OP == EW_SUBMUL || OP == EW_MAC;
    constexpr bool NEEDS_O = OP == EW_MAC || OP == EW_MULSCALAR_ADD;
    const u32 S = gridDim.x * blockDim.x;
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (4 * S <= HEC_N) {
        // four elements per thread with all their loads issued first (each thread reads and writes the same indices,
        // so in-place operands stay safe)
        for (; i < HEC_N; i += 4 * S) {
            u64 a[4], b[4], o[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const u32 x = i + k * S;
                a[k] = job.a[OP == EW_PERMUTE ? perm_index(x, job.g) : x];
                b[k] = NEEDS_B ? job.b[x] : 0;
                o[k] = NEEDS_O ? job.out[x] : 0;
            }
#pragma unroll
            for (int k = 0; k < 4; k++) job.out[i + k * S] = ew_apply<OP>(a[k], b[k], o[k], i + k * S, job, q, qinv, rmod);
        }
    } else {
        for (; i < HEC_N; i += S) {
            const u64 a = job.a[OP == EW_PERMUTE ? perm_index(i, job.g) : i];
            const u64 b = NEEDS_B ? job.b[i] : 0, o = NEEDS_O ? job.out[i] : 0;
            job.out[i] = ew_apply<OP>(a, b, o, i, job, q, qinv, rmod);
        }
    }
}

// ---- ct x ct tensor product (mulRelin, ciphertext branch, L:ckks/evaluator.go:1398-1432):
// c0 = a0 b0, c1 = a0 b1 + a1 b0, c2 = a1 b1 on canonical NTT-domain residues; MFormLvl(a) then
// MulCoeffsMontgomery(.., b) is the plain modular product.  4 limbs read, 3 written, one pass.
#define HEC_TNJOBS 32
struct TensorJob { const u64 *a0, *a1, *b0, *b1; u64 *c0, *c1, *c2; u64 r2; int mod; };
struct TensorJobs { TensorJob j[HEC_TNJOBS]; };
__global__ void __launch_bounds__(256) k_tensor(TensorJobs J, const ModC *__restrict__ mods) {
    HEC_PDL_SYNC();
    const TensorJob job = J.j[blockIdx.y];
    const u64 q = mods[job.mod].q, qinv = mods[job.mod].qinv;
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < HEC_N; i += gridDim.x * blockDim.x) {
        const u64 A0 = mred(job.a0[i], job.r2, q, qinv), A1 = mred(job.a1[i], job.r2, q, qinv);
        const u64 b0 = job.b0[i], b1 = job.b1[i];
        job.c0[i] = mred(A0, b0, q, qinv);
        job.c1[i] = addmod(mred(A0, b1, q, qinv), mred(A1, b0, q, qinv), q);
        job.c2[i] = mred(A1, b1, q, qinv);
    }
}

// ---- sum over taps: out = sum_t a_t * b_t * R^-1 (b != null: a chain of MulNew + Add, conv.go:168-171)
// or out = sum_t a_t (b == null: a chain of Add, eval.go:123).  Pointer lists live in device memory.
#ifndef HEC_DOT_U
#define HEC_DOT_U 4
#endif
struct DotJob { long long a_off; long long b_off; u64 *out; int mod; int T; }; // offsets into the pointer list; b_off < 0: none
__global__ void __launch_bounds__(256) k_dot(const DotJob *__restrict__ jobs, const u64 *const *__restrict__ ptrs,
                                             const ModC *__restrict__ mods) {
    HEC_PDL_SYNC();
    const DotJob jd = jobs[blockIdx.y];
    struct { const u64 *const *a; const u64 *const *b; u64 *out; int mod; int T; } job =
        {ptrs + jd.a_off, jd.b_off < 0 ? nullptr : ptrs + jd.b_off, jd.out, jd.mod, jd.T};
    const u64 q = mods[job.mod].q, qinv = mods[job.mod].qinv;
    // HEC_DOT_U coefficients per thread and term: 2 * HEC_DOT_U independent loads in flight per thread (the sum is bandwidth-bound and
    // one coefficient at a time leaves too few bytes in flight per SM)
    const u32 S = gridDim.x * blockDim.x;
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < HEC_N; i += HEC_DOT_U * S) {
        u64 acc[HEC_DOT_U] = {};
        for (int t = 0; t < job.T; t++) {
            const u64 *pa = job.a[t];
            u64 v[HEC_DOT_U];
#pragma unroll
            for (int k = 0; k < HEC_DOT_U; k++) v[k] = pa[i + k * S];
            if (job.b != nullptr) {
                const u64 *__restrict__ pb = job.b[t];
                u64 w[HEC_DOT_U];
#pragma unroll
                for (int k = 0; k < HEC_DOT_U; k++) w[k] = __ldg(pb + i + k * S);
#pragma unroll
                for (int k = 0; k < HEC_DOT_U; k++) v[k] = mred(v[k], w[k], q, qinv);
            }
#pragma unroll
            for (int k = 0; k < HEC_DOT_U; k++) acc[k] = addmod(acc[k], v[k], q);
        }
#pragma unroll
        for (int k = 0; k < HEC_DOT_U; k++) job.out[i + k * S] = acc[k];
    }
}

// out = cst + sum_t a_t * s_t * R^-1 with scalars s_t (Montgomery form): the whole linear combination of power-basis
// ciphertexts that EvaluatePoly's leaf forms (evaluatePolyFromPowerBasis: AddConst + one MultByGaussianIntegerAndAdd per
// coefficient) in ONE pass -- each term is read once and the sum written once, instead of a read-modify-write launch per
// term.  table[a_off + t] holds the operand pointers, table[s_off + t] the scalars.
struct LinJob { long long a_off, s_off; u64 *out; u64 cst; int mod; int T; };
__global__ void __launch_bounds__(256) k_lincomb(const LinJob *__restrict__ jobs, const u64 *__restrict__ table, const ModC *__restrict__ mods) {
    HEC_PDL_SYNC();
    const LinJob J = jobs[blockIdx.y];
    const u64 q = mods[J.mod].q, qinv = mods[J.mod].qinv;
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < HEC_N; i += gridDim.x * blockDim.x) {
        u64 acc = J.cst;
        for (int t = 0; t < J.T; t++) {
            const u64 *a = reinterpret_cast<const u64 *>(table[J.a_off + t]);
            acc = addmod(acc, mred(a[i], table[J.s_off + t], q, qinv), q);
        }
        J.out[i] = acc;
    }
}

// ---- the same sums with the operand tiles staged by the TMA engine --------------------------------------------
// k_dot is the one kernel of the key switch that only streams: per output limb it reads 2T limbs (T digits x
// {decomposed polynomial, key}) once and writes one.  Here a CTA owns HEC_DOT_TILE coefficients of an output limb and
// one thread feeds a ring of HEC_DOT_STAGES shared-memory stages with bulk asynchronous copies (cp.async.bulk ->
// UBLKCP), each completing on the stage's mbarrier (SYNCS); all threads consume stage t while the copies of stages
// t+1 .. t+STAGES-1 are in flight -- no registers hold data in flight and no thread issues per-element loads.
#define HEC_DOT_TILE 1024
#define HEC_DOT_STAGES 4
#define HEC_DOT_BULK_SMEM (2 * HEC_DOT_STAGES * HEC_DOT_TILE * sizeof(u64) + HEC_DOT_STAGES * sizeof(u64))
__device__ __forceinline__ u32 smem_u32(const void *p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64 *bar, u32 count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(u64 *bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(u64 *bar, u32 parity) {
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_LOOP:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra WAIT_DONE;\n\t"
                 "bra WAIT_LOOP;\n\tWAIT_DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, u32 bytes, u64 *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__global__ void __launch_bounds__(256) k_dot_bulk(const DotJob *__restrict__ jobs, const u64 *const *__restrict__ ptrs,
                                                  const ModC *__restrict__ mods) {
    HEC_PDL_SYNC();
    extern __shared__ __align__(128) u64 dsm[];
    u64 *sa = dsm, *sb = dsm + HEC_DOT_STAGES * HEC_DOT_TILE, *full = dsm + 2 * HEC_DOT_STAGES * HEC_DOT_TILE;
    const DotJob jd = jobs[blockIdx.y];
    const u64 *const *pa = ptrs + jd.a_off;
    const u64 *const *pb = jd.b_off < 0 ? nullptr : ptrs + jd.b_off;
    const int T = jd.T;
    const u64 q = mods[jd.mod].q, qinv = mods[jd.mod].qinv;
    const u32 base = blockIdx.x * HEC_DOT_TILE, tid = threadIdx.x;
    constexpr u32 BYTES = HEC_DOT_TILE * sizeof(u64);
    if (tid == 0) {
        for (int s = 0; s < HEC_DOT_STAGES; s++) mbar_init(full + s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](int t) { // thread 0: both operand tiles of term t into stage t % STAGES
        const int s = t % HEC_DOT_STAGES;
        mbar_expect_tx(full + s, pb ? 2 * BYTES : BYTES);
        bulk_g2s(sa + s * HEC_DOT_TILE, pa[t] + base, BYTES, full + s);
        if (pb) bulk_g2s(sb + s * HEC_DOT_TILE, pb[t] + base, BYTES, full + s);
    };
    if (tid == 0)
        for (int t = 0; t < HEC_DOT_STAGES && t < T; t++) issue(t);
    u64 acc[4] = {0, 0, 0, 0};
    for (int t = 0; t < T; t++) {
        const int s = t % HEC_DOT_STAGES;
        mbar_wait(full + s, (t / HEC_DOT_STAGES) & 1);
        const ulonglong2 *va = reinterpret_cast<const ulonglong2 *>(sa + s * HEC_DOT_TILE);
        const ulonglong2 a0 = va[tid], a1 = va[256 + tid];
        u64 v[4] = {a0.x, a0.y, a1.x, a1.y};
        if (pb) {
            const ulonglong2 *vb = reinterpret_cast<const ulonglong2 *>(sb + s * HEC_DOT_TILE);
            const ulonglong2 b0 = vb[tid], b1 = vb[256 + tid];
            v[0] = mred(v[0], b0.x, q, qinv); v[1] = mred(v[1], b0.y, q, qinv);
            v[2] = mred(v[2], b1.x, q, qinv); v[3] = mred(v[3], b1.y, q, qinv);
        }
#pragma unroll
        for (int k = 0; k < 4; k++) acc[k] = addmod(acc[k], v[k], q);
        if (t + HEC_DOT_STAGES < T) {   // the stage is free once every thread has read it
            __syncthreads();
            if (tid == 0) issue(t + HEC_DOT_STAGES);
        }
    }
    ulonglong2 *out = reinterpret_cast<ulonglong2 *>(jd.out + base);
    out[tid] = make_ulonglong2(acc[0], acc[1]);
    out[256 + tid] = make_ulonglong2(acc[2], acc[3]);
}

// ---- exact basis extension (modUpExact / reconstructRNS / multSum,
// L:ring/ring_basis_extension.go:438-457,670-779) -----------------------------------------
#define HEC_MAXA 5
#define HEC_MUJOBS 12
struct ModupJob {
    const u64 *src[HEC_MAXA]; // coefficient-domain source limbs (any representative, < 2^64)
    int smod[HEC_MAXA];
    u64 qib[HEC_MAXA];        // MForm((Q_d/q_i)^-1 mod q_i)
    u64 qisp[HEC_MAXA];       // MForm(Q_d/q_i mod p_t) for this target
    u64 qpjinv[HEC_MAXA + 1]; // -v*Q_d mod p_t
    u64 *dst;
    int tmod, n;
};
struct ModupJobs { ModupJob j[HEC_MUJOBS]; };
// The same extension for ALL targets of one source digit: y_i and the float overflow count v depend only on the
// source, so they are computed once per coefficient (the per-target form above recomputes alpha Montgomery products
// and alpha double divisions for each of the up to nQ + nP - alpha targets: it was 51 % of a full-level key switch).
#define HEC_M2_MAXT 48 // targets per group (host splits larger groups); nQ + nP <= 33 in the reference's sets
// qisp[s]: MForm(Q_d/q_s mod p_t); for a target below 2^31 instead the plain residue c in the low word and c * 2^32 mod p_t in
// the high word (k_modup2's narrow path)
struct Modup2Target { u64 *dst; u64 qisp[HEC_MAXA]; u64 qpjinv[HEC_MAXA + 1]; int tmod; };
struct Modup2Job {
    const u64 *src[HEC_MAXA];
    int smod[HEC_MAXA];
    u64 qib[HEC_MAXA];
    size_t first;                // index of the job's first entry in the launch's target table
    int n, ntargets;
};
__global__ void __launch_bounds__(256) k_modup2(const Modup2Job *__restrict__ jobs, const Modup2Target *__restrict__ targets,
                                                const ModC *__restrict__ mods) {
    HEC_PDL_SYNC();
    const Modup2Job &job = jobs[blockIdx.y];
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    u64 y[HEC_MAXA];
    double vi = 0.0;
#pragma unroll
    for (int s = 0; s < HEC_MAXA; s++) {
        y[s] = 0;
        if (s < job.n) {
            const u64 qs = mods[job.smod[s]].q;
            y[s] = mred(job.src[s][i], job.qib[s], qs, mods[job.smod[s]].qinv);
            // v = (uint64) sum_i float64(y_i)/float64(q_i): IEEE double, sequential, no FMA
            vi = __dadd_rn(vi, __ddiv_rn(__ull2double_rn(y[s]), __ull2double_rn(qs)));
        }
    }
    const u64 v = (u64)__double2ull_rz(vi);
    // the per-target tables (<= nQ + nP entries of 104 B) are the same for every thread: stage them in shared memory
    __shared__ Modup2Target sT[HEC_M2_MAXT];
    for (int k = threadIdx.x; k < job.ntargets * (int)(sizeof(Modup2Target) / 8); k += blockDim.x)
        reinterpret_cast<u64 *>(sT)[k] = reinterpret_cast<const u64 *>(targets + job.first)[k];
    __syncthreads();
    for (int t = 0; t < job.ntargets; t++) {
        const Modup2Target &T = sT[t];
        const u64 pt = mods[T.tmod].q, ptinv = mods[T.tmod].qinv;
        if (mods[T.tmod].small) {
            // target below 2^31: y * c = y_hi * (c 2^32 mod p) + y_lo * c, two 32 x 32 products per source (each term below
            // 2^63 + 2^60, the sum of five below 2^66: a 64-bit accumulator and a carry count), then one reduction
            u64 lo = 0;
            u32 carry = 0;
#pragma unroll
            for (int s = 0; s < HEC_MAXA; s++)
                if (s < job.n) {
                    const u32 c = (u32)T.qisp[s], c2 = (u32)(T.qisp[s] >> 32);
                    const u64 term = (u64)(u32)y[s] * c + (u64)(u32)(y[s] >> 32) * c2;
                    lo += term;
                    carry += lo < term;
                }
            const u64 rm = mods[T.tmod].rmod;                          // 2^64 mod p
            u64 acc = mred(lo, rm, pt, ptinv);                          // lo mod p
            for (u32 k = 0; k < carry; k++) acc = addmod(acc, rm, pt);
            T.dst[i] = addmod(acc, T.qpjinv[v], pt);
            continue;
        }
        // multSum (L:ring/ring_basis_extension.go:715-779): the alpha products are summed as 128-bit integers and
        // reduced once -- each is < 2^61 p_t, so the sum of <= 5 stays below p_t 2^64, the domain of one REDC
        u64 lo = 0, hi = 0;
#pragma unroll
        for (int s = 0; s < HEC_MAXA; s++)
            if (s < job.n) {
                const u64 pl = y[s] * T.qisp[s], ph = __umul64hi(y[s], T.qisp[s]);
                lo += pl;
                hi += ph + (lo < pl);
            }
        const u64 mq = __umul64hi(lo * ptinv, pt);
        u64 acc = hi - mq;
        if (hi < mq) acc += pt;
        T.dst[i] = addmod(acc, T.qpjinv[v], pt);
    }
}

// =========================================================================================
// (2) fused conv_then_pack kernels
// =========================================================================================
// Grid order of the fused kernels.  The tile index selects the twiddles (one 65 KB set per tile for the row kernels);
// with the job index fastest, the CTAs that follow each other on an SM share their tile, so the twiddles they stream stay
// in L1 instead of cycling through four tiles' worth (148 mod 16 = 4).
#ifdef HEC_GRID_TILE_FASTEST
#define HEC_BTILE blockIdx.x
#define HEC_BJOB blockIdx.y
#define HEC_GRID(tiles, jobs) dim3((tiles), (jobs))
#else
#define HEC_BTILE blockIdx.y
#define HEC_BJOB blockIdx.x
#define HEC_GRID(tiles, jobs) dim3((jobs), (tiles))
#endif
// ---- Stage A: for every active output channel i and poly c (conv.go:525-531):
//   MulNew(ct_in, pl_ker[i])   L:ckks/evaluator.go:1360-1444 (pt branch)
//   SetScale -> MultByConst    L:ckks/evaluator.go:782-863 (constants from the host)
//            -> Rescale -> divRoundByLastModulusNTT   L:ring/ring_scaling.go:442-513
// The fixed operands of the point-wise products (kernel plaintexts, key limbs) are kept by the plan as Shoup pairs
// (w, floor(w 2^64 / q)): the product is then the 5-wide-multiply approximate Shoup form (31.9 clk) instead of a
// Montgomery product (49 clk), at the price of 16 instead of 8 bytes per operand word (L2-resident tables).
struct ConvA {
    const u64 *const *ctin; // [M] -> [2 polys][2 limbs][N]
    const ulonglong2 *const *ptk; // [B] -> [2 limbs][N] pairs of pl_ker[i] * c_limb (the MultByConst constant is folded
                            //        into the plan's copy of the kernel plaintexts)
    u64 *w1, *w2;           // [M*na*2][N] scratch
    u64 *xout;              // [M*na][2][N] level-0 ciphertexts
    int na, norm, mq0, mq1;
    u64 half1;              // (q1-1)>>1
    u64 hneg0;              // q0 - (half1 mod q0)
    ulonglong2 resc0;       // q0 - q1^-1 mod q0  (RescaleParams) as a Shoup pair
    // deferred-transform plans (k_def*): the level-0 ciphertexts leave stage A as pairs (U, e), value = U - NTT(e)
    u64 *uout, *eout;       // [M*na][2][N] each: U in the NTT domain, e in the coefficient domain (natural order)
    ulonglong2 q1inv;       // q1^-1 mod q0 as a Shoup pair (the q0 limb of ptk carries the same factor)
};
struct AJob {
    int c, a, m;
    __device__ __forceinline__ AJob(int job, int na) { c = job & 1; a = (job >> 1) % na; m = (job >> 1) / na; }
};
// A1: limb q1:  ct*(pt*k1) -> inverse stages t = 1..128
#ifndef HEC_A1_MINB
#define HEC_A1_MINB HEC_MINB
#endif
__global__ void __launch_bounds__(HEC_THREADS, HEC_A1_MINB) k_convA1(ConvA P, const ModC *__restrict__ mods) {
    __shared__ u64 sm[16 * HEC_ROW_PITCH];
    const AJob J(HEC_BJOB, P.na);
    const ModC M = mods[P.mq1];
    RowGeom G(HEC_BTILE);
    const u64 *ct = P.ctin[J.m] + (size_t)(J.c * 2 + 1) * HEC_N;
    const ulonglong2 *pt = P.ptk[J.a * P.norm] + HEC_N;
    u64 x[16];
#pragma unroll
    for (int k = 0; k < 16; k++) {
        u32 i = G.gbase + 16 * k;
        x[k] = shoup4(ct[i], __ldg(pt + i), M.q); // < 4q
    }
    row_AtoB(x, sm, G);
    row_inv8(x, sm, G, M);
    row_storeA(x, P.w1 + (size_t)HEC_BJOB * HEC_N, G);
}
// A2: finish InvNTT_q1, centre, lift into q0, forward stages m = 1..128 under q0
__global__ void __launch_bounds__(HEC_THREADS, HEC_MINB) k_convA2(ConvA P, const ModC *__restrict__ mods) {
    __shared__ u64 sm[HEC_TILE];
    const ModC M1 = mods[P.mq1];
    const ModC M0 = mods[P.mq0];
    ColGeom G(HEC_BTILE);
    const u64 *in = P.w1 + (size_t)HEC_BJOB * HEC_N;
    u64 *out = P.w2 + (size_t)HEC_BJOB * HEC_N;
    u64 x[16];
#pragma unroll
    for (int k = 0; k < 16; k++) x[k] = in[G.gB(k)];
    col_inv8_final(x, sm, G, M1);                           // InvNTT incl. its final pass, canonical
    const bool fits = M1.q <= M0.q; // t < q1 <= q0 is already a canonical residue of q0
#pragma unroll
    for (int k = 0; k < 16; k++) {
        u64 t = cred(x[k] + P.half1, M1.q);                 // + (q1-1)/2 mod q1
        x[k] = (fits ? t : canon(t, M0)) + P.hneg0;         // (t mod q0) - half  in [0,2q0)
    }
    col_fwd8(x, sm, G, M0);
#pragma unroll
    for (int k = 0; k < 16; k++) out[G.gB(k)] = x[k];
}
// A3: finish NTT_q0, combine with limb q0 of ct*(pt*k0):  out = (p0 - u) * q1^-1
// The products ct*pt are formed FIRST and parked in shared memory (each thread's own 16 words): their operand loads
// then sit at the start of the CTA, where the other CTAs of the SM cover them, instead of behind the transform
// (profiles/r02a: long-scoreboard stalls 5.9 per issue with the loads after the transform).
#define HEC_A3_SMEM (2 * 16 * HEC_ROW_PITCH * sizeof(u64)) // dynamic: above the 48 KB static limit
__global__ void __launch_bounds__(HEC_THREADS, HEC_MINB) k_convA3(ConvA P, const ModC *__restrict__ mods) {
    extern __shared__ __align__(128) u64 dsm[];
    u64 *sm = dsm, *st = dsm + 16 * HEC_ROW_PITCH;
    const AJob J(HEC_BJOB, P.na);
    const ModC M = mods[P.mq0];
    RowGeom G(HEC_BTILE);
    const u64 *ct = P.ctin[J.m] + (size_t)(J.c * 2) * HEC_N;
    const ulonglong2 *pt = P.ptk[J.a * P.norm];
#pragma unroll
    for (int k = 0; k < 16; k++) {
        u32 i = G.gbase + 16 * k, e = G.p + 16 * k;
        st[G.sbase + e + (e >> 4)] = shoup4(ct[i], __ldg(pt + i), M.q); // < 4q
    }
    u64 x[16];
    row_loadA(x, P.w2 + (size_t)HEC_BJOB * HEC_N, G);
    row_fwd8(x, sm, G, M);
    row_BtoA(x, sm, G);
    u64 *out = P.xout + (size_t)HEC_BJOB * HEC_N; // ((m*na + a)*2 + c) == job
#pragma unroll
    for (int k = 0; k < 16; k++) {
        u32 i = G.gbase + 16 * k, e = G.p + 16 * k;
        out[i] = cred(shoup(x[k] + 2 * M.q2 - st[G.sbase + e + (e >> 4)], P.resc0, M.q), M.q);
    }
}

// plan set-up (once per plan): a Montgomery-form plaintext limb as Shoup pairs (w, floor(w * 2^64 / q)), or as plain
// residues.  The quotient is a 64-step restoring division -- q < 2^61, so the shifted remainder never overflows.
__global__ void __launch_bounds__(256) k_plan_tables(const u64 *__restrict__ in, ulonglong2 *pairs, u64 *plain, int mod,
                                                     const ModC *__restrict__ mods, u64 factor = 1ull) {
    const u64 q = mods[mod].q, qinv = mods[mod].qinv;
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < HEC_N; i += gridDim.x * blockDim.x) {
        const u64 w = mred(in[i], factor, q, qinv); // out of Montgomery form, times `factor` (a plain residue)
        if (plain) plain[i] = w;
        if (pairs) {
            u64 rem = w, quo = 0;
            for (int b = 0; b < 64; b++) {
                rem <<= 1; quo <<= 1;
                if (rem >= q) { rem -= q; quo |= 1; }
            }
            pairs[i] = make_ulonglong2(w, quo);
        }
    }
}

// plan set-up: number of positions where two limbs differ
__global__ void __launch_bounds__(256) k_count_diff(const u64 *__restrict__ a, const u64 *__restrict__ b, int *ndiff) {
    int d = 0;
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < HEC_N; i += gridDim.x * blockDim.x) d += a[i] != b[i];
    if (d) atomicAdd(ndiff, d);
}

// ---- Stage B: one level of pack_ctxts (conv.go:286-297).  For butterfly (a = ct[i], b = ct[i+step]):
//   tmp1 = b * X^step ; tmp2 = a - tmp1 ; tmp1 = a + tmp1 ; tmp2 = RotateGal(tmp2, 2^j+1) ; out = tmp1 + tmp2
// RotateGal -> permuteNTT (L:ckks/evaluator.go:1575-1597) -> SwitchKeysInPlace
// (L:rlwe/keyswitch.go:72-225; level 0, alpha = 1, beta = 1: single-limb digit = copy path)
// -> ModDownSplitNTTPQ (L:ring/ring_basis_extension.go:247-291) -> + c0 -> PermuteNTTWithIndexLvl.
struct ConvB {
    const u64 *xin;  // [M*n][2][N]
    u64 *xout;       // [M*n/2][2][N]
    const ulonglong2 *mono; // NTT(X^step) at q0 as Shoup pairs (plan-owned, built from pt_idx)
    const ulonglong2 *key; // the Q limb of the level-0 key slice as Shoup pairs (plan-owned): [2 polys][N]  (B5)
    const u64 *keyP;       // the P limb of the key itself, Montgomery form: poly c at keyP + c * keyPstride     (B3: measured
    size_t keyPstride;     //   10 % slower on 16-byte pairs -- its 16 contiguous words per thread load as 8 x 128 bit)
    const u64 *bias; // bias plaintext as plain residues (plan-owned copy; only the last level), or null
    u64 *w1, *w2, *w3, *w4;
    u64 *z;          // [M*n/2][N]: tmp2.c1 = a1 - b1*X^step of every butterfly, in [0,3q)
    int n, mq0, mp0;
    u32 galEl;
    ulonglong2 negpinv; // q0 - P^-1 mod q0 as a Shoup pair   [A] test_run 0x4e5049
    u64 qpj1;        // q0 - (p0 mod q0) = qpjInv[1]
    u64 vthr;        // smallest y < p0 with uint64(float64(y)/float64(p0)) = 1 (~0: none): the float overflow count
                     // of the single-prime extension is a step function of y (hec_float_quotient_threshold)
    u32 mu0;         // floor(2^64 / q0)
    // deferred-transform plans (k_def*): xin / xout hold the U halves, ein / eout the e halves; `key` carries P^-1
    const u64 *ein;  // [M*n][2][N]
    u64 *eout;       // [M*n/2][2][N]
    u32 ginv;        // galEl^-1 mod 2N
    u32 step;        // the level's monomial is X^step
    ulonglong2 pinv; // P^-1 mod q0 as a Shoup pair
    // first pack level of a deferred plan: the U halves of its inputs are ct_in * (plaintext table) and are formed where
    // they are used instead of being written by stage A and read back
    const u64 *const *ctin;        // [M] -> [2 polys][2 limbs][N]
    const ulonglong2 *ptkz, *ptks; // [n/2][N] pairs: ptk'[a] -+ X^step ptk'[b] per butterfly (ptk' = pl_ker * k0 / q1 on limb q0)
};
struct BJob {
    int c, u, m;
    const u64 *a, *b;
    __device__ __forceinline__ BJob(int job, bool has_c, const ConvB &P) {
        int nb = P.n >> 1;
        c = has_c ? (job & 1) : 0;
        int bu = has_c ? (job >> 1) : job;
        m = bu / nb; u = bu % nb;
        a = P.xin + (size_t)(m * P.n + u) * 2 * HEC_N;
        b = P.xin + (size_t)(m * P.n + u + nb) * 2 * HEC_N;
    }
};
// B1: z = tmp2.c1 = a1 - b1*mono (also stored) ; inverse stages t = 1..128 under q0      grid.y = M*nb
#ifndef HEC_B1_MINB
#define HEC_B1_MINB 2 // measured: B1 0.483 -> 0.440 ms per 64 convolutions (its loads are latency-bound: more of them in flight at 128 registers); A1 and k_defA2 lose with the same change
#endif
__global__ void __launch_bounds__(HEC_THREADS, HEC_B1_MINB) k_convB1(ConvB P, const ModC *__restrict__ mods) {
    __shared__ u64 sm[16 * HEC_ROW_PITCH];
    const BJob J(HEC_BJOB, false, P);
    const ModC M = mods[P.mq0];
    RowGeom G(HEC_BTILE);
    u64 x[16];
#pragma unroll
    for (int k = 0; k < 16; k++) {
        u32 i = G.gbase + 16 * k;
        x[k] = J.a[HEC_N + i] + M.q2 - shoup(J.b[HEC_N + i], __ldg(P.mono + i), M.q); // in (0,3q)
    }
    row_storeA(x, P.z + (size_t)HEC_BJOB * HEC_N, G); // kept for B5: the key product and tmp1 = 2a - tmp2 need it
    row_AtoB(x, sm, G);
    row_inv8(x, sm, G, M);
    row_storeA(x, P.w1 + (size_t)HEC_BJOB * HEC_N, G);
}
// B2: finish InvNTT_q0 (canonical digit), copy-path lift into p0, forward stages m = 1..128 under p0
__global__ void __launch_bounds__(HEC_THREADS, HEC_MINB) k_convB2(ConvB P, const ModC *__restrict__ mods) {
    __shared__ u64 sm[HEC_TILE];
    const ModC MQ = mods[P.mq0];
    const ModC MP = mods[P.mp0];
    ColGeom G(HEC_BTILE);
    const u64 *in = P.w1 + (size_t)HEC_BJOB * HEC_N;
    u64 *out = P.w2 + (size_t)HEC_BJOB * HEC_N;
    u64 x[16];
#pragma unroll
    for (int k = 0; k < 16; k++) x[k] = in[G.gB(k)];
    col_inv8_final(x, sm, G, MQ);
    const bool fits = MQ.q <= MP.q2; // the digit residue (< q0) is fed to the lazy forward as is
    if (!fits) {
#pragma unroll
        for (int k = 0; k < 16; k++) x[k] = canon(x[k], MP);
    }
    col_fwd8(x, sm, G, MP);
#pragma unroll
    for (int k = 0; k < 16; k++) out[G.gB(k)] = x[k];
}
// B3: finish NTT_p0 of the digit (once), then for both key polys: multiply by key[c] (P limb) and
//     run inverse stages t = 1..128 under p0                               grid.y = M*nb
#define HEC_B3_SMEM ((16 * HEC_ROW_PITCH + HEC_TILE) * sizeof(u64)) // dynamic: above the 48 KB static limit
#ifndef HEC_B3_MINB
#define HEC_B3_MINB 2 // measured: 1.268 -> 1.253 ms per 64 convolutions against 3 CTAs / 80 registers; unrolling the key-polynomial loop loses (spills at 80, no gain at 128)
#endif
#ifndef HEC_B3_UNROLL
#define HEC_B3_UNROLL 1
#endif
__global__ void __launch_bounds__(HEC_THREADS, HEC_B3_MINB) k_convB3(ConvB P, const ModC *__restrict__ mods) {
    extern __shared__ __align__(128) u64 dsm[];
    u64 *sm = dsm;
    u64 *stash = dsm + 16 * HEC_ROW_PITCH; // NTT_p0(digit), kept for the second key poly
    const ModC M = mods[P.mp0];
    RowGeom G(HEC_BTILE);
    {
        u64 x[16];
        row_loadA(x, P.w2 + (size_t)HEC_BJOB * HEC_N, G);
        row_fwd8(x, sm, G, M);
#pragma unroll
        for (int k = 0; k < 16; k++) stash[k * HEC_THREADS + threadIdx.x] = x[k]; // own slots only
    }
    constexpr int kUnrollC = HEC_B3_UNROLL; // (a macro is not expanded inside the pragma)
#pragma unroll kUnrollC
    for (int c = 0; c < 2; c++) {
        const ulonglong2 *kv = reinterpret_cast<const ulonglong2 *>(P.keyP + c * P.keyPstride + G.b * 256 + 16 * G.p);
        u64 y[16];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            ulonglong2 t = __ldg(kv + k);
            y[2 * k] = mred_lazy(stash[(2 * k) * HEC_THREADS + threadIdx.x], t.x, M.q, M.qinv);
            y[2 * k + 1] = mred_lazy(stash[(2 * k + 1) * HEC_THREADS + threadIdx.x], t.y, M.q, M.qinv);
        }
        row_inv8(y, sm, G, M);
        row_storeA(y, P.w3 + (size_t)(HEC_BJOB * 2 + c) * HEC_N, G);
    }
}
// B4: finish InvNTTLazy_p0, exact basis extension P -> q0 (float64 overflow count v),
//     forward stages m = 1..128 under q0                                   grid.y = M*nb*2
__global__ void __launch_bounds__(HEC_THREADS, HEC_MINB) k_convB4(ConvB P, const ModC *__restrict__ mods) {
    __shared__ u64 sm[HEC_TILE];
    const ModC MP = mods[P.mp0];
    const ModC MQ = mods[P.mq0];
    ColGeom G(HEC_BTILE);
    const u64 *in = P.w3 + (size_t)HEC_BJOB * HEC_N;
    u64 *out = P.w4 + (size_t)HEC_BJOB * HEC_N;
    u64 x[16];
#pragma unroll
    for (int k = 0; k < 16; k++) x[k] = in[G.gB(k)];
    col_inv8_final(x, sm, G, MP);
#pragma unroll
    for (int k = 0; k < 16; k++) {
        // y = MRed(e, qibMont) with qibMont = MForm(1): the canonical residue mod p0; v = uint64(float64(y)/float64(p0))
        // is 0 below P.vthr and 1 from there on
        const u64 y = x[k];
        x[k] = reduce_lazy(y, MQ.q, P.mu0) + (y >= P.vthr ? P.qpj1 : 0ull); // in [0,3q0)
    }
    col_fwd8(x, sm, G, MQ);
#pragma unroll
    for (int k = 0; k < 16; k++) out[G.gB(k)] = x[k];
}
// B5: finish NTT_q0, mod-down combine with acc_Q = z*key[c] (Q limb), + tmp2.c0 (c = 0),
//     apply sigma_g inside the 256-word block, add tmp1 (+ bias)           grid.y = M*nb*2
//     Sums are lazy (every term a known multiple of q away from canonical) and reduced once at the end.
#define HEC_B5_SMEM (2 * 16 * HEC_ROW_PITCH * sizeof(u64)) // dynamic: above the 48 KB static limit
// the point-wise half of B5 for one polynomial.  A template on the polynomial rather than a test inside the loop: with
// a branch in every iteration the compiler cannot move the loads of iteration k+1 above the arithmetic of iteration k,
// and the kernel waits for every operand in turn (profiles/r02a: 71 % of its stall samples sat on these loads).
template <bool C0>
__device__ __forceinline__ void b5_pointwise(const u64 (&x)[16], u64 *sm, u64 *st, const ConvB &P, const BJob &J, const ModC &M,
                                             const RowGeom &G, const u64 *__restrict__ zb, const ulonglong2 *__restrict__ kq) {
    const u64 *__restrict__ a = C0 ? J.a : J.a + HEC_N;
    const u64 *__restrict__ b = J.b;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const u32 i = G.gbase + 16 * k, e = G.p + 16 * k;
        const u64 z = zb[i];                                                // tmp2.c1 in (0,3q)
        const u64 accq = shoup4(z, __ldg(kq + i), M.q);                     // MulCoeffsMontgomeryConstant (+ Reduce), < 4q
        u64 d = shoup(x[k] + 2 * M.q2 - accq, P.negpinv, M.q);              // ModDownSplitNTTPQ combine, [0,2q)
        u64 t1;
        if (C0) {
            const u64 a0 = a[i];
            const u64 m0 = shoup(b[i], __ldg(P.mono + i), M.q);             // [0,2q)
            d += a0 + M.q2 - m0;                                            // + tmp2.c0  (AddLvl), < 5q
            t1 = a0 + m0;                                                   // tmp1.c0, < 3q
            if (P.bias != nullptr) t1 += __ldg(P.bias + i);                 // < 4q
        } else {
            t1 = 2 * a[i] + 3 * M.q - z;                                    // tmp1.c1 = a1 + b1*X^step = 2 a1 - tmp2.c1, in (0,5q)
        }
        sm[G.sbase + e + (e >> 4)] = d;
        st[G.sbase + e + (e >> 4)] = t1;
    }
}
#ifndef HEC_B5_MINB
#define HEC_B5_MINB 2 // 128 registers: the point-wise half keeps more operand loads in flight (measured: B5 1.51 -> 1.37 ms
                      // per 64 convolutions against 3 CTAs / 80 registers; the transform half loses less than that gains)
#endif
__global__ void __launch_bounds__(HEC_THREADS, HEC_B5_MINB) k_convB5(ConvB P, const ModC *__restrict__ mods) {
    extern __shared__ __align__(128) u64 dsm[];
    u64 *sm = dsm;                         // exchange buffer of the transform, then the mod-down results d
    u64 *st = dsm + 16 * HEC_ROW_PITCH;    // tmp1 of the thread's own coefficients: parked here rather than in 32 registers, so
                                           // that the loads of the epilogue's operand streams can run ahead of their use
    const BJob J(HEC_BJOB, true, P);
    const ModC M = mods[P.mq0];
    RowGeom G(HEC_BTILE);
    u64 x[16];
    row_loadA(x, P.w4 + (size_t)HEC_BJOB * HEC_N, G);
    row_fwd8(x, sm, G, M);
    row_BtoA(x, sm, G);
    const ulonglong2 *kq = P.key + (size_t)J.c * HEC_N; // Q limb of key poly c
    const u64 *zb = P.z + (size_t)(HEC_BJOB >> 1) * HEC_N;
    if (J.c == 0) b5_pointwise<true>(x, sm, st, P, J, M, G, zb, kq);
    else b5_pointwise<false>(x, sm, st, P, J, M, G, zb, kq);
    // sigma_g for g = 2^j + 1 keeps the top j - 1 bits of the (bit-reversed) slot index: for j >= 9 a slot stays inside its
    // 256-word block (this half-warp's own), for 5 <= j < 9 inside the CTA's 4096-word tile (the wide packings: B > 256)
    const bool in_block = P.galEl > 512u;
    if (in_block) __syncwarp(); else __syncthreads();
    u64 *out = P.xout + (size_t)HEC_BJOB * HEC_N;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        u32 i = G.gbase + 16 * k;
        u32 s = perm_index(i, P.galEl) & (HEC_TILE - 1);             // position inside the tile
        u32 sb = in_block ? G.sbase : (s >> 8) * HEC_ROW_PITCH;
        s &= 255u;
        u32 e = G.p + 16 * k;
        out[i] = canon16(st[G.sbase + e + (e >> 4)] + sm[sb + s + (s >> 4)], M.q); // < 9q
    }
}

// =========================================================================================
// (3) the same path with the forward transforms deferred (tests/defer_model.py is the CPU model)
// =========================================================================================
// The rescale of stage A and both mod-downs of a butterfly END with a forward transform under q0 whose result is only
// ever added to, multiplied by a monomial, permuted by sigma_g -- all of which can be done on its coefficients -- or fed
// to the NEXT key switch, which starts by transforming back.  A level-0 polynomial is therefore carried as a pair
// (U, e), value = U - NTT(e) mod q0: U in the NTT domain (what came out of point-wise products), e in the coefficient
// domain (what came out of a basis change).  On e a monomial product is a negacyclic shift, sigma_g the index map
// n -> n g mod 2N with a sign, both exact; the value is formed only where it is needed: as d = InvNTT(U) - e for the c1
// polynomial that enters a key switch (d IS the digit; its transforms under q0 and p0 feed the key products), and once
// per output polynomial at the end.  Transforms per convolution: 3B + 5(B-1) + 2 instead of 4B + 6(B-1); every step is
// exact arithmetic modulo q0 and what passes through the special prime is unchanged, so the canonical result is the
// reference's bit for bit.  Requires the pack monomials to BE monomials (checked at plan creation) and g = 1 mod 256
// (sigma_g on coefficients then stays inside a column tile): B <= 256.

// X^step * e at coefficient n (e canonical): e[n - step], or -e[n - step + N] for the `step` coefficients that wrap
__device__ __forceinline__ u64 shifted_coeff(const u64 *__restrict__ e, u32 n, u32 step, u64 q) {
    return n >= step ? e[n - step] : q - e[n + HEC_N - step]; // [0,q]
}
// dA1: limb q0: U = ct*(pt*k0/q1) (stored) ; limb q1: ct*(pt*k1) -> inverse stages t = 1..128   (A1 + the product of A3)
__global__ void __launch_bounds__(HEC_THREADS, HEC_MINB) k_defA1(ConvA P, const ModC *__restrict__ mods) {
    __shared__ u64 sm[16 * HEC_ROW_PITCH];
    const AJob J(HEC_BJOB, P.na);
    const ModC M = mods[P.mq1];
    const u64 q0 = mods[P.mq0].q;
    RowGeom G(HEC_BTILE);
    const u64 *ct = P.ctin[J.m] + (size_t)(J.c * 2) * HEC_N;
    const ulonglong2 *pt = P.ptk[J.a * P.norm];
    u64 *uo = P.uout + (size_t)HEC_BJOB * HEC_N;
    u64 x[16];
#pragma unroll
    for (int k = 0; k < 16; k++) {
        u32 i = G.gbase + 16 * k;
        x[k] = shoup4(ct[HEC_N + i], __ldg(pt + HEC_N + i), M.q); // < 4q1
    }
#pragma unroll
    for (int k = 0; k < 16; k++) {
        u32 i = G.gbase + 16 * k;
        uo[i] = cred(shoup(ct[i], __ldg(pt + i), q0), q0);
    }
    row_AtoB(x, sm, G);
    row_inv8(x, sm, G, M);
    row_storeA(x, P.w1 + (size_t)HEC_BJOB * HEC_N, G);
}
// dA2: finish InvNTT_q1, centre, lift into q0, divide by q1: e (coefficient domain, natural order)
#ifndef HEC_DA2_MINB
#define HEC_DA2_MINB HEC_MINB
#endif
__global__ void __launch_bounds__(HEC_THREADS, HEC_DA2_MINB) k_defA2(ConvA P, const ModC *__restrict__ mods) {
    __shared__ u64 sm[HEC_TILE];
    const ModC M1 = mods[P.mq1];
    const ModC M0 = mods[P.mq0];
    ColGeom G(HEC_BTILE);
    const u64 *in = P.w1 + (size_t)HEC_BJOB * HEC_N;
    u64 *out = P.eout + (size_t)HEC_BJOB * HEC_N;
    u64 x[16];
#pragma unroll
    for (int k = 0; k < 16; k++) x[k] = in[G.gB(k)];
    col_inv8_final(x, sm, G, M1);
    const bool fits = M1.q <= M0.q;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        u64 t = cred(x[k] + P.half1, M1.q);
        u64 r = (fits ? t : canon(t, M0)) + P.hneg0;                  // [0,2q0)
        out[G.gA(k)] = cred(shoup(r, P.q1inv, M0.q), M0.q);
    }
}
// dB1 = k_convB1 on the U halves (z = Ua1 - Ub1*mono, inverse stages t = 1..128); on the FIRST pack level Ua1 and Ub1 are
// both ct_in.c1 times a kernel plaintext, so z = ct_in.c1 * (ptk'[a] - X^step ptk'[b]): one product with a table the plan
// prepared per butterfly, no U halves to read (or to have been written), no z to store (k_defB5<true> does not need it)
__global__ void __launch_bounds__(HEC_THREADS, HEC_B1_MINB) k_defB1f(ConvB P, const ModC *__restrict__ mods) {
    __shared__ u64 sm[16 * HEC_ROW_PITCH];
    const ModC M = mods[P.mq0];
    RowGeom G(HEC_BTILE);
    const int nb = P.n >> 1, m = HEC_BJOB / nb, u = HEC_BJOB % nb;
    const u64 *ct1 = P.ctin[m] + (size_t)2 * HEC_N;            // limb q0 of c1
    const ulonglong2 *tz = P.ptkz + (size_t)u * HEC_N;
    u64 x[16];
#pragma unroll
    for (int k = 0; k < 16; k++) {
        u32 i = G.gbase + 16 * k;
        x[k] = shoup4(ct1[i], __ldg(tz + i), M.q); // < 4q
    }
    row_AtoB(x, sm, G);
    row_inv8(x, sm, G, M);
    row_storeA(x, P.w1 + (size_t)HEC_BJOB * HEC_N, G);
}
// dB2: finish InvNTT_q0, subtract the e half of tmp2.c1 (ea1 - X^step eb1): the digit d, canonical; forward stages
//      m = 1..128 of d under p0 (-> w2) AND under q0 (-> w4): the value of tmp2.c1 is NTT_q0(d)
#define HEC_DB2_SMEM (2 * HEC_TILE * sizeof(u64))
#ifndef HEC_DB2_MINB
#define HEC_DB2_MINB HEC_MINB
#endif
#ifndef HEC_DB4_MINB
#define HEC_DB4_MINB 2 // measured: 1.389 -> 1.331 ms per 64 convolutions against 3 CTAs / 80 registers (k_defB2 and k_defB5 lose with the other setting)
#endif
#ifndef HEC_DB5_MINB
#define HEC_DB5_MINB HEC_B5_MINB
#endif
__global__ void __launch_bounds__(HEC_THREADS, HEC_DB2_MINB) k_defB2(ConvB P, const ModC *__restrict__ mods) {
    extern __shared__ __align__(128) u64 dsm[];
    u64 *sm = dsm, *stash = dsm + HEC_TILE;
    const ModC MQ = mods[P.mq0];
    const ModC MP = mods[P.mp0];
    ColGeom G(HEC_BTILE);
    const int nb = P.n >> 1, m = HEC_BJOB / nb, u = HEC_BJOB % nb;
    const u64 *ea = P.ein + ((size_t)(m * P.n + u) * 2 + 1) * HEC_N;
    const u64 *eb = P.ein + ((size_t)(m * P.n + u + nb) * 2 + 1) * HEC_N;
    const u64 *in = P.w1 + (size_t)HEC_BJOB * HEC_N;
    // the e operands FIRST, parked in the thread's own stash slots: their loads sit at the start of the CTA, where the
    // other CTAs of the SM cover them, instead of behind the inverse transform (as in k_convA3 / k_convB5)
    const u64 *__restrict__ ebs = eb - P.step; // only slot k = 0 can hold a coefficient n < step (row 0)
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const u32 n = G.gA(k);
        stash[k * HEC_THREADS + threadIdx.x] = (MQ.q - ea[n]) + (k == 0 ? shifted_coeff(eb, n, P.step, MQ.q) : ebs[n]); // [0,2q]
    }
    u64 x[16];
#pragma unroll
    for (int k = 0; k < 16; k++) x[k] = in[G.gB(k)];
    col_inv8_final(x, sm, G, MQ);
#pragma unroll
    for (int k = 0; k < 16; k++) {
        u64 d = x[k] + stash[k * HEC_THREADS + threadIdx.x]; // [0,3q]
        d = cred(cred(d, MQ.q2), MQ.q);
        x[k] = d;
        stash[k * HEC_THREADS + threadIdx.x] = d; // own slots only
    }
    if (!(MQ.q <= MP.q2)) {
#pragma unroll
        for (int k = 0; k < 16; k++) x[k] = canon(x[k], MP);
    }
    col_fwd8(x, sm, G, MP);
    u64 *outp = P.w2 + (size_t)HEC_BJOB * HEC_N;
#pragma unroll
    for (int k = 0; k < 16; k++) outp[G.gB(k)] = x[k];
#pragma unroll
    for (int k = 0; k < 16; k++) x[k] = stash[k * HEC_THREADS + threadIdx.x];
    col_fwd8(x, sm, G, MQ);
    u64 *outq = P.w4 + (size_t)HEC_BJOB * HEC_N;
#pragma unroll
    for (int k = 0; k < 16; k++) outq[G.gB(k)] = x[k];
}
// dB3 = k_convB3 (NTT_p0 of the digit, key products on the P limb, inverse stages t = 1..128 for both key polys)
// dB4: finish InvNTTLazy_p0, exact basis extension P -> q0, divide by P: the e half of the key-switch output; apply
//      sigma_g on coefficients (row permutation with sign inside the column tile); add the e halves of tmp1 and, for
//      c = 0, of sigma(tmp2.c0)                                                     grid.y = M*nb*2
#define HEC_DB4_SMEM (2 * HEC_TILE * sizeof(u64))
// everything of the result that does not pass through the key switch, for the thread's own 16 coefficients:
//   (ea + X^step eb)[n]  +  (c = 0) sigma(ea - X^step eb)[n]     -- sigma read straight from global memory: it maps a row of
// the column tile onto another row (same 16 columns), so the permuted loads are as coalesced as the straight ones
template <bool C0>
__device__ __forceinline__ void defb4_operands(u64 *S, const u64 *__restrict__ ea, const u64 *__restrict__ eb, const ConvB &P,
                                               const ColGeom &G, u64 q) {
    // only coefficients n < step wrap around (row 0 of the limb: slot k = 0 of the threads with pg = 0; step <= 128), and
    // a permuted row index is never 0 unless the row itself is: everything else is a plain load at n - step
    const u64 *__restrict__ ebs = eb - P.step;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const u32 n = G.gA(k);
        u64 s = ea[n] + (k == 0 ? shifted_coeff(eb, n, P.step, q) : ebs[n]); // <= 2q
        if (C0) {
            const u32 src = (n * P.ginv) & (2u * HEC_N - 1u);
            const u32 n2 = (src & (HEC_N - 1u));                              // same column, row (src >> 8) & 255
            const u64 d2 = ea[n2] + (q - shifted_coeff(eb, n2, P.step, q));   // [0,2q]
            s += (src & HEC_N) ? 2 * q - d2 : d2;                             // <= 4q
        }
        S[k * HEC_THREADS + threadIdx.x] = s;
    }
}
__global__ void __launch_bounds__(HEC_THREADS, HEC_DB4_MINB) k_defB4(ConvB P, const ModC *__restrict__ mods) {
    extern __shared__ __align__(128) u64 dsm[];
    u64 *sm = dsm;              // exchange buffer of the transform, then the e half of the key-switch output (for sigma)
    u64 *S = dsm + HEC_TILE;    // own slots: the operand sums, loaded before the transform
    const ModC MP = mods[P.mp0];
    const ModC MQ = mods[P.mq0];
    ColGeom G(HEC_BTILE);
    const int nb = P.n >> 1, c = HEC_BJOB & 1, bu = HEC_BJOB >> 1, m = bu / nb, u = bu % nb;
    const u64 *ea = P.ein + ((size_t)(m * P.n + u) * 2 + c) * HEC_N;
    const u64 *eb = P.ein + ((size_t)(m * P.n + u + nb) * 2 + c) * HEC_N;
    const u64 *in = P.w3 + (size_t)HEC_BJOB * HEC_N;
    if (c == 0) defb4_operands<true>(S, ea, eb, P, G, MQ.q);
    else defb4_operands<false>(S, ea, eb, P, G, MQ.q);
    u64 x[16];
#pragma unroll
    for (int k = 0; k < 16; k++) x[k] = in[G.gB(k)];
    col_inv8_final(x, sm, G, MP);                           // ends past its last barrier: sm is free again
#pragma unroll
    for (int k = 0; k < 16; k++) {
        // (y - v p0) / P mod q0 with P = p0: y / P - v -- the float overflow count v is 0 below P.vthr and 1 from there on,
        // and the Shoup product takes any y < 2^64, so the extension needs no reduction of its own
        const u64 y = x[k];
        sm[G.sA(k)] = shoup(y, P.pinv, MQ.q) + (y >= P.vthr ? MQ.q - 1 : 0ull);      // [0,3q)
    }
    __syncthreads();
    u64 *out = P.eout + (size_t)HEC_BJOB * HEC_N;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const u32 n = G.gA(k);
        const u32 src = (n * P.ginv) & (2u * HEC_N - 1u);    // sigma(T)[n] = +-T[n g^-1 mod 2N]; same column for g = 1 mod 256
        const u64 t = sm[((src >> 8) & 255u) * 16 + G.cc];
        out[n] = canon8(S[k * HEC_THREADS + threadIdx.x] + ((src & HEC_N) ? 3 * MQ.q - t : t), MQ.q); // <= 7q
    }
}
// dB5: finish NTT_q0 of the digit = the value of tmp2.c1, ONCE for both key polys; for each: product with key[c]/P
//      (Q limb), + the U half of tmp2.c0 (c = 0), sigma_g inside the 256-word block, + the U half of tmp1   grid.y = M*nb
template <bool C0, bool FIRST>
__device__ __forceinline__ void defb5_pointwise(const u64 *xs, u64 *sm, u64 *st, const ConvB &P, const BJob &J, const ModC &M,
                                                const RowGeom &G, const u64 *__restrict__ zb, const ulonglong2 *__restrict__ kq,
                                                const u64 *__restrict__ ct, const ulonglong2 *__restrict__ tz,
                                                const ulonglong2 *__restrict__ ts) {
    const u64 *__restrict__ a = C0 ? J.a : J.a + HEC_N;
    const u64 *__restrict__ b = J.b;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const u32 i = G.gbase + 16 * k, e = G.p + 16 * k;
        u64 d = shoup4(xs[k * HEC_THREADS + threadIdx.x], __ldg(kq + i), M.q);  // U half of the key-switch output, < 4q
        u64 t1;
        if (FIRST) {
            const u64 cv = __ldg(ct + i);                                       // limb q0 of ct_in.c0 / .c1
            if (C0) d += shoup(cv, __ldg(tz + i), M.q);                         // + U half of tmp2.c0, < 6q
            t1 = shoup(cv, __ldg(ts + i), M.q);                                 // U half of tmp1, [0,2q)
        } else if (C0) {
            const u64 a0 = a[i];
            const u64 m0 = shoup(b[i], __ldg(P.mono + i), M.q);                 // [0,2q)
            d += a0 + M.q2 - m0;                                                // + U half of tmp2.c0, < 7q
            t1 = a0 + m0;                                                       // < 3q
        } else {
            t1 = 2 * a[i] + 3 * M.q - zb[i];                                    // Ua1 + Ub1*mono = 2 Ua1 - z, in (0,5q)
        }
        sm[G.sbase + e + (e >> 4)] = d;
        st[G.sbase + e + (e >> 4)] = t1;
    }
}
// The transform's result is parked in shared memory (the thread's own 16 slots) rather than kept in 32 registers across
// both polynomials: with it in registers the point-wise loops had no room to run their operand loads ahead (first
// capture: every iteration waited for its own 128-bit load, 6.3 warps per issue on the long scoreboard, 43 % pipe use).
#define HEC_DB5_SMEM ((2 * 16 * HEC_ROW_PITCH + HEC_TILE) * sizeof(u64))
template <bool FIRST>
__global__ void __launch_bounds__(HEC_THREADS, HEC_DB5_MINB) k_defB5(ConvB P, const ModC *__restrict__ mods) {
    extern __shared__ __align__(128) u64 dsm[];
    u64 *sm = dsm, *st = dsm + 16 * HEC_ROW_PITCH, *xs = dsm + 2 * 16 * HEC_ROW_PITCH;
    const BJob J(HEC_BJOB, false, P);
    const ModC M = mods[P.mq0];
    RowGeom G(HEC_BTILE);
    {
        u64 x[16];
        row_loadA(x, P.w4 + (size_t)HEC_BJOB * HEC_N, G);
        row_fwd8(x, sm, G, M);
        row_BtoA(x, sm, G);
#pragma unroll
        for (int k = 0; k < 16; k++) xs[k * HEC_THREADS + threadIdx.x] = x[k];
    }
    const u64 *zb = P.z + (size_t)HEC_BJOB * HEC_N;
    const u64 *ct = FIRST ? P.ctin[J.m] : nullptr;
    const ulonglong2 *tz = FIRST ? P.ptkz + (size_t)J.u * HEC_N : nullptr, *ts = FIRST ? P.ptks + (size_t)J.u * HEC_N : nullptr;
    const bool in_block = P.galEl > 512u;
#pragma unroll 1
    for (int c = 0; c < 2; c++) {
        if (c == 0) defb5_pointwise<true, FIRST>(xs, sm, st, P, J, M, G, zb, P.key, ct, tz, ts);
        else defb5_pointwise<false, FIRST>(xs, sm, st, P, J, M, G, zb, P.key + HEC_N, FIRST ? ct + 2 * HEC_N : nullptr, tz, ts);
        if (in_block) __syncwarp(); else __syncthreads();
        u64 *out = P.xout + (size_t)(HEC_BJOB * 2 + c) * HEC_N;
#pragma unroll
        for (int k = 0; k < 16; k++) {
            u32 i = G.gbase + 16 * k;
            u32 s = perm_index(i, P.galEl) & (HEC_TILE - 1);
            u32 sb = in_block ? G.sbase : (s >> 8) * HEC_ROW_PITCH;
            s &= 255u;
            u32 e = G.p + 16 * k;
            out[i] = canon16(st[G.sbase + e + (e >> 4)] + sm[sb + s + (s >> 4)], M.q); // < 10q
        }
        if (in_block) __syncwarp(); else __syncthreads();
    }
}
// dF1 / dF2: the one forward transform per output polynomial: out = U - NTT(e) (+ bias on c0), canonical
__global__ void __launch_bounds__(HEC_THREADS, HEC_MINB) k_defF1(const u64 *__restrict__ ein, u64 *__restrict__ w, int mq0,
                                                                 const ModC *__restrict__ mods) {
    __shared__ u64 sm[HEC_TILE];
    const ModC M = mods[mq0];
    ColGeom G(HEC_BTILE);
    const u64 *in = ein + (size_t)HEC_BJOB * HEC_N;
    u64 *out = w + (size_t)HEC_BJOB * HEC_N;
    u64 x[16];
#pragma unroll
    for (int k = 0; k < 16; k++) x[k] = in[G.gA(k)];
    col_fwd8(x, sm, G, M);
#pragma unroll
    for (int k = 0; k < 16; k++) out[G.gB(k)] = x[k];
}
__global__ void __launch_bounds__(HEC_THREADS, HEC_MINB) k_defF2(const u64 *__restrict__ w, const u64 *__restrict__ uin,
                                                                 const u64 *__restrict__ bias, u64 *__restrict__ xout, int mq0,
                                                                 const ModC *__restrict__ mods) {
    __shared__ u64 sm[16 * HEC_ROW_PITCH];
    const ModC M = mods[mq0];
    RowGeom G(HEC_BTILE);
    u64 x[16];
    row_loadA(x, w + (size_t)HEC_BJOB * HEC_N, G);
    row_fwd8(x, sm, G, M);
    row_BtoA(x, sm, G);
    const u64 *u = uin + (size_t)HEC_BJOB * HEC_N;
    u64 *out = xout + (size_t)HEC_BJOB * HEC_N;
    const bool add_bias = bias != nullptr && (HEC_BJOB & 1) == 0;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const u32 i = G.gbase + 16 * k;
        u64 r = u[i] + M.q - canon(x[k], M);                 // (0,2q)
        if (add_bias) r += __ldg(bias + i);                  // < 3q
        out[i] = cred(cred(r, M.q2), M.q);
    }
}
