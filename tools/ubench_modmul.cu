// Micro-benchmark: throughput of candidate modular-multiply formulations (sm_100a), in
// modmuls per clock per SM.  Evidence for the butterfly design; not part of the product.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
typedef unsigned long long u64;
typedef unsigned int u32;
#define ITERS 2048
#define UNR 8

__device__ __forceinline__ u64 mred_lazy(u64 x, u64 y, u64 q, u64 qinv) {
    u64 lo = x * y, hi = __umul64hi(x, y);
    return hi - __umul64hi(lo * qinv, q) + q;
}
__device__ __forceinline__ u64 shoup_c(u64 y, u64 w, u64 ws, u64 q) {
    u64 qe = __umul64hi(y, ws);
    return y * w - qe * q;
}
// hi64(a*b) with explicit 32-bit partial products and carry chain
__device__ __forceinline__ u64 mulhi_ptx(u64 a, u64 b) {
    u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32);
    u32 r2, r3;
    asm("{\n\t.reg .u32 t0,t1,t2,t3;\n\t"
        "mul.hi.u32 t1, %2, %4;\n\t"          // hi(a0*b0)
        "mad.lo.cc.u32 t1, %2, %5, t1;\n\t"   // + lo(a0*b1)
        "madc.hi.u32 t2, %2, %5, 0;\n\t"      // hi(a0*b1) + c
        "mad.lo.cc.u32 t1, %3, %4, t1;\n\t"   // + lo(a1*b0)
        "madc.hi.cc.u32 t2, %3, %4, t2;\n\t"  // + hi(a1*b0) + c
        "addc.u32 t3, 0, 0;\n\t"
        "mad.lo.cc.u32 %0, %3, %5, t2;\n\t"   // + lo(a1*b1)
        "madc.hi.u32 %1, %3, %5, t3;\n\t"     // + hi(a1*b1)
        "}" : "=r"(r2), "=r"(r3) : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
    return ((u64)r3 << 32) | r2;
}
__device__ __forceinline__ u64 shoup_ptx(u64 y, u64 w, u64 ws, u64 nq) {
    u64 qe = mulhi_ptx(y, ws);
    return y * w + qe * nq; // nq = -q mod 2^64
}
// approximate quotient: drop the a0*b0 partial product (qe off by at most 1 -> result < 3q)
__device__ __forceinline__ u64 mulhi_approx(u64 a, u64 b) {
    u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32);
    u64 m = (u64)a0 * b1;
    u64 n = (u64)a1 * b0;
    u64 s = (m >> 32) + (n >> 32) + (((m & 0xffffffffull) + (n & 0xffffffffull)) >> 32);
    return (u64)a1 * b1 + s;
}
__device__ __forceinline__ u64 shoup_approx(u64 y, u64 w, u64 ws, u64 nq) {
    u64 qe = mulhi_approx(y, ws);
    return y * w + qe * nq;
}
// quotient from the high 32 bits only of the Shoup companion: ws32 = floor(w*2^32/q) (w<q<2^61):
// qe = hi64(y * ws32 * 2^32) = (y * ws32) >> 32  -> error <= ~2, result < 4q.  2 wide mults.
__device__ __forceinline__ u64 shoup32(u64 y, u64 w, u32 ws32, u64 nq) {
    u32 y0 = (u32)y, y1 = (u32)(y >> 32);
    u64 qe = (u64)y1 * ws32 + (((u64)y0 * ws32) >> 32);
    return y * w + qe * nq;
}

// explicit 32-bit partial products: only mad.wide / mad.lo (full-rate IMAD), accumulating in place
__device__ __forceinline__ u64 shoup_ptx2(u64 y, u64 w, u64 ws, u64 nq) {
    u64 r;
    asm("{\n\t"
        ".reg .u32 y0,y1,w0,w1,s0,s1,n0,n1,t0,t1,u0,u1,v0,v1,e0,e1,r0,r1;\n\t"
        ".reg .u64 t,u,v,e,c,acc;\n\t"
        "mov.b64 {y0,y1}, %1;\n\t"
        "mov.b64 {w0,w1}, %2;\n\t"
        "mov.b64 {s0,s1}, %3;\n\t"
        "mov.b64 {n0,n1}, %4;\n\t"
        "mul.wide.u32 t, y0, s0;\n\t"
        "mov.b64 {t0,t1}, t;\n\t"
        "mov.b64 c, {t1, 0};\n\t"
        "mad.wide.u32 u, y0, s1, c;\n\t"        // y0*s1 + hi32(y0*s0)
        "mov.b64 {u0,u1}, u;\n\t"
        "mov.b64 c, {u0, 0};\n\t"
        "mad.wide.u32 v, y1, s0, c;\n\t"        // y1*s0 + lo32(u)
        "mov.b64 {v0,v1}, v;\n\t"
        "mov.b64 c, {u1, 0};\n\t"
        "mad.wide.u32 e, y1, s1, c;\n\t"        // y1*s1 + hi32(u)
        "mov.b64 c, {v1, 0};\n\t"
        "add.u64 e, e, c;\n\t"                  // + hi32(v)   -> qe = hi64(y*ws)
        "mov.b64 {e0,e1}, e;\n\t"
        "mul.wide.u32 acc, y0, w0;\n\t"
        "mad.wide.u32 acc, e0, n0, acc;\n\t"
        "mov.b64 {r0,r1}, acc;\n\t"
        "mad.lo.u32 r1, y0, w1, r1;\n\t"
        "mad.lo.u32 r1, y1, w0, r1;\n\t"
        "mad.lo.u32 r1, e0, n1, r1;\n\t"
        "mad.lo.u32 r1, e1, n0, r1;\n\t"
        "mov.b64 %0, {r0,r1};\n\t"
        "}" : "=l"(r) : "l"(y), "l"(w), "l"(ws), "l"(nq));
    return r;
}
__device__ __forceinline__ u64 shoup_nq(u64 y, u64 w, u64 ws, u64 nq) {
    return y * w + __umul64hi(y, ws) * nq;
}
// full butterflies (CT without correction, GS) on register pairs, as in fwd4/inv4
__device__ __forceinline__ void ct_free(u64 &X, u64 &Y, u64 w, u64 ws, u64 q, u64 q2) {
    u64 t = shoup_c(Y, w, ws, q);
    u64 x = X;
    X = x + t; Y = x - t + q2;
}
__device__ __forceinline__ void gs(u64 &X, u64 &Y, u64 w, u64 ws, u64 q, u64 q2) {
    u64 u = X, v = Y;
    u64 s = u + v;
    X = s >= q2 ? s - q2 : s;
    Y = shoup_c(u - v + q2, w, ws, q);
}
template <int OP>
__global__ void k(u64 *out, u64 q, u64 qinv, u64 w, u64 ws) {
    u64 y[UNR];
    u64 nq = 0 - q;
#pragma unroll
    for (int i = 0; i < UNR; i++) y[i] = (threadIdx.x * 977 + i * 31 + blockIdx.x) * 0x9E3779B97F4A7C15ull % q;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < UNR; i++) {
            if (OP == 0) y[i] = mred_lazy(y[i], w, q, qinv);
            if (OP == 1) y[i] = shoup_c(y[i], w, ws, q);
            if (OP == 2) y[i] = shoup_ptx(y[i], w, ws, nq);
            if (OP == 3) y[i] = shoup_approx(y[i], w, ws, nq);
            if (OP == 4) y[i] = shoup32(y[i], w, (u32)(ws >> 32), nq);
            if (OP == 5) y[i] = shoup_nq(y[i], w, ws, nq);
            if (OP == 8) y[i] = shoup_ptx2(y[i], w, ws, nq);
            if (OP == 6 && (i & 1) == 0) ct_free(y[i], y[i + 1], w + it, ws, q, 2 * q);
            if (OP == 7 && (i & 1) == 0) gs(y[i], y[i + 1], w + it, ws, q, 2 * q);
        }
    }
    u64 s = 0;
#pragma unroll
    for (int i = 0; i < UNR; i++) s += y[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
static int g_blocks = 4, g_threads = 512;
template <int OP>
void run(const char *name) {
    u64 *d;
    cudaMalloc(&d, 148 * 4 * 512 * 8);
    u64 q = 0x80000000080001ull, w = 0x1234567890abcdull % q;
    u64 qinv = 1;
    for (int i = 0; i < 7; i++) qinv *= 2 - q * qinv;
    u64 ws = (u64)((((unsigned __int128)w) << 64) / q);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<OP><<<148 * g_blocks, g_threads>>>(d, q, qinv, w, ws);
    cudaEventRecord(e0);
    k<OP><<<148 * g_blocks, g_threads>>>(d, q, qinv, w, ws);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double ops = 148.0 * g_blocks * g_threads * ITERS * UNR;
    double cycles = ms * 1e-3 * 1.965e9;
    printf("%-34s %8.3f ms  %6.2f modmul/clk/SM   %5.1f clk per warp-modmul per SMSP\n", name, ms, ops / cycles / 148,
           cycles * 148 * 4 / (ops / 32));
    cudaFree(d);
}
int main(int argc, char **argv) {
    if (argc > 2) { g_blocks = atoi(argv[1]); g_threads = atoi(argv[2]); }
    printf("warps per SM: %d\n", g_blocks * g_threads / 32);
    run<0>("montgomery lazy (C)");
    run<1>("shoup (C, __umul64hi)");
    run<2>("shoup (PTX mad/madc, +nq)");
    run<3>("shoup approx hi (3 wide)");
    run<4>("shoup32 (32-bit companion)");
    run<5>("shoup (C, + qe*(-q))");
    run<8>("shoup (PTX mad.wide/mad.lo only)");
    run<6>("CT butterfly free (x2 = per bfly)");
    run<7>("GS butterfly (x2 = per bfly)");
    return 0;
}
