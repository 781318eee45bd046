// Micro-benchmark: cycles per butterfly of the real fwd4 / inv4 register kernels (hec_dev.cuh) with
// twiddles from global memory (L1/L2 hits), without any global data traffic.  Evidence for where the
// fused kernels lose time relative to the bare butterfly (tools/ubench_modmul.cu).
#include <cstdio>
#include <cstdlib>
#include "../optimal_conv_b200/csrc/hec_dev.cuh"
#define ITERS 256

template <int OP>
__global__ void __launch_bounds__(256, 3) k(u64 *out, const ModC *mods, int it_n) {
    const ModC M = mods[0];
    u64 x[16];
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = (threadIdx.x * 977 + i * 31 + blockIdx.x) * 0x9E3779B97F4A7C15ull % M.q;
    for (int it = 0; it < it_n; it++) {
        u32 base = 16 + ((threadIdx.x >> 4) + it) % 240;
        if (OP == 0) fwd4<false, 1>(x, M.psi + 15 * base, M.q, M.q2);
        if (OP == 1) fwd4<true, 1>(x, M.psi + 15 * base, M.q, M.q2);
        if (OP == 2) inv4<false, 1>(x, M.psi_inv + 15 * base, M.q, M.q2);
        if (OP == 10) inv4<true, 1>(x, M.psi_inv + 15 * base, M.q, M.q2);
        if (OP == 3) { fwd4<false, 1>(x, M.psi, M.q, M.q2); }               // uniform twiddles (column layout A)
        if (OP == 4) {   // row layout B' with the natural NttPsi order: lane p reads psi[ng*(4096+16b+p)+gi]
            u32 b = (blockIdx.x * 16 + (threadIdx.x >> 4) + it) & 255, base = 4096 + 16 * b + (threadIdx.x & 15);
#pragma unroll
            for (int lg = 0; lg < 4; lg++) {
                const int d = 8 >> lg, ng = 1 << lg;
#pragma unroll
                for (int gi = 0; gi < ng; gi++) {
                    ulonglong2 w = __ldg(M.psi + ((ng * base + gi) & 65535));
#pragma unroll
                    for (int k = 0; k < d; k++) ct_bfly<false>(x[gi * 2 * d + k], x[gi * 2 * d + k + d], w, M.q, M.q2);
                }
            }
        }
        if (OP == 5) {   // the same stages with the thread-order table the kernels use (stride 16)
            u32 b = (blockIdx.x * 16 + (threadIdx.x >> 4) + it) & 255;
            fwd4<false, 16>(x, M.psi + HEC_TW_ROWB + 240 * b + (threadIdx.x & 15), M.q, M.q2);
        }
        if (OP >= 6) {   // bare products on 16 independent chains: exact Shoup, approximate Shoup, lazy Montgomery
            const ulonglong2 w = __ldg(M.psi + ((it * 7) & 255));
#pragma unroll
            for (int i = 0; i < 16; i++) {
                if (OP == 6) x[i] = shoup(x[i], w, M.q);
                if (OP == 7) x[i] = shoup4(x[i], w, M.q);
                if (OP == 8) x[i] = mred_lazy(x[i], w.x, M.q, M.qinv);
                if (OP == 9) x[i] = mred(x[i], w.x, M.q, M.qinv);
            }
        }
        if (OP == 0 || (OP >= 3 && OP < 6)) {                              // keep the free-mode values bounded
#pragma unroll
            for (int i = 0; i < 16; i++) x[i] &= 0x00ffffffffffffffull;
        }
    }
    u64 s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int OP>
void run(const char *name, const ModC *dm) {
    u64 *d;
    cudaMalloc(&d, 148 * 3 * 256 * 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<OP><<<148 * 3, 256>>>(d, dm, ITERS);
    cudaEventRecord(e0);
    k<OP><<<148 * 3, 256>>>(d, dm, ITERS);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double bfly = 148.0 * 3 * 256 * ITERS * (OP >= 6 ? 16 : 32); // 32 butterflies per fwd4/inv4 per thread, 16 bare products
    double cycles = ms * 1e-3 * 1.965e9;
    printf("%-48s %7.3f ms  %6.1f clk per warp-butterfly per SMSP\n", name, ms, cycles * 148 * 4 / (bfly / 32));
    cudaFree(d);
}
int main() {
    printf("HEC_ARITH = %d\n", HEC_ARITH);
    // a fake table is fine for timing: (w, ws) pairs
    u64 q = 0x80000000080001ull;
    ulonglong2 *tab;
    cudaMalloc(&tab, 65536 * sizeof(ulonglong2));
    ulonglong2 *h = (ulonglong2 *)malloc(65536 * sizeof(ulonglong2));
    for (int i = 0; i < 65536; i++) { u64 w = (0x1234567ull * (i + 1)) % q; h[i] = make_ulonglong2(w, (u64)((((unsigned __int128)w) << 64) / q)); }
    cudaMemcpy(tab, h, 65536 * sizeof(ulonglong2), cudaMemcpyHostToDevice);
    ModC m;
    m.q = q; m.q2 = 2 * q; m.qinv = 0xff7fffffbff7ffffull /* any odd value: timing only */; m.rmod = 1; m.ninv_w = 1; m.ninv_s = 1; m.psi = tab; m.psi_inv = tab; m.tight = 0; m.small = 0; m.mu = 0; m.pad = 0;
    ModC *dm;
    cudaMalloc(&dm, sizeof(ModC));
    cudaMemcpy(dm, &m, sizeof(ModC), cudaMemcpyHostToDevice);
    run<0>("fwd4<free>, per-half-warp twiddles", dm);
    run<1>("fwd4<tight>", dm);
    run<2>("inv4", dm);
    run<10>("inv4, deferred range corrections", dm);
    run<3>("fwd4<free>, uniform twiddles", dm);
    run<4>("fwd4<free>, row B', NttPsi order (16 lines/load)", dm);
    run<5>("fwd4<free>, row B', thread-order table", dm);
    run<6>("shoup (exact quotient), per product", dm);
    run<7>("shoup4 (approximate quotient), per product", dm);
    run<8>("mred_lazy (Montgomery), per product", dm);
    run<9>("mred (Montgomery, canonical), per product", dm);
    return 0;
}
