// Micro-benchmark 3: 64-bit mulhi / Shoup modmul formulations that avoid IMAD.WIDE with a 64-bit addend
// and carry (the form __umul64hi compiles to), sm_100a.  Results in profiles/r01_ubench_int_pipe.txt.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
typedef unsigned int u32;
#define ITERS 2048
#define UNR 8

__device__ __forceinline__ u64 wide(u32 a, u32 b) { u64 r; asm("mul.wide.u32 %0, %1, %2;" : "=l"(r) : "r"(a), "r"(b)); return r; }
// hi64(x*y): four addend-free wide products, carries on the ALU pipe
__device__ __forceinline__ u64 mulhi_alu(u64 x, u64 y) {
    u32 x0 = (u32)x, x1 = (u32)(x >> 32), y0 = (u32)y, y1 = (u32)(y >> 32);
    u64 p00 = wide(x0, y0), p01 = wide(x0, y1), p10 = wide(x1, y0), p11 = wide(x1, y1);
    u64 mid = (p00 >> 32) + (u32)p01 + (u32)p10;
    return p11 + (p01 >> 32) + (p10 >> 32) + (mid >> 32);
}
// same, dropping the low product (error <= 2 in the quotient estimate: lazy Shoup)
__device__ __forceinline__ u64 mulhi_alu3(u64 x, u64 y) {
    u32 x0 = (u32)x, x1 = (u32)(x >> 32), y0 = (u32)y, y1 = (u32)(y >> 32);
    u64 p01 = wide(x0, y1), p10 = wide(x1, y0), p11 = wide(x1, y1);
    return p11 + (p01 >> 32) + (p10 >> 32);
}
__device__ __forceinline__ u64 mullo_alu(u64 x, u64 y) {
    u32 x0 = (u32)x, x1 = (u32)(x >> 32), y0 = (u32)y, y1 = (u32)(y >> 32);
    u64 p00 = wide(x0, y0);
    u32 c = x0 * y1 + x1 * y0;
    return p00 + ((u64)c << 32);
}
template <int OP>
__device__ __forceinline__ u64 shoup(u64 y, u64 w, u64 ws, u64 q) {
    if (OP == 0) { u64 qe = __umul64hi(y, ws); return y * w - qe * q; }
    if (OP == 1) { u64 qe = mulhi_alu(y, ws); return y * w - qe * q; }
    if (OP == 2) { u64 qe = mulhi_alu(y, ws); return mullo_alu(y, w) - mullo_alu(qe, q); }
    if (OP == 3) { u64 qe = mulhi_alu3(y, ws); return y * w - qe * q; }
    if (OP == 4) return __umul64hi(y, ws);
    if (OP == 5) return mulhi_alu(y, ws);
    if (OP == 6) return mulhi_alu3(y, ws);
    return y * w;
}
template <int OP>
__global__ void k(u64 *out, u64 q, u64 w, u64 ws) {
    u64 y[UNR];
#pragma unroll
    for (int i = 0; i < UNR; i++) y[i] = (threadIdx.x * 977 + i * 31 + blockIdx.x) * 0x9E3779B97F4A7C15ull % q;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < UNR; i++) y[i] = shoup<OP>(y[i], w + it, ws + it, q) + i;
    }
    u64 s = 0;
#pragma unroll
    for (int i = 0; i < UNR; i++) s += y[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int OP>
void run(const char *name) {
    u64 *d;
    cudaMalloc(&d, 148 * 4 * 512 * 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    u64 q = 0x80000000080001ull, w = 0x123456789abcdull, ws = (u64)(((unsigned __int128)w << 64) / q);
    k<OP><<<148 * 4, 512>>>(d, q, w, ws);
    cudaEventRecord(e0);
    k<OP><<<148 * 4, 512>>>(d, q, w, ws);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    int clk;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double ops = 148.0 * 4 * 16 * ITERS * UNR;   // warp-ops
    double cycles = ms * 1e-3 * clk * 1e3;
    printf("%-56s %8.3f ms  %5.2f /clk/SM  %5.1f clk per warp-op per SMSP\n", name, ms, ops * 32 / cycles / 148 / 32, 4.0 / (ops / cycles / 148));
    cudaFree(d);
}
int main() {
    run<4>("mulhi64  __umul64hi");
    run<5>("mulhi64  4 wide (no addend) + ALU carries");
    run<6>("mulhi64  3 wide (approx)    + ALU");
    run<7>("mullo64  y*w");
    run<0>("shoup    __umul64hi, y*w - qe*q");
    run<1>("shoup    mulhi 4 wide + ALU");
    run<2>("shoup    mulhi 4 wide + ALU, mullo via wide+2 imad");
    run<3>("shoup    mulhi 3 wide approx");
    return 0;
}
