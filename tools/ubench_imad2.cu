// Micro-benchmark 2: issue cost of the IMAD forms a 64-bit modular multiply is made of (sm_100a), with the
// SASS each PTX form compiles to checked by cuobjdump (see profiles/r01_ubench_int_pipe.txt).
// Every chain feeds its own result back as a multiplicand so nothing is loop-invariant.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
typedef unsigned int u32;
#define ITERS 4096
#define UNR 8

template <int OP>
__global__ void k(u32 *out, u32 a0, u32 b0) {
    u32 a[UNR], b[UNR];
    u64 w[UNR];
#pragma unroll
    for (int i = 0; i < UNR; i++) { a[i] = a0 + threadIdx.x + i; b[i] = b0 + i; w[i] = a[i] * 77ull; }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < UNR; i++) {
            u32 lo = (u32)w[i], hi = (u32)(w[i] >> 32);
            if (OP == 0) asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w[i]) : "r"(lo), "r"(b[i]));            // IMAD.WIDE.U32 d, a, b, RZ
            if (OP == 1) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(lo), "r"(b[i]));        // IMAD.WIDE.U32 d, a, b, c
            if (OP == 2) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(hi));          // IMAD d, a, b, c
            if (OP == 3) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i]));                       // IMAD.HI.U32
            if (OP == 4) asm volatile("mad.lo.cc.u32 %0, %0, %2, %3; madc.hi.u32 %1, %0, %2, %1;" : "+r"(a[i]), "+r"(b[i]) : "r"(b0), "r"(a0));  // carry chain
            if (OP == 5) {  // 64x64 -> hi 64 from four addend-free wide products + ALU adds
                u32 x0 = lo, x1 = hi, y0 = b[i], y1 = a[i];
                u64 p00, p01, p10, p11;
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p00) : "r"(x0), "r"(y0));
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p01) : "r"(x0), "r"(y1));
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p10) : "r"(x1), "r"(y0));
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p11) : "r"(x1), "r"(y1));
                u64 mid = (p00 >> 32) + (u32)p01 + (u32)p10;
                w[i] = p11 + (p01 >> 32) + (p10 >> 32) + (mid >> 32);
            }
            if (OP == 6) w[i] = __umul64hi(w[i], ((u64)a[i] << 32) | b[i]);
        }
    }
    u32 s = 0;
#pragma unroll
    for (int i = 0; i < UNR; i++) s += a[i] + b[i] + (u32)w[i] + (u32)(w[i] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP>
void run(const char *name) {
    u32 *d;
    cudaMalloc(&d, 148 * 8 * 1024 * 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<OP><<<148 * 2, 1024>>>(d, 3, 5);
    cudaEventRecord(e0);
    k<OP><<<148 * 2, 1024>>>(d, 3, 5);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    int clk;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double ops = 148.0 * 2 * 32 * ITERS * UNR;
    double cycles = ms * 1e-3 * clk * 1e3;
    printf("%-52s %8.3f ms  %6.2f op/clk/SM  = %5.2f clk per warp-op per SMSP\n", name, ms, ops / cycles / 148, 4.0 / (ops / cycles / 148));
    cudaFree(d);
}
int main() {
    run<0>("mul.wide.u32          IMAD.WIDE.U32 d,a,b,RZ");
    run<1>("mad.wide.u32 (+c64)   IMAD.WIDE.U32 d,a,b,c");
    run<2>("mad.lo.u32            IMAD d,a,b,c");
    run<3>("mul.hi.u32            IMAD.HI.U32");
    run<4>("mad.lo.cc + madc.hi   (2 instr)");
    run<5>("mulhi64: 4 wide (no addend) + ALU adds");
    run<6>("mulhi64: __umul64hi");
    return 0;
}
