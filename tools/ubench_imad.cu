// Micro-benchmark: issue throughput of the integer instructions the NTT butterflies use
// (sm_100a).  Prints warp-instructions per cycle per SM for each opcode.  Not part of the
// product; evidence for DESIGN.md's "integer-multiply pipe" ceiling.
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
    for (int i = 0; i < UNR; i++) { a[i] = a0 + threadIdx.x + i; b[i] = b0 + i; w[i] = a[i]; }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < UNR; i++) {
            if (OP == 0) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a[i]), "r"(b[i]));
            if (OP == 1) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(a[i]), "r"(b[i]));
            if (OP == 2) asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(a[i]), "r"(b[i]));
            if (OP == 3) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i]));
            if (OP == 4) asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(a[i]), "+r"(b[i]) : "r"(b0), "r"(a0));
            if (OP == 5) asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w[i]) : "r"(a[i]), "r"((u32)w[i]));
            if (OP == 6) { u64 t; asm volatile("mul.hi.u64 %0, %1, %2;" : "=l"(t) : "l"(w[i]), "l"((u64)b[i] << 32 | a[i])); w[i] = t + it; }
            if (OP == 7) { u64 t; asm volatile("mul.lo.u64 %0, %1, %2;" : "=l"(t) : "l"(w[i]), "l"((u64)b[i] << 32 | a[i])); w[i] = t + it; }
        }
    }
    u32 s = 0;
#pragma unroll
    for (int i = 0; i < UNR; i++) s += a[i] + b[i] + (u32)w[i] + (u32)(w[i] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP>
void run(const char *name, double ops_per_iter) {
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
    double warp_instr = 148.0 * 2 * 32 * ITERS * UNR * ops_per_iter;
    double cycles = ms * 1e-3 * clk * 1e3;
    printf("%-28s %8.3f ms  %6.2f warp-instr/clk/SM (assuming %d kHz)\n", name, ms, warp_instr / cycles / 148, clk);
    cudaFree(d);
}
int main() {
    run<0>("mad.wide.u32 (IMAD.WIDE)", 1);
    run<1>("mad.lo.u32 (IMAD)", 1);
    run<2>("mad.hi.u32 (IMAD.HI)", 1);
    run<3>("add.u32 (IADD3)", 1);
    run<4>("add.cc+addc (IADD3,IADD3.X)", 2);
    run<5>("mul.wide.u32", 1);
    run<6>("mul.hi.u64 (+add)", 1);
    run<7>("mul.lo.u64 (+add)", 1);
    return 0;
}
