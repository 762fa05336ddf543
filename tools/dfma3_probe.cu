// Does a DFMA with three distinct register operands cost more than one with two (register-bank reads)?
#include <cstdio>
#include <cuda_runtime.h>
#define REP8(x) x x x x x x x x
template <int OP>
__global__ void k(double *out, int iters, double s) {
    double a0 = threadIdx.x + 1.5, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    double b0 = a0 * s, b1 = a1 * s, b2 = a2 * s, b3 = a3 * s, b4 = a4 * s, b5 = a5 * s, b6 = a6 * s, b7 = a7 * s;
    double c0 = b0 * s, c1 = b1 * s, c2 = b2 * s, c3 = b3 * s, c4 = b4 * s, c5 = b5 * s, c6 = b6 * s, c7 = b7 * s;
    for (int it = 0; it < iters; ++it) {
        if (OP == 0) {  // a = a*b + c : three distinct registers per instruction
            REP8(asm volatile("fma.rn.f64 %0,%0,%8,%16; fma.rn.f64 %1,%1,%9,%17; fma.rn.f64 %2,%2,%10,%18; fma.rn.f64 %3,%3,%11,%19;"
                              "fma.rn.f64 %4,%4,%12,%20; fma.rn.f64 %5,%5,%13,%21; fma.rn.f64 %6,%6,%14,%22; fma.rn.f64 %7,%7,%15,%23;"
                              : "+d"(a0), "+d"(a1), "+d"(a2), "+d"(a3), "+d"(a4), "+d"(a5), "+d"(a6), "+d"(a7)
                              : "d"(b0), "d"(b1), "d"(b2), "d"(b3), "d"(b4), "d"(b5), "d"(b6), "d"(b7), "d"(c0), "d"(c1), "d"(c2), "d"(c3),
                                "d"(c4), "d"(c5), "d"(c6), "d"(c7));)
        } else if (OP == 1) {  // a = a*b + const : two distinct registers
            REP8(asm volatile("fma.rn.f64 %0,%0,%8,0d3FF0000000000001; fma.rn.f64 %1,%1,%9,0d3FF0000000000001; fma.rn.f64 %2,%2,%10,0d3FF0000000000001;"
                              "fma.rn.f64 %3,%3,%11,0d3FF0000000000001; fma.rn.f64 %4,%4,%12,0d3FF0000000000001; fma.rn.f64 %5,%5,%13,0d3FF0000000000001;"
                              "fma.rn.f64 %6,%6,%14,0d3FF0000000000001; fma.rn.f64 %7,%7,%15,0d3FF0000000000001;"
                              : "+d"(a0), "+d"(a1), "+d"(a2), "+d"(a3), "+d"(a4), "+d"(a5), "+d"(a6), "+d"(a7)
                              : "d"(b0), "d"(b1), "d"(b2), "d"(b3), "d"(b4), "d"(b5), "d"(b6), "d"(b7));)
        } else if (OP == 2) {  // a = a*a + const : one register
            REP8(asm volatile("fma.rn.f64 %0,%0,%0,0d3FF0000000000001; fma.rn.f64 %1,%1,%1,0d3FF0000000000001; fma.rn.f64 %2,%2,%2,0d3FF0000000000001;"
                              "fma.rn.f64 %3,%3,%3,0d3FF0000000000001; fma.rn.f64 %4,%4,%4,0d3FF0000000000001; fma.rn.f64 %5,%5,%5,0d3FF0000000000001;"
                              "fma.rn.f64 %6,%6,%6,0d3FF0000000000001; fma.rn.f64 %7,%7,%7,0d3FF0000000000001;"
                              : "+d"(a0), "+d"(a1), "+d"(a2), "+d"(a3), "+d"(a4), "+d"(a5), "+d"(a6), "+d"(a7));)
        } else if (OP == 3) {  // a = b * c (DMUL, 2 regs, dest distinct)
            REP8(asm volatile("mul.rn.f64 %0,%8,%16; mul.rn.f64 %1,%9,%17; mul.rn.f64 %2,%10,%18; mul.rn.f64 %3,%11,%19;"
                              "mul.rn.f64 %4,%12,%20; mul.rn.f64 %5,%13,%21; mul.rn.f64 %6,%14,%22; mul.rn.f64 %7,%15,%23;"
                              : "+d"(a0), "+d"(a1), "+d"(a2), "+d"(a3), "+d"(a4), "+d"(a5), "+d"(a6), "+d"(a7)
                              : "d"(b0), "d"(b1), "d"(b2), "d"(b3), "d"(b4), "d"(b5), "d"(b6), "d"(b7), "d"(c0), "d"(c1), "d"(c2), "d"(c3),
                                "d"(c4), "d"(c5), "d"(c6), "d"(c7));)
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
template <int OP>
void run(const char *name) {
    double *out; const int threads = 128, ctas = 148 * 4, iters = 4096;
    cudaMalloc(&out, sizeof(double) * ctas * threads);
    k<OP><<<ctas, threads>>>(out, 16, 1.0000001);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<OP><<<ctas, threads>>>(out, iters, 1.0000001);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); cudaFree(out);
    printf("%-34s %.3f ms -> %.2f cycles (at 1.9 GHz) per warp-instruction per SMSP\n", name, ms, ms * 1e-3 * 1.9e9 / (iters * 64.0 * 4));
}
int main() {
    run<0>("DFMA 3 distinct registers");
    run<1>("DFMA 2 registers + immediate");
    run<2>("DFMA 1 register + immediate");
    run<3>("DMUL 2 registers");
    return 0;
}
