// Microbenchmark: FP64 pipe ceiling on this GPU (DFMA alone, DFMA mixed with ALU work, with MUFU.64H seeds).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu && ./fp64_peak
#include <cstdio>
#include <cuda_runtime.h>

template <int MIX, int CHAINS>
__global__ void k(double *out, int iters, double a, double b, unsigned int m) {
    double x[CHAINS];
    unsigned int z[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) { x[i] = threadIdx.x * 1e-3 + i; z[i] = threadIdx.x + i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) {
            x[i] = fma(x[i], a, b);
            if (MIX >= 1) z[i] = (z[i] ^ m) + (z[i] >> 3);           // 2-3 ALU ops per DFMA
            if (MIX >= 2) z[i] = (z[i] & 0xffff) | (z[i] << 7);
            if (MIX == 3 && (i & 7) == 0) {                           // one MUFU.RSQ64H per 8 DFMA
                double y; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x[i])); x[i] = fma(y, 1e-30, x[i]);
            }
        }
    }
    double s = 0; unsigned int t = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) { s += x[i]; t += z[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + t;
}

template <int MIX, int CHAINS>
void run(const char *name, int ctas_per_sm, int threads) {
    int sms = 148; double *out; cudaMalloc(&out, sizeof(double) * sms * ctas_per_sm * threads);
    const int iters = 4096;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MIX, CHAINS><<<sms * ctas_per_sm, threads>>>(out, 16, 1.0000001, 1e-9, 0x9e3779b9u);
    cudaEventRecord(e0);
    k<MIX, CHAINS><<<sms * ctas_per_sm, threads>>>(out, iters, 1.0000001, 1e-9, 0x9e3779b9u);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double dfma = (double)sms * ctas_per_sm * threads * iters * CHAINS * (MIX == 3 ? 1.125 : 1.0);
    printf("%-34s warps/SM=%2d  %.3f ms  %.2f T DFMA/s  (%.1f DFMA/clk/SM at 1.9 GHz)\n", name, ctas_per_sm * threads / 32, ms,
           dfma / ms / 1e9, dfma / (ms * 1e-3) / 148 / 1.9e9);
    cudaFree(out);
}

int main() {
    run<0, 8>("DFMA only, 8 chains", 4, 128);
    run<0, 8>("DFMA only, 8 chains", 8, 128);
    run<0, 2>("DFMA only, 2 chains", 4, 128);
    run<0, 2>("DFMA only, 2 chains", 8, 128);
    run<1, 8>("DFMA + 3 ALU, 8 chains", 4, 128);
    run<2, 8>("DFMA + 6 ALU, 8 chains", 4, 128);
    run<1, 2>("DFMA + 3 ALU, 2 chains", 4, 128);
    run<3, 8>("DFMA + 3 ALU + MUFU64/8, 8 chains", 4, 128);
    run<3, 8>("DFMA + 3 ALU + MUFU64/8, 8 chains", 6, 128);
    return 0;
}
