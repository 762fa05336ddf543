// Marginal cost (SM sub-partition cycles) of one extra instruction inserted into an issue-bound FP64/ALU mix
// (8 DFMA + ~24 integer ops per iteration and warp, 4 or 5 warps per sub-partition).
#include <cstdio>
#include <cuda_runtime.h>

template <int EXTRA>
__global__ void k(double *out, int iters, double a, double b, unsigned int m) {
    __shared__ double tab[64];
    if (threadIdx.x < 64) tab[threadIdx.x] = threadIdx.x;
    __syncthreads();
    double x[8];
    unsigned int z[8];
    float f = threadIdx.x;
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = threadIdx.x * 1e-3 + i + 1; z[i] = threadIdx.x + i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            x[i] = fma(x[i], a, b);
            z[i] = (z[i] ^ m) + (z[i] >> 3);
        }
        if (EXTRA == 1) { double y; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x[0])); z[0] += __double2hiint(y); }
        if (EXTRA == 2) { double y; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x[0])); z[0] += __double2hiint(y); }
        if (EXTRA == 3) { float y; asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(f)); f = y + 1.0f; }
        if (EXTRA == 4) { float y; asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(y) : "d"(x[0])); z[0] += __float_as_int(y); }
        if (EXTRA == 5) { z[0] += __shfl_up_sync(0xffffffffu, z[1], 1); }
        if (EXTRA == 6) { z[0] += __double2loint(tab[z[1] & 63]); }
        if (EXTRA == 7) { double y; asm volatile("cvt.rn.f64.u32 %0, %1;" : "=d"(y) : "r"(z[1])); z[0] += __double2hiint(y); }
        if (EXTRA == 8) { double y; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x[0])); z[0] += __double2hiint(y);
                          asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x[1])); z[1] += __double2hiint(y);
                          asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x[2])); z[2] += __double2hiint(y);
                          asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x[3])); z[3] += __double2hiint(y); }
        if (EXTRA == 9) { double y; asm volatile("fma.rn.f64 %0, %1, %1, %1;" : "=d"(y) : "d"(x[0])); z[0] += __double2hiint(y); }
    }
    double s = f; unsigned int t = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { s += x[i]; t += z[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + t;
}

template <int EXTRA>
double run(int warps_per_sm) {
    double *out; const int threads = 128, ctas = 148 * warps_per_sm / 4, iters = 8192;
    cudaMalloc(&out, sizeof(double) * ctas * threads);
    k<EXTRA><<<ctas, threads>>>(out, 64, 1.0000001, 1e-9, 0x9e3779b9u);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<EXTRA><<<ctas, threads>>>(out, iters, 1.0000001, 1e-9, 0x9e3779b9u);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); cudaFree(out);
    return ms * 1e-3 * 1.9e9 / iters / (warps_per_sm / 4);  // cycles (at 1.9 GHz) per warp-iteration per sub-partition
}

int main() {
    for (int w : {16, 20}) {
        const double base = run<0>(w);
        printf("warps/SM=%d  base iteration (8 DFMA + ALU): %.1f cycles per warp-iteration per SMSP\n", w, base);
        printf("  + MUFU.RSQ64H      %+.1f\n", run<1>(w) - base);
        printf("  + MUFU.RCP64H      %+.1f\n", run<2>(w) - base);
        printf("  + MUFU.RSQ (fp32)  %+.1f\n", run<3>(w) - base);
        printf("  + F2F.F32.F64      %+.1f\n", run<4>(w) - base);
        printf("  + SHFL             %+.1f\n", run<5>(w) - base);
        printf("  + LDS.64           %+.1f\n", run<6>(w) - base);
        printf("  + I2F.F64.U32      %+.1f\n", run<7>(w) - base);
        printf("  + 4x MUFU.RSQ64H   %+.1f\n", run<8>(w) - base);
        printf("  + DFMA             %+.1f\n", run<9>(w) - base);
    }
    return 0;
}
