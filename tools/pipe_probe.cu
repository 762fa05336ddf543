// Per-instruction throughput probe (cycles per warp-instruction per SM sub-partition) for the instruction classes
// the CSV kernel is made of.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_probe pipe_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

#define REP16(x) x x x x x x x x x x x x x x x x
template <int OP>
__global__ void probe(double *out, long long *cyc, int iters) {
    double a0 = threadIdx.x + 1.5, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double b = 1.0000001, c = 1e-9;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (OP == 0) {  // 8 independent DFMA chains
            REP16(asm volatile("fma.rn.f64 %0,%0,%8,%9; fma.rn.f64 %1,%1,%8,%9; fma.rn.f64 %2,%2,%8,%9; fma.rn.f64 %3,%3,%8,%9;"
                               "fma.rn.f64 %4,%4,%8,%9; fma.rn.f64 %5,%5,%8,%9; fma.rn.f64 %6,%6,%8,%9; fma.rn.f64 %7,%7,%8,%9;"
                               : "+d"(a0), "+d"(a1), "+d"(a2), "+d"(a3), "+d"(a4), "+d"(a5), "+d"(a6), "+d"(a7) : "d"(b), "d"(c));)
        } else if (OP == 1) {  // one dependent DFMA chain (latency)
            REP16(asm volatile("fma.rn.f64 %0,%0,%1,%2; fma.rn.f64 %0,%0,%1,%2; fma.rn.f64 %0,%0,%1,%2; fma.rn.f64 %0,%0,%1,%2;"
                               "fma.rn.f64 %0,%0,%1,%2; fma.rn.f64 %0,%0,%1,%2; fma.rn.f64 %0,%0,%1,%2; fma.rn.f64 %0,%0,%1,%2;"
                               : "+d"(a0) : "d"(b), "d"(c));)
        } else if (OP == 2) {  // 8 independent MUFU.RSQ64H
            REP16(asm volatile("rsqrt.approx.ftz.f64 %0,%0; rsqrt.approx.ftz.f64 %1,%1; rsqrt.approx.ftz.f64 %2,%2; rsqrt.approx.ftz.f64 %3,%3;"
                               "rsqrt.approx.ftz.f64 %4,%4; rsqrt.approx.ftz.f64 %5,%5; rsqrt.approx.ftz.f64 %6,%6; rsqrt.approx.ftz.f64 %7,%7;"
                               : "+d"(a0), "+d"(a1), "+d"(a2), "+d"(a3), "+d"(a4), "+d"(a5), "+d"(a6), "+d"(a7));)
        } else if (OP == 3) {  // 8 independent MUFU.RCP64H
            REP16(asm volatile("rcp.approx.ftz.f64 %0,%0; rcp.approx.ftz.f64 %1,%1; rcp.approx.ftz.f64 %2,%2; rcp.approx.ftz.f64 %3,%3;"
                               "rcp.approx.ftz.f64 %4,%4; rcp.approx.ftz.f64 %5,%5; rcp.approx.ftz.f64 %6,%6; rcp.approx.ftz.f64 %7,%7;"
                               : "+d"(a0), "+d"(a1), "+d"(a2), "+d"(a3), "+d"(a4), "+d"(a5), "+d"(a6), "+d"(a7));)
        } else if (OP == 4) {  // 8 independent 32-bit shuffles
            int i0 = __double2loint(a0), i1 = i0 + 1, i2 = i0 + 2, i3 = i0 + 3, i4 = i0 + 4, i5 = i0 + 5, i6 = i0 + 6, i7 = i0 + 7;
            REP16(i0 = __shfl_up_sync(0xffffffffu, i0, 1); i1 = __shfl_up_sync(0xffffffffu, i1, 1); i2 = __shfl_up_sync(0xffffffffu, i2, 1);
                  i3 = __shfl_up_sync(0xffffffffu, i3, 1); i4 = __shfl_up_sync(0xffffffffu, i4, 1); i5 = __shfl_up_sync(0xffffffffu, i5, 1);
                  i6 = __shfl_up_sync(0xffffffffu, i6, 1); i7 = __shfl_up_sync(0xffffffffu, i7, 1);)
            a0 += i0 + i1 + i2 + i3 + i4 + i5 + i6 + i7;
        } else if (OP == 5) {  // DFMA + DADD + DMUL mix, 8 chains
            REP16(asm volatile("fma.rn.f64 %0,%0,%8,%9; add.rn.f64 %1,%1,%9; mul.rn.f64 %2,%2,%8; fma.rn.f64 %3,%3,%8,%9;"
                               "add.rn.f64 %4,%4,%9; mul.rn.f64 %5,%5,%8; fma.rn.f64 %6,%6,%8,%9; add.rn.f64 %7,%7,%9;"
                               : "+d"(a0), "+d"(a1), "+d"(a2), "+d"(a3), "+d"(a4), "+d"(a5), "+d"(a6), "+d"(a7) : "d"(b), "d"(c));)
        } else if (OP == 6) {  // 2 dependent chains (the ILP of two pixels)
            REP16(asm volatile("fma.rn.f64 %0,%0,%2,%3; fma.rn.f64 %1,%1,%2,%3; fma.rn.f64 %0,%0,%2,%3; fma.rn.f64 %1,%1,%2,%3;"
                               "fma.rn.f64 %0,%0,%2,%3; fma.rn.f64 %1,%1,%2,%3; fma.rn.f64 %0,%0,%2,%3; fma.rn.f64 %1,%1,%2,%3;"
                               : "+d"(a0), "+d"(a1) : "d"(b), "d"(c));)
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int OP>
void run(const char *name, int warps_per_sm) {
    double *out; long long *cyc, h;
    const int threads = 128, ctas = 148 * warps_per_sm / 4, iters = 256;
    cudaMalloc(&out, sizeof(double) * ctas * threads); cudaMalloc(&cyc, 8);
    probe<OP><<<ctas, threads>>>(out, cyc, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe<OP><<<ctas, threads>>>(out, cyc, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double inst_per_warp = (double)iters * 16 * 8;
    const int warps_per_smsp = warps_per_sm / 4;
    printf("%-30s warps/SMSP=%d  %8.2f cycles per warp-instruction per SMSP   (%.3f ms)\n", name, warps_per_smsp,
           (double)h / (inst_per_warp * warps_per_smsp), ms);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int w : {4, 16, 32}) {
        printf("---- %d warps per SM\n", w);
        if (w == 4) { run<0>("DFMA x8 independent", 4); run<1>("DFMA dependent chain", 4); run<6>("DFMA 2 chains", 4); run<5>("DFMA/DADD/DMUL mix", 4);
                      run<2>("MUFU.RSQ64H", 4); run<3>("MUFU.RCP64H", 4); run<4>("SHFL.UP", 4); }
        if (w == 16) { run<0>("DFMA x8 independent", 16); run<1>("DFMA dependent chain", 16); run<6>("DFMA 2 chains", 16); run<5>("DFMA/DADD/DMUL mix", 16);
                       run<2>("MUFU.RSQ64H", 16); run<3>("MUFU.RCP64H", 16); run<4>("SHFL.UP", 16); }
        if (w == 32) { run<0>("DFMA x8 independent", 32); run<1>("DFMA dependent chain", 32); run<2>("MUFU.RSQ64H", 32); run<4>("SHFL.UP", 32); }
    }
    return 0;
}
