// Device side of the NCCL-free multi-GPU step loop: release/acquire flags in peer memory over NVLink.
//
// Row-slab protocol (one process per GPU, peers' buffers mapped with CUDA IPC):
//   * a CTA that writes one of the slab's first/last HALO rows also stores it into the neighbour's halo rows (those CTAs
//     are scheduled first, csv_kernels.cu);
//   * the warp that completes the last local reduction group copies this rank's group sums into every peer's group-sum
//     array, raises arrive[rank] in every peer's CommBox (st.release.sys after a warp barrier) and then -- in the tail of
//     the SAME launch -- waits until all ranks have arrived and adds the 32 group sums in index order (identical on every
//     rank => bit-identical c1/c2 and stop decision everywhere): a CSV step is one launch on every rank (reduce.cuh);
//   * PM: the last boundary CTA of a launch raises the neighbour's flag, pm_wait_kernel polls it before the next launch.
// Nothing ever waits for a kernel on the SAME GPU that might not be resident, and peers only wait for launches that are
// queued before anything that waits for them.  Every wait is bounded (spin_until).
#pragma once
#include "common.cuh"

namespace cvb {

__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned int *p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned int *p, unsigned int v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_relaxed(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// Spin until the flag *p (written by a peer, only ever growing) reaches `need`.  Bounded: after SPIN_TIMEOUT_NS the
// wait gives up, marks this rank's CommBox and returns false -- a lost peer becomes an error status on the host
// (CVB_ERR_COMM) instead of a kernel that never ends.
__device__ __forceinline__ bool spin_until(const unsigned int *p, unsigned int need, CommBox *box) {
    if (ld_acquire_sys(p) >= need) return true;
    if (ld_relaxed(&box->timed_out)) return false;
    const unsigned long long t0 = global_ns();
    for (;;) {
#pragma unroll 1
        for (int i = 0; i < 256; ++i)
            if (ld_acquire_sys(p) >= need) return true;
        if (global_ns() - t0 > SPIN_TIMEOUT_NS) {
            *reinterpret_cast<volatile unsigned int *>(&box->timed_out) = 1u;
            __threadfence();
            return false;
        }
    }
}

}  // namespace cvb
