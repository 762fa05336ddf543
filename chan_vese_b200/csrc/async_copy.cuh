// Asynchronous global -> shared copies (cp.async / LDGSTS) used by the row rings of the CSV and PM kernels.
// 16-byte chunks, L2-only (.cg): the rows are streamed once per launch, L1 would not help.
#pragma once
#include <cuda_runtime.h>

namespace cvb {

__device__ __forceinline__ void cp_async16(unsigned int dst_shared, const void *src_global) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_shared), "l"(src_global) : "memory");
}
// 8-byte chunks (fp32 rows: two columns per lane); .cg exists for 16 bytes only, 8 bytes go through L1 (.ca)
__device__ __forceinline__ void cp_async8(unsigned int dst_shared, const void *src_global) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst_shared), "l"(src_global) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
// wait until at most N of this thread's committed groups are still in flight
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

}  // namespace cvb
