// Host-side launchers of the sm_100a kernels (definitions in csv_kernels.cu / pm_kernels.cu).
#pragma once
#include <algorithm>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "common.cuh"

namespace cvb {

// Programmatic dependent launch of the step kernels: the next step's CTAs become resident during the tail of the
// running one (they load the atan table, then block in griddepcontrol.wait until the running step has completed and
// flushed).  On for every job, row slabs included: a slab's step kernel ends in a wait for the peers' sums, but every
// CTA of that kernel is already resident or done when the dependent launch may start, and the peers wait only for
// kernels that are queued BEFORE anything that waits for them (no cycle).  CVB_PDL=0 turns it off.
inline bool use_pdl(bool /*multi_rank*/) {
    static const int env = [] {
        const char *e = getenv("CVB_PDL");
        return e ? (e[0] == '0' ? 0 : 1) : -1;
    }();
    return env != 0;
}

// cudaFuncAttributePreferredSharedMemoryCarveout applies per DEVICE: remember it per device ordinal, not per process.
template <typename F>
inline cudaError_t prefer_max_shared(F *func) {
    static bool done[64] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64 && done[dev]) return cudaSuccess;
    e = cudaFuncSetAttribute(func, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
    if (e == cudaSuccess && dev >= 0 && dev < 64) done[dev] = true;
    return e;
}

cudaError_t launch_csv_step(const CsvArgs &A, bool strict, cudaStream_t s);
cudaError_t launch_csv_kappa(const CsvArgs &A, bool strict, cudaStream_t s);
// final_mode 1: reset counters + stop condition; 2: region means only
cudaError_t launch_csv_init(const CsvArgs &A, int final_mode, cudaStream_t s);
cudaError_t launch_csv_finalize(const CsvArgs &A, int mode, cudaStream_t s);
cudaError_t launch_delta_map(double *data, size_t n, double eps, cudaStream_t s);
cudaError_t launch_mask(const double *u, uint8_t *mask, int rows, int w, int pitch, int invert, cudaStream_t s);
// rule 0: separate()'s float32(u) > 0; rule 1: the video contour's saturate_cast<uchar>(u) > 0 (u > 0.5)
cudaError_t launch_mask_packed(const void *u, int f32, uint8_t *bits, int rows, int w, int pitch, int invert, cudaStream_t s,
                               int rule = 0);
// all images of a batch: bits = [count][rows][(w+7)/8]; each image's current buffer is picked from its own step counter
cudaError_t launch_mask_packed_batch(const void *u0, const void *u1, const CsvState *state, int f32, uint8_t *bits, int count,
                                     int rows, int w, int pitch, size_t plane_bytes, int invert, cudaStream_t s);
// border halo rows := copies of the first / last image row (planes of any element size; row_bytes multiple of 16)
cudaError_t launch_replicate_halo(void *base, size_t plane_bytes, size_t row_bytes, int nplanes, int rows, int top, int bottom,
                                  cudaStream_t s);
cudaError_t launch_checkerboard(double *u, const signed char *si, const signed char *sj, int row_lo, int rows, int w,
                                int pitch, cudaStream_t s);

// in_u8 / out_u8: the first step reads the uint8 image, the last one writes it (fused quantisation)
cudaError_t launch_pm_step(const PmArgs &A, bool in_u8, bool out_u8, bool strict, cudaStream_t s);
// two diffusion steps fused into one pass (temporal blocking), fp64 planes in and out; tiling g.ncb_pm2 / pm2_seg_rows
cudaError_t launch_pm2_step(const PmArgs &A, cudaStream_t s);
// P2P slab runs: one warp polls the neighbours' flags between two PM launches
cudaError_t launch_pm_wait(CommBox *box, unsigned int need, int has_up, int has_dn, cudaStream_t s);
cudaError_t launch_pm_quantise(const double *in, uint8_t *out, size_t n, cudaStream_t s);

// fp32 variant (f32_kernels.cu)
cudaError_t launch_csv_step_f32(const CsvArgs &A, cudaStream_t s);
cudaError_t launch_csv_init_f32(const CsvArgs &A, int final_mode, cudaStream_t s);
cudaError_t launch_pm_step_f32(const PmArgs &A, bool in_u8, bool out_u8, cudaStream_t s);
cudaError_t launch_convert_d2f(const double *in, float *out, size_t n, cudaStream_t s);
cudaError_t launch_convert_f2d(const float *in, double *out, size_t n, cudaStream_t s);
cudaError_t launch_mask_f32(const float *u, uint8_t *mask, int rows, int w, int pitch, int invert, cudaStream_t s);
cudaError_t launch_checkerboard_f32(float *u, const signed char *si, const signed char *sj, int row_lo, int rows, int w,
                                    int pitch, cudaStream_t s);
cudaError_t launch_quantise_f32(const float *in, uint8_t *out, size_t n, cudaStream_t s);

}  // namespace cvb
