// Deterministic fused reductions of the CSV kernels.
//
// Every CTA (one warp: one row segment x one column strip of one image) leaves one partial vector;
// the last CTA of each of the NGROUPS fixed row groups sums that group's partials in a fixed order;
// the last group finisher adds the NGROUPS group sums in index order and derives c1/c2, the norm and
// the stop flag.  Groups are keyed to GLOBAL segment indices, so a row-slab run over 1/2/4/8 GPUs
// (each rank owning NGROUPS/G consecutive groups, the group sums all-gathered) adds exactly the same
// numbers in exactly the same order as the single-GPU run.  No floating-point atomics anywhere.
#pragma once
#include "common.cuh"

namespace cvb {

__device__ __forceinline__ double ld_cg(const double *p) {
    double v;
    asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}

// Slots used by a kernel: [0, 1+NCH) + ACC_SQ (+ [ACC_I, ACC_I+NCH) for the init pass).
template <int NCH, bool INIT>
__device__ __forceinline__ bool slot_used(int v) {
    return v <= NCH || v == ACC_SQ || (INIT && v >= ACC_I && v < ACC_I + NCH);
}

// src/main.cpp:255-281 with H = a + 1/2: c1 = sum I*H / sum H, c2 = sum I*(1-H) / sum (1-H).
__device__ inline void region_means_from_sums(CsvState *st, const double *tot, int nch, double npix) {
    const double hs = 0.5 * npix + tot[ACC_A];
    const double gs = 0.5 * npix - tot[ACC_A];
    for (int k = 0; k < nch; ++k) {
        st->c1[k] = (0.5 * st->sumI[k] + tot[ACC_IA + k]) / hs;
        st->c2[k] = (0.5 * st->sumI[k] - tot[ACC_IA + k]) / gs;
    }
}

// One thread per image: add the group sums in index order and update the state.
// mode 0: after a step; 1: init with reset (stop condition, counters); 2: init without reset (means only)
__device__ inline void csv_finalize_image(const CsvArgs &A, int img, int mode) {
    const Geom &G = A.g;
    CsvState *st = A.state + img;
    double tot[NACC];
    for (int v = 0; v < NACC; ++v) tot[v] = 0.0;
    for (int grp = 0; grp < NGROUPS; ++grp)
        for (int v = 0; v < NACC; ++v) tot[v] += ld_cg(A.group_sums + ((size_t)grp * G.count + img) * NACC + v);
    for (int v = 0; v < NACC; ++v) st->sums[v] = tot[v];
    if (mode != 0)
        for (int k = 0; k < G.nch; ++k) st->sumI[k] = tot[ACC_I + k];
    if (mode == 1) {
        st->stop = A.tol * sqrt(tot[ACC_SQ]);  // src/main.cpp:959
        st->norm = __longlong_as_double(0x7ff8000000000000LL);
        st->done = 0;
        st->steps_done = 0;
    } else if (mode == 0) {
        const double nrm = sqrt(tot[ACC_SQ]);  // src/main.cpp:993
        st->norm = nrm;
        st->steps_done += 1;
        st->done = (nrm <= st->stop) ? 1 : 0;  // src/main.cpp:1000
    }
    region_means_from_sums(st, tot, G.nch, (double)G.h * (double)G.w);
}

// Called by the whole warp (= the whole CTA) after the row loop.  acc holds per-lane sums.
// No block-level barrier anywhere: a CTA is one warp, so warps never wait for one another and a finished warp frees
// its slot at once.  The last warp of a group adds the group's partial vectors in index order.
template <int NCH, bool INIT>
__device__ __forceinline__ void finish_tile(const CsvArgs &A, int img, int seg, int cb, int ncb, double (&acc)[NACC],
                                            int final_mode) {
    const Geom &G = A.g;
    const int lane = threadIdx.x & 31;
    CsvState *st = A.state + img;
#pragma unroll
    for (int v = 0; v < NACC; ++v)
        if (slot_used<NCH, INIT>(v)) {
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) acc[v] += __shfl_xor_sync(0xffffffffu, acc[v], off);
        }
    double *part = A.partials + (((size_t)img * G.nseg + seg) * ncb + cb) * NACC;
    const int gseg = G.seg0 + seg;
    const int grp = (int)(((long long)gseg * NGROUPS) / G.nseg_global);
    // local segments of this group
    const int sb = max(group_seg_begin(grp, G.nseg_global), G.seg0) - G.seg0;
    const int se = min(group_seg_begin(grp + 1, G.nseg_global), G.seg0 + G.nseg) - G.seg0;
    int last = 0;
    if (lane == 0) {
#pragma unroll
        for (int v = 0; v < NACC; ++v)
            if (slot_used<NCH, INIT>(v)) part[v] = acc[v];
        __threadfence();
        const unsigned int old = atomicAdd(&st->group_ticket[grp], 1u);
        last = (old == (unsigned int)((se - sb) * ncb) - 1u) ? 1 : 0;
    }
    last = __shfl_sync(0xffffffffu, last, 0);
    if (!last) return;
    __threadfence();
    // ---- this warp finishes group grp: fixed-order sum of its partial vectors (lane l takes l, l+32, ...)
    const int nvec = (se - sb) * ncb;
    const double *base = A.partials + (((size_t)img * G.nseg + sb) * ncb) * NACC;
    double sum[NACC];
#pragma unroll
    for (int v = 0; v < NACC; ++v) sum[v] = 0.0;
    for (int i = lane; i < nvec; i += 32) {
#pragma unroll
        for (int v = 0; v < NACC; ++v)
            if (slot_used<NCH, INIT>(v)) sum[v] += ld_cg(base + (size_t)i * NACC + v);
    }
#pragma unroll
    for (int v = 0; v < NACC; ++v)
        if (slot_used<NCH, INIT>(v)) {
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) sum[v] += __shfl_xor_sync(0xffffffffu, sum[v], off);
        }
    if (lane == 0) {
        double *gsum = A.group_sums + ((size_t)grp * G.count + img) * NACC;
#pragma unroll
        for (int v = 0; v < NACC; ++v) gsum[v] = slot_used<NCH, INIT>(v) ? sum[v] : 0.0;
        st->group_ticket[grp] = 0u;
        if (!A.multi_rank) {
            __threadfence();
            const unsigned int old = atomicAdd(&st->final_ticket, 1u);
            if (old == (unsigned int)A.ngroups_local - 1u) {
                __threadfence();
                st->final_ticket = 0u;
                csv_finalize_image(A, img, final_mode);
            }
        }
    }
}

}  // namespace cvb
