// Deterministic fused reductions of the CSV kernels, single- and multi-GPU.
//
// Every CTA (one warp: one row segment x one column strip of one image) leaves one partial vector; the last CTA of
// each of the NGROUPS fixed row groups sums that group's partials in a fixed order ("group sums"); the warp that
// completes the last LOCAL group either
//   * single GPU: FOLDS the reduction at once -- adds the NGROUPS group sums with a fixed shuffle tree (lane = group)
//     and updates CsvState (c1/c2, norm, stop flag, step counter) in device memory; or
//   * multi-GPU row slabs (P2P): copies this rank's group sums into every peer's group-sum array over NVLink and
//     bumps arrive[rank] in every peer's CommBox (st.release.sys after a system fence), then waits -- one warp, in
//     the tail of the same launch -- until all ranks have arrived and folds: the same numbers in the same order on
//     every rank => bit-identical c1/c2 and stop decision everywhere.  (csv_wait_fold remains for completeness.)
// Groups are keyed to GLOBAL segment indices, so a slab run over 1/2/4/8 GPUs adds exactly the same numbers in
// exactly the same order as the single-GPU run.  No floating-point atomics anywhere.
#pragma once
#include "comm.cuh"
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

// Whole warp: add the group sums of reduction `prod` of image `img` (lane g holds group g, fixed xor tree) and let
// lane 0 update the state.  mode 0: after a step; 1: init with reset (stop condition, counters); 2: means only.
// src/main.cpp:255-281 with H = a + 1/2: c1 = sum I*H / sum H, c2 = sum I*(1-H) / sum (1-H).
__device__ __forceinline__ void csv_fold(const CsvArgs &A, int img, int mode, unsigned int prod) {
    const Geom &G = A.g;
    const int lane = threadIdx.x & 31;
    static_assert(NGROUPS == 32, "one lane per reduction group");
    const double *gs = A.group_sums + ((size_t)(prod & 1u) * NGROUPS + lane) * G.count * NACC + (size_t)img * NACC;
    double tot[NACC];
#pragma unroll
    for (int v = 0; v < NACC; ++v) tot[v] = ld_cg(gs + v);
#pragma unroll
    for (int v = 0; v < NACC; ++v) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) tot[v] += __shfl_xor_sync(0xffffffffu, tot[v], off);
    }
    if (lane != 0) return;
    CsvState *st = A.state + img;
#pragma unroll
    for (int v = 0; v < NACC; ++v) st->sums[v] = tot[v];
    if (mode != 0) {
#pragma unroll
        for (int k = 0; k < MAX_CH; ++k) st->sumI[k] = tot[ACC_I + k];
    }
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
    const double npix = (double)G.h * (double)G.w;
    const double hs = 0.5 * npix + tot[ACC_A], gsum = 0.5 * npix - tot[ACC_A];
#pragma unroll
    for (int k = 0; k < MAX_CH; ++k) {
        st->c1[k] = (0.5 * st->sumI[k] + tot[ACC_IA + k]) / hs;
        st->c2[k] = (0.5 * st->sumI[k] - tot[ACC_IA + k]) / gsum;
    }
}

// Whole warp (P2P multi-GPU, between two launches): wait until every rank has pushed reduction `produced`, fold it.
__device__ __forceinline__ void csv_wait_fold(const CsvArgs &A) {
    CommBox *b = A.cv.box;
    const int lane = threadIdx.x & 31;
    const unsigned int prod = ld_relaxed(&b->produced);
    if (ld_relaxed(&b->finalized) == prod) return;  // nothing new (e.g. the image is frozen: launches were no-ops)
    if (lane < A.cv.nranks) spin_until(&b->arrive[lane], prod, b);
    __syncwarp();
    __threadfence_system();
    csv_fold(A, 0, b->pending_mode, prod);
    if (lane == 0) b->finalized = prod;
}

// Whole warp: fixed-order sum of `nvec` partial vectors (NACC doubles each, contiguous at `base`): lane l adds vectors
// l, l+32, l+64, ... in that order, then the lanes are combined with a fixed xor tree.  Four vectors per lane are
// requested before the first addition (the additions keep their order; only the loads overlap), so a few hundred
// vectors cost two or three L2 round trips instead of one per vector -- this sum sits in the tail of every step.
template <int NCH, bool INIT>
__device__ __forceinline__ void sum_vectors(const double *base, int nvec, int lane, double (&sum)[NACC]) {
    constexpr int NP = NACC / 2;
#pragma unroll
    for (int v = 0; v < NACC; ++v) sum[v] = 0.0;
    auto pair_used = [](int q) { return slot_used<NCH, INIT>(2 * q) || slot_used<NCH, INIT>(2 * q + 1); };
    int i = lane;
    for (; i + 96 < nvec; i += 128) {
        double2 t[4][NP];
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int q = 0; q < NP; ++q)
                if (pair_used(q)) t[k][q] = __ldcg(reinterpret_cast<const double2 *>(base + (size_t)(i + 32 * k) * NACC) + q);
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int q = 0; q < NP; ++q) {
                if (slot_used<NCH, INIT>(2 * q)) sum[2 * q] += t[k][q].x;
                if (slot_used<NCH, INIT>(2 * q + 1)) sum[2 * q + 1] += t[k][q].y;
            }
    }
    for (; i < nvec; i += 32) {
        double2 t[NP];
#pragma unroll
        for (int q = 0; q < NP; ++q)
            if (pair_used(q)) t[q] = __ldcg(reinterpret_cast<const double2 *>(base + (size_t)i * NACC) + q);
#pragma unroll
        for (int q = 0; q < NP; ++q) {
            if (slot_used<NCH, INIT>(2 * q)) sum[2 * q] += t[q].x;
            if (slot_used<NCH, INIT>(2 * q + 1)) sum[2 * q + 1] += t[q].y;
        }
    }
#pragma unroll
    for (int v = 0; v < NACC; ++v)
        if (slot_used<NCH, INIT>(v)) {
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) sum[v] += __shfl_xor_sync(0xffffffffu, sum[v], off);
        }
}

// Whole warp, after the row loop.  acc holds per-lane sums.
// Tree: CTA partial vectors -> (large jobs only: A.seg_level) one sum per row segment, added by the last CTA of the
// segment -> one sum per reduction group, added by the last arrival of the group -> fold.  Every level adds its
// inputs in index order, so the result depends on the tiling (tile rows, strips) and on nothing else: not on the
// order in which CTAs finish, not on the number of GPUs.  The segment level keeps the chain that is exposed at the
// end of a step short: without it the last group's finisher adds (segments per group) x (strips) vectors alone.
template <int NCH, bool INIT>
__device__ __forceinline__ void finish_tile(const CsvArgs &A, int img, int seg, int cb, int ncb, double (&acc)[NACC],
                                            int final_mode, bool pushed_rows = false) {
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
    const bool seg_level = A.seg_level != 0;
    int last = 0;
    if (pushed_rows) __threadfence_system();  // this warp stored boundary rows into a neighbour's halo
    __syncwarp();
    if (lane == 0) {
#pragma unroll
        for (int v = 0; v < NACC; ++v)
            if (slot_used<NCH, INIT>(v)) part[v] = acc[v];
        __threadfence();
        if (seg_level) {
            const unsigned int old = atomicAdd(&A.seg_ticket[(size_t)img * G.nseg + seg], 1u);
            last = (old == (unsigned int)ncb - 1u) ? 1 : 0;
        } else {
            const unsigned int old = atomicAdd(&st->group_ticket[grp], 1u);
            last = (old == (unsigned int)((se - sb) * ncb) - 1u) ? 1 : 0;
        }
    }
    last = __shfl_sync(0xffffffffu, last, 0);
    if (!last) return;
    __threadfence();
    double sum[NACC];
    if (seg_level) {
        // ---- this warp finishes segment seg: fixed-order sum of its strips' partial vectors
        sum_vectors<NCH, INIT>(A.partials + (((size_t)img * G.nseg + seg) * ncb) * NACC, ncb, lane, sum);
        if (lane == 0) {
            double *ss = A.seg_sums + ((size_t)img * G.nseg + seg) * NACC;
#pragma unroll
            for (int v = 0; v < NACC; ++v)
                if (slot_used<NCH, INIT>(v)) ss[v] = sum[v];
            A.seg_ticket[(size_t)img * G.nseg + seg] = 0u;
            __threadfence();
            const unsigned int old = atomicAdd(&st->group_ticket[grp], 1u);
            last = (old == (unsigned int)(se - sb) - 1u) ? 1 : 0;
        }
        last = __shfl_sync(0xffffffffu, last, 0);
        if (!last) return;
        __threadfence();
        // ---- ... and group grp: fixed-order sum of its segments' sums
        sum_vectors<NCH, INIT>(A.seg_sums + ((size_t)img * G.nseg + sb) * NACC, se - sb, lane, sum);
    } else {
        // ---- this warp finishes group grp: fixed-order sum of its partial vectors
        sum_vectors<NCH, INIT>(A.partials + (((size_t)img * G.nseg + sb) * ncb) * NACC, (se - sb) * ncb, lane, sum);
    }
    const bool p2p = A.cv.p2p != 0;
    // multi-GPU: this reduction's number; its parity picks the group-sum buffer (double buffering against slow peers)
    const unsigned int prod = p2p ? ld_relaxed(&A.cv.box->produced) + 1u : 0u;
    double *gbase = A.group_sums + (size_t)(prod & 1u) * NGROUPS * G.count * NACC;
    int all_done = 0;
    if (lane == 0) {
        double *gsum = gbase + ((size_t)grp * G.count + img) * NACC;
#pragma unroll
        for (int v = 0; v < NACC; ++v) gsum[v] = slot_used<NCH, INIT>(v) ? sum[v] : 0.0;
        st->group_ticket[grp] = 0u;
        if (!A.multi_rank || p2p) {
            __threadfence();
            const unsigned int old = atomicAdd(&st->final_ticket, 1u);
            if (old == (unsigned int)A.ngroups_local - 1u) {
                st->final_ticket = 0u;
                all_done = 1;
            }
        }
    }
    all_done = __shfl_sync(0xffffffffu, all_done, 0);
    if (!all_done) return;
    // ---- every local group of this image is summed
    __threadfence();
    if (!p2p) {
        csv_fold(A, img, final_mode, 0u);
        return;
    }
    // slab session (count == 1): this rank's group sums go to every peer (the warp copies them together: one L2 read per
    // value, nranks - 1 posted stores over NVLink), then lane p raises peer p's flag with st.release.sys -- the release,
    // after the warp barrier, orders every lane's stores before the flag: ONE system-scope round trip, no separate fence.
    const size_t first = (size_t)A.group_lo * G.count * NACC;
    const int nval2 = (A.group_hi - A.group_lo) * G.count * NACC / 2;  // in double2 (NACC is even)
    const size_t poff = (size_t)(prod & 1u) * NGROUPS * G.count * NACC + first;
    const double2 *src = reinterpret_cast<const double2 *>(gbase + first);
    for (int i = lane; i < nval2; i += 64) {
        const bool two = i + 32 < nval2;
        const double2 v0 = __ldcg(src + i), v1 = two ? __ldcg(src + i + 32) : make_double2(0.0, 0.0);
        for (int p = 0; p < A.cv.nranks; ++p)
            if (p != A.cv.rank) {
                double2 *dst = reinterpret_cast<double2 *>(A.cv.peer_group[p] + poff);
                dst[i] = v0;
                if (two) dst[i + 32] = v1;
            }
    }
    if (lane == 0) {
        A.cv.box->pending_mode = final_mode;
        *reinterpret_cast<volatile unsigned int *>(&A.cv.box->produced) = prod;
    }
    __syncwarp();
    const unsigned long long t_wait = global_ns();
    if (lane < A.cv.nranks) {
        st_release_sys(&A.cv.peer_box[lane]->arrive[A.cv.rank], prod);
        // ... and fold right here, in the tail of the same launch: wait until every rank has arrived (the other SMs of
        // this GPU are idle by now; the peers only need THEIR OWN launch to finish, so nobody waits in a circle), then
        // add the group sums -- the same numbers in the same order on every rank.  No extra launch per step.
        spin_until(&A.cv.box->arrive[lane], prod, A.cv.box);
    }
    __syncwarp();  // lane p has acquired peer p's flag: with the warp barrier every lane may read every peer's sums
    csv_fold(A, 0, final_mode, prod);
    if (lane == 0) {
        A.cv.box->finalized = prod;
        A.cv.box->wait_ns += global_ns() - t_wait;
        A.cv.box->wait_count += 1u;
    }
}

}  // namespace cvb
