// Chan-Sandberg-Vese level-set kernels for sm_100a.
//
// csv_step_kernel fuses one whole iteration of the reference's time-step loop (src/main.cpp:963-1001)
// into a single pass over HBM: curvature (:342-375), data term (:299-312, :968-980), the addWeighted
// combine (:985), the regularised delta (:204-210 through ParallelPixelFunction, :988-992), the norm
// (:993), the update (:994) AND the sums that give the next step's region means c1/c2 (:255-281,
// :973-974) of the updated level set.  Algorithmic traffic: read u 8 B + write u 8 B + N bytes of image
// per pixel per step (the reference moves ~0.9 kB).
//
// Mapping: a warp (= one CTA) marches down a strip of 64 columns, each lane owning two adjacent columns; ny(i-1)
// and u(i) - u(i-1) are carried in registers from the previous row.  Lane 0 is a halo lane (it supplies nx of the
// column left of the strip), so a strip owns 62 columns.  Production path (csv_rows_ring): rows stream
// HBM -> shared memory through a cp.async ring 7 rows ahead of the row being computed, east / west neighbours are
// read from the ring; a CTA may march through several segments and delivers the fused sums per segment
// (Geom::seg_mult).  Build variant -DCSV_TMA (csv_rows_tma): the same loop fed by cp.async.bulk.tensor + mbarrier --
// measured equal (profiles/README.md), not the default.  Generic path (strict math, curvature-only mode): explicit
// clamps, register prefetch CSV_D rows ahead and L2 prefetch CSV_PF rows ahead.
#include <limits.h>
#include <string.h>

#include "async_copy.cuh"
#include "common.cuh"
#include "kernels.h"
#include "math.cuh"
#include "reduce.cuh"

namespace cvb {

#ifndef CSV_D
#define CSV_D 2
#endif
#ifndef CSV_PF
#define CSV_PF 8
#endif
#ifndef CSV_MIN_CTAS
#define CSV_MIN_CTAS 16  // one-warp CTAs: 16 resident warps per SM = 128 registers per thread (more did not help)
#endif
#ifndef CSV_NO_UNIFORM_COEF
#define CSV_UNIFORM_COEF 1
#endif

enum { MODE_STEP = 0, MODE_KAPPA = 1 };


// P2P multi-GPU: rows that are a neighbouring slab's halo are also stored straight into that neighbour's buffer.
__device__ __forceinline__ bool push_halo_rows(const CsvArgs &A, int buf, int i, int a, int lane, double v0, double v1) {
    const Geom &G = A.g;
    bool pushed = false;
    if (i < G.row_lo + HALO && A.cv.up_u[buf] != nullptr) {
        double *d = A.cv.up_u[buf] + (size_t)(HALO + A.cv.up_rows + (i - G.row_lo)) * G.pitch + a;
        if (lane >= 1 && a + 1 < G.w)
            *reinterpret_cast<double2 *>(d) = make_double2(v0, v1);
        else if (lane >= 1 && a < G.w)
            *d = v0;
        pushed = true;
    }
    if (i >= G.row_hi - HALO && A.cv.dn_u[buf] != nullptr) {
        double *d = A.cv.dn_u[buf] + (size_t)(i - (G.row_hi - HALO)) * G.pitch + a;
        if (lane >= 1 && a + 1 < G.w)
            *reinterpret_cast<double2 *>(d) = make_double2(v0, v1);
        else if (lane >= 1 && a < G.w)
            *d = v0;
        pushed = true;
    }
    return pushed;
}

// Read back the boundary rows this warp just wrote (L2 hits) and store them straight into the neighbour's buffer --
// outside the hot loop, and out of line so that the hot loop's register allocation does not see it.
__device__ __noinline__ bool push_boundary_rows(const CsvArgs &A, const double *uout, int buf, int ra, int rb, int a, int lane) {
    const Geom &G = A.g;
    bool pushed = false;
    const int lo_end = min(rb, G.row_lo + HALO), hi_beg = max(ra, G.row_hi - HALO);
    for (int i = ra; i < rb; ++i) {
        if (i >= lo_end && i < hi_beg) continue;
        const double2 v = __ldcg(reinterpret_cast<const double2 *>(uout + (size_t)(i - G.row_lo + HALO) * G.pitch + a));
        pushed |= push_halo_rows(A, buf, i, a, lane, v.x, v.y);
    }
    return pushed;
}

// BORDER_REPLICATE in i for the fast path: the halo rows above row 0 and below row h-1 of the buffer just written
// get copies of those rows (each lane copies the two columns it owns; read back from L2, outside the hot loop).
__device__ __noinline__ void replicate_border_rows(double *uout, const Geom &G, int ra, int rb, int a) {
    if (ra == 0) {
        const double2 v = __ldcg(reinterpret_cast<const double2 *>(uout + (size_t)(0 - G.row_lo + HALO) * G.pitch + a));
        for (int k = 0; k < HALO; ++k) *reinterpret_cast<double2 *>(uout + (size_t)k * G.pitch + a) = v;
    }
    if (rb == G.h) {
        const size_t last = (size_t)(G.h - 1 - G.row_lo + HALO);
        const double2 v = __ldcg(reinterpret_cast<const double2 *>(uout + last * G.pitch + a));
        for (int k = 1; k <= HALO; ++k) *reinterpret_cast<double2 *>(uout + (last + k) * G.pitch + a) = v;
    }
}

// Coefficients of one step, derived from the region means once per thread.
template <int NCH>
struct StepCoef {
    double cA[NCH], cB[NCH];  // per-channel data term as a quadratic in I: A_k I^2 + B_k I
    double q0;                // constant part: gamma' + sum_k C_k
    double alphap;            // mu*dt * eps/pi
    double eps2, inv_eps;
    bool linear;              // all A_k == 0 (lambda1 == lambda2, the reference's default): the data term is linear in I
};

// ---- fast path with an asynchronous shared-memory row ring ---------------------------------------------
// Rows travel HBM -> shared memory with cp.async (LDGSTS, 16-byte chunks, L2-only) RING_NS - 1 rows ahead of the row
// being computed: no registers are held by data in flight (ptxas sank the register prefetch of the earlier versions
// next to its first use to stay within 128 registers, profiles/README.md r1e), the prefetch distance is deep enough for
// the HBM latency-bandwidth product (7 rows x 0.8 KB x 16 warps per SM), and the east / west neighbours are read from
// the ring instead of being shuffled.
// Slot of row k (k = i - ra; RING_SLOT bytes at (k % RING_NS) * RING_SLOT):
//   [16 + 16*j, +16)          chunk j = 0..32 of the u row: columns cs-2+2j, cs-1+2j (lane j's two columns; chunk 32
//                             holds lane 31's east neighbour)
//   [RING_IMG + 80*c, +80)    image row of channel c from the 16-byte aligned column (cs-2) & ~15
// Per row: two LDGSTS warp-instructions (lanes 0..31: u chunks 0..31; lanes 0..5*NCH-1: image chunks, lane 15+: u
// chunk 32), one commit, one wait.  Rows past the end of the segment are fetched too and never used (TAIL_ROWS).
// With RING_NS = 8 and the loop unrolled 4x the slot offsets are compile-time constants on top of one toggling base.
constexpr int RING_NS = 8;
constexpr int RING_SLOT = 1024;
constexpr int RING_U = 16;
constexpr int RING_IMG = RING_U + 33 * 16;
constexpr int RING_BYTES = RING_NS * RING_SLOT;
static_assert(RING_IMG + 80 * MAX_CH <= RING_SLOT, "slot too small");
static_assert(RING_NS - 1 <= TAIL_ROWS, "tail padding too small for the ring");

// A CTA that marches through several segments (Geom::seg_mult) delivers the sums of a finished segment from inside its
// row loop: rare (once per seg_rows rows), out of line so that the hot loop's register allocation does not see it.
// sums of lane 0 (a halo lane) are dropped, as at the end of the row loop.
struct SegCursor {
    const CsvArgs *A;
    int img, seg, cb;   // seg: the local segment the rows being computed belong to
    int next;           // first row (relative to the CTA's first row) of the next segment; INT_MAX: no further segment
};
template <int NCH>
__device__ __noinline__ void csv_flush_segment(const CsvArgs &A, int img, int seg, int cb, double a, double s, double i0, double i1,
                                               double i2) {
    double acc[NACC];
#pragma unroll
    for (int v = 0; v < NACC; ++v) acc[v] = 0.0;
    if (threadIdx.x & 31) {
        acc[ACC_A] = a;
        acc[ACC_SQ] = s;
        acc[ACC_IA] = i0;
        if (NCH > 1) acc[ACC_IA + 1] = i1;
        if (NCH > 2) acc[ACC_IA + 2] = i2;
    }
    finish_tile<NCH, false>(A, img, seg, cb, A.g.ncb_csv, acc, 0, false);
}
#define CSV_SEGMENT_BOUNDARY(r)                                                                                         \
    if ((r) == cur.next && (r) < n) { /* the sums of the CTA's LAST segment are delivered after the loop */            \
        csv_flush_segment<NCH>(*cur.A, cur.img, cur.seg, cur.cb, accA, accS, accI[0], NCH > 1 ? accI[NCH > 1 ? 1 : 0] : 0.0, \
                               NCH > 2 ? accI[NCH > 2 ? 2 : 0] : 0.0);                                                  \
        ++cur.seg;                                                                                                     \
        cur.next += G.seg_rows;                                                                                        \
        accA = accS = 0.0;                                                                                             \
        _Pragma("unroll") for (int c = 0; c < NCH; ++c) accI[c] = 0.0;                                                 \
    }

template <int NCH, bool EDGE, bool LINEAR>
__device__ __forceinline__ void csv_rows_ring(const double *__restrict__ uin, double *__restrict__ uout,
                                              const uint8_t *__restrict__ im, const Geom &G, const StepCoef<NCH> &K,
                                              const double2 *s_tab, unsigned char *ring, int ra, int rb, int cs, int lane,
                                              double (&acc)[NACC], SegCursor &cur) {
    const int w = G.w;
    const int a = cs - 2 + 2 * lane;
    const bool first = EDGE && a == 0, last0 = EDGE && a == w - 1, last1 = EDGE && a + 1 == w - 1;
    const bool v0 = lane >= 1 && (!EDGE || a < w), v1 = lane >= 1 && (!EDGE || a + 1 < w);
    const unsigned int p2 = (unsigned int)G.pitch >> 1;  // row pitch in 16-byte (u) / 2-byte (image) units
    const size_t pe = (size_t)G.plane_elems;
    // 16-byte index of (row ra, column a) in a u plane; 2 * o = byte index of the same position in an image plane
    unsigned int o = (unsigned int)(ra - G.row_lo + HALO) * p2 + (unsigned int)(a >> 1);
    const double2 *bu = reinterpret_cast<const double2 *>(uin);
    double2 *bo = reinterpret_cast<double2 *>(uout);

    // ---- the two copy instructions of a row: per-lane source = base + o * scale, per-lane slot offset
    const int s_al = (cs - 2) & ~15;            // first column of the image chunks
    const int dsh = (cs - 2) - s_al;            // byte offset of column cs-2 inside the image strip
    const bool ok1 = !EDGE || (a >= 0 && a < G.pitch);
    const unsigned int dst1 = RING_U + 16 * lane;
    const char *base2 = reinterpret_cast<const char *>(uin) + 16 * (32 - lane);  // u chunk 32 (lanes >= 5*NCH)
    unsigned int scale2 = 16, dst2 = RING_U + 16 * 32;
    bool ok2 = lane == 15 && (!EDGE || a + 2 * (32 - lane) < G.pitch);
    if (lane < 5 * NCH) {
        const int c = lane / 5, q = lane % 5, col = s_al + 16 * q;
        base2 = reinterpret_cast<const char *>(im) + (size_t)c * pe + (col - a);
        scale2 = 2;
        dst2 = RING_IMG + 80 * c + 16 * q;
        ok2 = !EDGE || (col >= 0 && col < G.pitch);
    }
    const unsigned int ring_s = (unsigned int)__cvta_generic_to_shared(ring);
    auto issue = [&](unsigned int orow, unsigned int slot_off) {  // orow: index of (row, column a)
        if (ok1) cp_async16(ring_s + slot_off + dst1, bu + orow);
        if (ok2) cp_async16(ring_s + slot_off + dst2, base2 + (size_t)orow * scale2);
        cp_async_commit();
    };
    // prologue: rows ra .. ra+NS-2 into slots 0 .. NS-2
#pragma unroll
    for (int k = 0; k < RING_NS - 1; ++k) issue(o + k * p2, k * RING_SLOT);

    // rows ra-2, ra-1 (own columns only) straight from global memory
    const double2 R0 = ok1 ? __ldg(bu + o - 2 * p2) : make_double2(0.0, 0.0);
    const double2 R1 = ok1 ? __ldg(bu + o - p2) : make_double2(0.0, 0.0);
    const unsigned char *my16 = ring + 16 * lane;        // + slot: own chunk is at RING_U, west at RING_U - 8, east + 16
    const unsigned char *my2 = ring + dsh + 2 * lane;    // + slot + RING_IMG + 80 c: own two image bytes
    cp_async_wait<RING_NS - 3>();  // rows ra and ra+1 have landed
    __syncwarp();
    double2 C = *reinterpret_cast<const double2 *>(my16 + RING_U);
    double CW = *reinterpret_cast<const double *>(my16 + RING_U - 8);
    double CE = *reinterpret_cast<const double *>(my16 + RING_U + 16);
    double dN0 = C.x - R1.x, dN1 = C.y - R1.y;
    double nyp0 = normal_component<false>(dN0, dN0 + (R1.x - R0.x));
    double nyp1 = normal_component<false>(dN1, dN1 + (R1.y - R0.y));
    if (ra == 0) {
        // image top: the halo rows hold copies of row 0 (BORDER_REPLICATE), so dN = 0; the y-term of kappa must vanish
        // in row 0 (ny(-1) := ny(0), src/main.cpp:372): start from the very value the loop will compute for ny(0)
        const double2 q = *reinterpret_cast<const double2 *>(my16 + RING_SLOT + RING_U);
        nyp0 = normal_component<false>(q.x - C.x, (q.x - C.x) + dN0);
        nyp1 = normal_component<false>(q.y - C.y, (q.y - C.y) + dN1);
    }

    double accA = 0.0, accS = 0.0, accI[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) accI[c] = 0.0;
    const int n = rb - ra;
    o += (RING_NS - 1) * p2;  // the row being fetched; the row being computed is RING_NS - 1 rows behind
    const double2 *bo_c = bo - (size_t)(RING_NS - 1) * p2;

    // one row: s_wr = slot of row i-1 (free, receives row i+NS-1), s_img = slot of row i, s_u = slot of row i+1
    auto row = [&](unsigned int s_wr, unsigned int s_img, unsigned int s_u) {
        issue(o, s_wr);
        cp_async_wait<RING_NS - 2>();  // row i+1 has landed
        __syncwarp();
        const double2 S = *reinterpret_cast<const double2 *>(my16 + s_u + RING_U);
        const double SW = *reinterpret_cast<const double *>(my16 + s_u + RING_U - 8);
        const double SE = *reinterpret_cast<const double *>(my16 + s_u + RING_U + 16);
        unsigned int Ib[NCH];
#pragma unroll
        for (int c = 0; c < NCH; ++c) Ib[c] = *reinterpret_cast<const unsigned short *>(my2 + s_img + RING_IMG + 80 * c);

        // curvature (:342-375)
        const double upy0 = S.x - C.x, upy1 = S.y - C.y;
        const double ny0 = normal_component<false>(upy0, upy0 + dN0);
        const double ny1 = normal_component<false>(upy1, upy1 + dN1);
        double Wn = CW, E2 = CE, E0 = C.y;
        if (EDGE) {
            Wn = first ? C.x : Wn;
            E0 = last0 ? C.x : C.y;
            E2 = last1 ? C.y : E2;
        }
        const double nx0 = normal_component<false>(E0 - C.x, E0 - Wn);
        const double nx1 = normal_component<false>(E2 - C.y, E2 - C.x);
        const double nxw = __shfl_up_sync(0xffffffffu, nx1, 1);
        double kx0 = nx0 - nxw;
        if (EDGE) kx0 = first ? 0.0 : kx0;
        const double kap0 = kx0 + (ny0 - nyp0);
        const double kap1 = (nx1 - nx0) + (ny1 - nyp1);
        // data term + combine (:968-985), delta (:988-992), update (:994)
        double I0[NCH], I1[NCH];
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            I0[c] = u8_to_double(Ib[c] & 0xffu);
            I1[c] = u8_to_double(Ib[c] >> 8);
        }
        double t0 = K.q0, t1 = K.q0;
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            if (LINEAR) {
                t0 = fma(K.cB[c], I0[c], t0);
                t1 = fma(K.cB[c], I1[c], t1);
            } else {
                t0 = fma(fma(K.cA[c], I0[c], K.cB[c]), I0[c], t0);
                t1 = fma(fma(K.cA[c], I1[c], K.cB[c]), I1[c], t1);
            }
        }
        t0 = fma(kap0, K.alphap, t0);
        t1 = fma(kap1, K.alphap, t1);
        // one reciprocal for both pixels: 1/s0 = s1/(s0*s1), 1/s1 = s0/(s0*s1)
        const double s0 = fma(C.x, C.x, K.eps2), s1 = fma(C.y, C.y, K.eps2);
        const double rs = fast_rcp(s0 * s1);
        const double du0 = t0 * (rs * s1);
        const double du1 = t1 * (rs * s0);
        const double un0 = C.x + du0, un1 = C.y + du1;
        double2 *po = const_cast<double2 *>(bo_c) + o;
        if (EDGE) {
            if (v1)
                *po = make_double2(un0, un1);
            else if (v0)
                *reinterpret_cast<double *>(po) = un0;
        } else if (lane) {
            *po = make_double2(un0, un1);
        }
        o += p2;

        // sums of the updated level set and of du^2 (lane 0 is a halo lane: its sums are dropped at the end)
        double a0, a1;
        atan_over_pi2(un0 * K.inv_eps, un1 * K.inv_eps, s_tab, a0, a1);
        double dq0 = du0, dq1 = du1;
        if (EDGE) {  // columns beyond the image do not count
            a0 = v0 ? a0 : 0.0;
            a1 = v1 ? a1 : 0.0;
            dq0 = v0 ? du0 : 0.0;
            dq1 = v1 ? du1 : 0.0;
        }
        accA += a0;
        accA += a1;
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            accI[c] = fma(I0[c], a0, accI[c]);
            accI[c] = fma(I1[c], a1, accI[c]);
        }
        accS = fma(dq0, dq0, accS);
        accS = fma(dq1, dq1, accS);
        // next row
        dN0 = upy0;
        dN1 = upy1;
        nyp0 = ny0;
        nyp1 = ny1;
        C = S;
        CW = SW;
        CE = SE;
    };

    int r = 0;
    unsigned int tog = 0;  // offset of slot 0 or slot 4: the slot of the first row of a group of four
#pragma unroll 1
    for (; r + 4 <= n; r += 4) {
        CSV_SEGMENT_BOUNDARY(r)  // seg_rows is a multiple of 4 whenever seg_mult > 1
        const unsigned int t2 = tog ^ (4 * RING_SLOT);
        row(t2 + 3 * RING_SLOT, tog, tog + RING_SLOT);
        row(tog, tog + RING_SLOT, tog + 2 * RING_SLOT);
        row(tog + RING_SLOT, tog + 2 * RING_SLOT, tog + 3 * RING_SLOT);
        row(tog + 2 * RING_SLOT, tog + 3 * RING_SLOT, t2);
        tog = t2;
    }
    CSV_SEGMENT_BOUNDARY(r)
#pragma unroll 1
    for (; r < n; ++r) {
        const unsigned int k = (unsigned int)r;
        row(((k + RING_NS - 1) % RING_NS) * RING_SLOT, (k % RING_NS) * RING_SLOT, ((k + 1) % RING_NS) * RING_SLOT);
    }
    cp_async_wait<0>();  // nothing may land in the ring after the CTA has gone
    if (lane) {
        acc[ACC_A] = accA;
        acc[ACC_SQ] = accS;
#pragma unroll
        for (int c = 0; c < NCH; ++c) acc[ACC_IA + c] = accI[c];
    }
}

#ifdef CSV_TMA
// ---- build variant: the same row loop fed by the tensor memory accelerator ---------------------------------------------
// One cp.async.bulk.tensor.2d per TMA_R-row x 66-column tile of u and one .3d per tile of the image channels, issued by
// one lane, completion on an mbarrier per stage; TMA_NST stages.  A tile is exactly the four rows of one unrolled loop
// iteration, so per iteration the warp executes one barrier wait and (lane 0) one expect_tx + two bulk copies instead of
// 4 x (2 LDGSTS + commit + wait) and their address arithmetic.  Boxes that reach outside the tensor (strips at the left /
// right border, rows past the last plane) are zero-filled by the hardware.
constexpr int TMA_R = 4;
constexpr int TMA_NST = 3;
constexpr int TMA_UROW = 33 * 16;                        // 66 doubles
constexpr int TMA_UB = (TMA_R * TMA_UROW + 127) / 128 * 128;
constexpr int TMA_IROW = 80;
constexpr int TMA_IB = (MAX_CH * TMA_R * TMA_IROW + 127) / 128 * 128;
constexpr int TMA_STAGE = TMA_UB + TMA_IB;
constexpr int TMA_LEAD = 128;                            // mbarriers + the 8 bytes lane 0 reads west of a tile
constexpr int TMA_BYTES = TMA_LEAD + TMA_NST * TMA_STAGE;

__device__ __forceinline__ void mbar_init(unsigned int bar, unsigned int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned int bar, unsigned int bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned int bar, unsigned int parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "CVB_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra CVB_DONE;\n\t"
        "bra CVB_WAIT;\n\t"
        "CVB_DONE:\n\t"
        "}" ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(unsigned int dst, const CUtensorMap *m, unsigned int bar, int x, int y) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
                 "l"(m), "r"(bar), "r"(x), "r"(y)
                 : "memory");
}
__device__ __forceinline__ void tma_load_3d(unsigned int dst, const CUtensorMap *m, unsigned int bar, int x, int y, int z) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
                 "l"(m), "r"(bar), "r"(x), "r"(y), "r"(z)
                 : "memory");
}

template <int NCH, bool EDGE, bool LINEAR>
__device__ __forceinline__ void csv_rows_tma(const double *__restrict__ uin, double *__restrict__ uout, const CUtensorMap *tm_u,
                                             const CUtensorMap *tm_img, int img, const Geom &G, const StepCoef<NCH> &K,
                                             const double2 *s_tab, unsigned char *ring, int ra, int rb, int cs, int lane,
                                             double (&acc)[NACC], SegCursor &cur) {
    const int w = G.w;
    const int a = cs - 2 + 2 * lane;
    const bool first = EDGE && a == 0, last0 = EDGE && a == w - 1, last1 = EDGE && a + 1 == w - 1;
    const bool v0 = lane >= 1 && (!EDGE || a < w), v1 = lane >= 1 && (!EDGE || a + 1 < w);
    const unsigned int p2 = (unsigned int)G.pitch >> 1;
    unsigned int o = (unsigned int)(ra - G.row_lo + HALO) * p2 + (unsigned int)(a >> 1);
    const double2 *bu = reinterpret_cast<const double2 *>(uin);
    double2 *bo = reinterpret_cast<double2 *>(uout);
    const int s_al = (cs - 2) & ~15;  // first column of the image box
    const int dsh = (cs - 2) - s_al;
    const bool ok1 = !EDGE || (a >= 0 && a < G.pitch);
    const int n = rb - ra;
    const int ntiles = (n + 1 + TMA_R - 1) / TMA_R;  // rows ra .. rb (the row below the segment is the last south row)
    const unsigned int ring_s = (unsigned int)__cvta_generic_to_shared(ring);
    const int urow0 = img * G.rows_alloc + (ra - G.row_lo + HALO), irow0 = ra - G.row_lo + HALO, z0 = img * G.nch;
    constexpr unsigned int kBytes = TMA_R * TMA_UROW + NCH * TMA_R * TMA_IROW;
    auto issue_tile = [&](int t) {  // one lane
        const unsigned int st = ring_s + TMA_LEAD + (unsigned int)(t % TMA_NST) * TMA_STAGE, bar = ring_s + 8u * (unsigned int)(t % TMA_NST);
        mbar_expect_tx(bar, kBytes);
        tma_load_2d(st, tm_u, bar, cs - 2, urow0 + TMA_R * t);
        tma_load_3d(st + TMA_UB, tm_img, bar, s_al, irow0 + TMA_R * t, z0);
    };
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < TMA_NST; ++k) mbar_init(ring_s + 8u * k, 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < TMA_NST; ++k)
            if (k < ntiles) issue_tile(k);
    }
    // rows ra-2, ra-1 (own columns only) straight from global memory
    const double2 R0 = ok1 ? __ldg(bu + o - 2 * p2) : make_double2(0.0, 0.0);
    const double2 R1 = ok1 ? __ldg(bu + o - p2) : make_double2(0.0, 0.0);
    const unsigned char *mu = ring + TMA_LEAD + 16 * lane;                   // + stage + row * TMA_UROW: own chunk
    const unsigned char *mi = ring + TMA_LEAD + TMA_UB + dsh + 2 * lane;     // + stage + c * TMA_R * TMA_IROW + row * TMA_IROW
    mbar_wait(ring_s, 0u);  // tile 0
    double2 C = *reinterpret_cast<const double2 *>(mu);
    double CW = *reinterpret_cast<const double *>(mu - 8);
    double CE = *reinterpret_cast<const double *>(mu + 16);
    double dN0 = C.x - R1.x, dN1 = C.y - R1.y;
    double nyp0 = normal_component<false>(dN0, dN0 + (R1.x - R0.x));
    double nyp1 = normal_component<false>(dN1, dN1 + (R1.y - R0.y));
    if (ra == 0) {  // see csv_rows_ring
        const double2 q = *reinterpret_cast<const double2 *>(mu + TMA_UROW);
        nyp0 = normal_component<false>(q.x - C.x, (q.x - C.x) + dN0);
        nyp1 = normal_component<false>(q.y - C.y, (q.y - C.y) + dN1);
    }
    double accA = 0.0, accS = 0.0, accI[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) accI[c] = 0.0;

    // one row: pu = own chunk of row i+1, pi = own image bytes of row i (channel 0; channel stride TMA_R * TMA_IROW)
    auto row = [&](const unsigned char *pu, const unsigned char *pi) {
        const double2 S = *reinterpret_cast<const double2 *>(pu);
        const double SW = *reinterpret_cast<const double *>(pu - 8);
        const double SE = *reinterpret_cast<const double *>(pu + 16);
        unsigned int Ib[NCH];
#pragma unroll
        for (int c = 0; c < NCH; ++c) Ib[c] = *reinterpret_cast<const unsigned short *>(pi + c * (TMA_R * TMA_IROW));
        const double upy0 = S.x - C.x, upy1 = S.y - C.y;
        const double ny0 = normal_component<false>(upy0, upy0 + dN0);
        const double ny1 = normal_component<false>(upy1, upy1 + dN1);
        double Wn = CW, E2 = CE, E0 = C.y;
        if (EDGE) {
            Wn = first ? C.x : Wn;
            E0 = last0 ? C.x : C.y;
            E2 = last1 ? C.y : E2;
        }
        const double nx0 = normal_component<false>(E0 - C.x, E0 - Wn);
        const double nx1 = normal_component<false>(E2 - C.y, E2 - C.x);
        const double nxw = __shfl_up_sync(0xffffffffu, nx1, 1);
        double kx0 = nx0 - nxw;
        if (EDGE) kx0 = first ? 0.0 : kx0;
        const double kap0 = kx0 + (ny0 - nyp0);
        const double kap1 = (nx1 - nx0) + (ny1 - nyp1);
        double I0[NCH], I1[NCH];
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            I0[c] = u8_to_double(Ib[c] & 0xffu);
            I1[c] = u8_to_double(Ib[c] >> 8);
        }
        double t0 = K.q0, t1 = K.q0;
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            if (LINEAR) {
                t0 = fma(K.cB[c], I0[c], t0);
                t1 = fma(K.cB[c], I1[c], t1);
            } else {
                t0 = fma(fma(K.cA[c], I0[c], K.cB[c]), I0[c], t0);
                t1 = fma(fma(K.cA[c], I1[c], K.cB[c]), I1[c], t1);
            }
        }
        t0 = fma(kap0, K.alphap, t0);
        t1 = fma(kap1, K.alphap, t1);
        const double s0 = fma(C.x, C.x, K.eps2), s1 = fma(C.y, C.y, K.eps2);
        const double rs = fast_rcp(s0 * s1);
        const double du0 = t0 * (rs * s1);
        const double du1 = t1 * (rs * s0);
        const double un0 = C.x + du0, un1 = C.y + du1;
        double2 *po = bo + o;
        if (EDGE) {
            if (v1)
                *po = make_double2(un0, un1);
            else if (v0)
                *reinterpret_cast<double *>(po) = un0;
        } else if (lane) {
            *po = make_double2(un0, un1);
        }
        o += p2;
        double a0, a1;
        atan_over_pi2(un0 * K.inv_eps, un1 * K.inv_eps, s_tab, a0, a1);
        double dq0 = du0, dq1 = du1;
        if (EDGE) {
            a0 = v0 ? a0 : 0.0;
            a1 = v1 ? a1 : 0.0;
            dq0 = v0 ? du0 : 0.0;
            dq1 = v1 ? du1 : 0.0;
        }
        accA += a0;
        accA += a1;
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            accI[c] = fma(I0[c], a0, accI[c]);
            accI[c] = fma(I1[c], a1, accI[c]);
        }
        accS = fma(dq0, dq0, accS);
        accS = fma(dq1, dq1, accS);
        dN0 = upy0;
        dN1 = upy1;
        nyp0 = ny0;
        nyp1 = ny1;
        C = S;
        CW = SW;
        CE = SE;
    };

    int r = 0, t = 0;
    unsigned int sc = 0, sn = TMA_STAGE;  // byte offsets of the stages of tile t and tile t+1
    unsigned int par_next = 0;            // parity of the next wait on tile t+1's stage
#pragma unroll 1
    for (; r + TMA_R <= n; r += TMA_R, ++t) {
        CSV_SEGMENT_BOUNDARY(r)
        // tile t+1 (its first row is the south row of this tile's last row) -- parity of its stage's (t+1)/NST-th completion
        par_next = (unsigned int)(((t + 1) / TMA_NST) & 1);
        mbar_wait(ring_s + 8u * (unsigned int)((t + 1) % TMA_NST), par_next);
        row(mu + sc + TMA_UROW, mi + sc);
        row(mu + sc + 2 * TMA_UROW, mi + sc + TMA_IROW);
        row(mu + sc + 3 * TMA_UROW, mi + sc + 2 * TMA_IROW);
        row(mu + sn, mi + sc + 3 * TMA_IROW);
        __syncwarp();  // every lane has read tile t: its stage may be refilled
        if (lane == 0 && t + TMA_NST < ntiles) issue_tile(t + TMA_NST);
        sc = sn;
        sn = (sn == (TMA_NST - 1) * TMA_STAGE) ? 0u : sn + TMA_STAGE;
    }
    // tail (< TMA_R rows): tile t has been waited for; the south row of local row 3 of a tile lives in tile t+1
    CSV_SEGMENT_BOUNDARY(r)
#pragma unroll 1
    for (int k = 0; r < n; ++r, ++k) {
        if (k == TMA_R - 1) mbar_wait(ring_s + 8u * (unsigned int)((t + 1) % TMA_NST), (unsigned int)(((t + 1) / TMA_NST) & 1));
        const unsigned char *pu = (k == TMA_R - 1) ? mu + sn : mu + sc + (k + 1) * TMA_UROW;
        row(pu, mi + sc + k * TMA_IROW);
    }
    // every issued tile has been waited for unless the tail ended before its last tile was needed: drain
#pragma unroll 1
    for (int q = t + 1; q < ntiles; ++q) mbar_wait(ring_s + 8u * (unsigned int)(q % TMA_NST), (unsigned int)((q / TMA_NST) & 1));
    if (lane) {
        acc[ACC_A] = accA;
        acc[ACC_SQ] = accS;
#pragma unroll
        for (int c = 0; c < NCH; ++c) acc[ACC_IA + c] = accI[c];
    }
}
#endif  // CSV_TMA

template <int NCH, bool STRICT, int MODE>
__global__ void __launch_bounds__(CTA_THREADS, CSV_MIN_CTAS) csv_step_kernel(const __grid_constant__ CsvArgs A) {
    const Geom &G = A.g;
    __shared__ double2 s_tab[ATAN_TAB_N];  // {atan(c_q)/pi, c_q}
#ifdef CSV_TMA
    __shared__ __align__(128) unsigned char s_ring[(MODE == MODE_STEP && !STRICT) ? TMA_BYTES : 16];
#else
    __shared__ __align__(16) unsigned char s_ring[(MODE == MODE_STEP && !STRICT) ? RING_BYTES : 16];
#endif
    const int lane = threadIdx.x;
    constexpr int warp = 0;
    int bid = blockIdx.x;
    const int cb = bid % G.ncb_csv;
    bid /= G.ncb_csv;
    // the production kernel marches through seg_mult consecutive segments per CTA (long row loops amortise the priming;
    // how long is best depends on how many waves the job has, i.e. on the GPU count) and delivers one partial vector per
    // segment; the other instantiations take one segment per CTA
    const int mult = (MODE == MODE_STEP && !STRICT) ? G.seg_mult : 1;
    const int nsup = (G.nseg + mult - 1) / mult;
    // row slabs: the CTAs of the LAST segments go first (then the first ones): they store the slab's bottom rows into the
    // lower neighbour's halo and fence at system scope -- in the first wave that is hidden behind the bulk of the launch,
    // in the last one it would sit in the tail that every rank waits for
    int sup = bid % nsup;
    if (MODE == MODE_STEP && A.cv.p2p && nsup > 1) sup = (sup == 0) ? nsup - 1 : sup - 1;
    const int seg = sup * mult;                       // first local segment
    const int seg_end = min(seg + mult, G.nseg);      // one past the last
    const int img = bid / nsup;
    CsvState *st = A.state + img;
#pragma unroll
    for (int q = 0; q < ATAN_TAB_N; q += 32) s_tab[q + lane] = make_double2(A.atan_tab[q + lane], atan_centre(q + lane));
    if (MODE == MODE_STEP && !STRICT) {
        // launched with programmatic stream serialization: everything above is independent of the previous launch;
        // wait for it (c1/c2, stop flag, level set, halo rows), then let the next launch start filling freed slots
        asm volatile("griddepcontrol.wait;" ::: "memory");
        asm volatile("griddepcontrol.launch_dependents;");
    }
    const int2 ds = *reinterpret_cast<const int2 *>(&st->done);  // {done, steps_done}
    if (MODE == MODE_STEP && ds.x) return;  // frozen image: the launch is a no-op (src/main.cpp:1000)
    __syncwarp();

    const int par = (MODE == MODE_STEP) ? A.par : (ds.y & 1);
    const double *__restrict__ uin = A.u[par] + (size_t)img * G.plane_elems;
    double *__restrict__ uout = (MODE == MODE_KAPPA ? A.kappa_out : A.u[par ^ 1]) + (size_t)img * G.plane_elems;
    const uint8_t *__restrict__ im = A.img + (size_t)img * G.nch * G.plane_elems;

    const int gseg = G.seg0 + seg;
    const int ra = max(gseg * G.seg_rows, G.row_lo);
    const int rb = min((G.seg0 + seg_end) * G.seg_rows, G.row_hi);
    SegCursor cur;
    cur.A = &A;
    cur.img = img;
    cur.seg = seg;
    cur.cb = cb;
    cur.next = (seg_end - seg > 1) ? (gseg + 1) * G.seg_rows - ra : INT_MAX;
    const int cs = cb * CSV_CB + warp * CSV_STRIP_OWN;
    const int a = cs - 2 + 2 * lane;  // this lane's columns: a, a+1
    const int w = G.w, h = G.h;
    const bool colok = a >= 0 && a < G.pitch;
    const bool e2ok = lane == 31 && a + 2 < G.pitch;

    // ---- per-step coefficients
    const double eps = A.eps;
    const double inv_eps = A.inv_eps;
    // fast: du = (kappa*alpha' + sum_k (A_k I^2 + B_k I) + q0) / (eps^2 + u^2), everything pre-scaled by eps/pi
    StepCoef<NCH> K;
    K.q0 = 0.0;
    K.alphap = 0.0;
    K.eps2 = eps * eps;
    K.inv_eps = inv_eps;
    K.linear = true;
    double c1[NCH], c2[NCH];
    if (MODE == MODE_STEP) {
        const double kd = eps * CVB_INV_PI;
        const double bk = A.beta * kd;
        K.alphap = A.alpha * kd;
        K.q0 = A.gamma * kd;
#pragma unroll
        for (int k = 0; k < NCH; ++k) {
            c1[k] = st->c1[k];
            c2[k] = st->c2[k];
            const double l1 = A.lambda1[k], l2 = A.lambda2[k];
            K.cA[k] = bk * (l2 - l1);
            K.linear = K.linear && (l1 == l2);
            K.cB[k] = 2.0 * bk * (l1 * c1[k] - l2 * c2[k]);
            K.q0 += bk * (l2 * c2[k] * c2[k] - l1 * c1[k] * c1[k]);
        }
    }
#ifdef CSV_UNIFORM_COEF
    // every lane computed the same numbers; reading them back from lane 0 tells the compiler they are warp-uniform,
    // so they can live in uniform registers instead of 20 per-thread registers
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
        K.cA[k] = __shfl_sync(0xffffffffu, K.cA[k], 0);
        K.cB[k] = __shfl_sync(0xffffffffu, K.cB[k], 0);
    }
    K.q0 = __shfl_sync(0xffffffffu, K.q0, 0);
    K.alphap = __shfl_sync(0xffffffffu, K.alphap, 0);
    K.eps2 = __shfl_sync(0xffffffffu, K.eps2, 0);
    K.inv_eps = __shfl_sync(0xffffffffu, K.inv_eps, 0);
#endif
    const double q0 = K.q0, alphap = K.alphap, eps2 = K.eps2;
    const double *cA = K.cA, *cB = K.cB;
    const StepCoef<NCH> &K2 = K;

    double acc[NACC];
#pragma unroll
    for (int v = 0; v < NACC; ++v) acc[v] = 0.0;

#ifdef CSV_TMA
#define CSV_ROWS(E, L) csv_rows_tma<NCH, E, L>(uin, uout, &A.tm_u[par], &A.tm_img, img, G, K2, s_tab, s_ring, ra, rb, cs, lane, acc, cur)
#else
#define CSV_ROWS(E, L) csv_rows_ring<NCH, E, L>(uin, uout, im, G, K2, s_tab, s_ring, ra, rb, cs, lane, acc, cur)
#endif
    // CTAs whose stencils stay inside the image take the fast path (all but the outermost ring)
    // (image top and bottom included: the halo rows there hold copies of the border rows, see replicate_border_rows)
    const bool interior = !STRICT && MODE == MODE_STEP && cb > 0 && (cb + 1) * CSV_CB < w;
    const bool edge_fast = !STRICT && MODE == MODE_STEP && !interior && cs < w;
    if (interior) {
        if (K.linear)
            CSV_ROWS(false, true);
        else
            CSV_ROWS(false, false);
    } else if (edge_fast) {
        if (K.linear)
            CSV_ROWS(true, true);
        else
            CSV_ROWS(true, false);
    } else if (cs < w) {
        const int nk = rb - ra + 3;  // streamed rows ra-2 .. rb
        double2 pq[CSV_D];
        double pe[CSV_D];
        unsigned int pI[CSV_D][NCH];
        auto row_off = [&](int k) -> size_t {
            int gr = ra - 2 + k;
            gr = min(max(gr, 0), h - 1);  // BORDER_REPLICATE in i (src/main.cpp:352,354)
            return (size_t)(gr - G.row_lo + HALO) * G.pitch;
        };
        auto issue = [&](int k, int j) {
            pq[j] = make_double2(0.0, 0.0);
            pe[j] = 0.0;
#pragma unroll
            for (int c = 0; c < NCH; ++c) pI[j][c] = 0u;
            if (k < nk) {
                const size_t off = row_off(k);
                if (colok) pq[j] = __ldg(reinterpret_cast<const double2 *>(uin + off + a));
                if (e2ok) pe[j] = __ldg(uin + off + a + 2);
                if (MODE == MODE_STEP && k >= 3 && colok) {
                    const size_t offi = (size_t)(ra - 3 + k - G.row_lo + HALO) * G.pitch + a;
#pragma unroll
                    for (int c = 0; c < NCH; ++c)
                        pI[j][c] = __ldg(reinterpret_cast<const unsigned short *>(im + (size_t)c * G.plane_elems + offi));
                }
            }
        };
#pragma unroll
        for (int j = 0; j < CSV_D; ++j) issue(j, j);

        double2 N = make_double2(0.0, 0.0), C = make_double2(0.0, 0.0);
        double e2c = 0.0, nyp0 = 0.0, nyp1 = 0.0;
#pragma unroll 1
        for (int k0 = 0; k0 < nk; k0 += CSV_D) {
#pragma unroll
            for (int j = 0; j < CSV_D; ++j) {
                const int k = k0 + j;
                if (k < nk) {
                    const double2 S = pq[j];
                    const double e2s = pe[j];
                    unsigned int Ib[NCH];
#pragma unroll
                    for (int c = 0; c < NCH; ++c) Ib[c] = pI[j][c];
                    issue(k + CSV_D, j);
                    if (k + CSV_PF < nk) {
                        const size_t off = row_off(k + CSV_PF);
                        if (colok) prefetch_l2(uin + off + a);
                        if (MODE == MODE_STEP && lane < 3 * NCH) {
                            const int c = lane / 3, o = (lane % 3) * 31;
                            prefetch_l2(im + (size_t)c * G.plane_elems +
                                        (size_t)(ra - 3 + k + CSV_PF - G.row_lo + HALO) * G.pitch + max(cs, 0) + o);
                        }
                    }
                    if (k >= 2) {
                        const int i = ra - 3 + k;  // ny of row i; output row when k >= 3
                        const double ny0 = normal_component<STRICT>(S.x - C.x, S.x - N.x);  // :352,354,367-368
                        const double ny1 = normal_component<STRICT>(S.y - C.y, S.y - N.y);
                        if (k >= 3) {
                            const double Wn = __shfl_up_sync(0xffffffffu, C.y, 1);
                            double E2 = __shfl_down_sync(0xffffffffu, C.x, 1);
                            if (lane == 31) E2 = e2c;
                            // BORDER_REPLICATE in j (:351,353)
                            const double E0 = (a + 1 < w) ? C.y : C.x;
                            const double E1 = (a + 2 < w) ? E2 : C.y;
                            const double W0 = (a >= 1) ? Wn : C.x;
                            const double nx0 = normal_component<STRICT>(E0 - C.x, E0 - W0);  // :351,353,365-366
                            const double nx1 = normal_component<STRICT>(E1 - C.y, E1 - C.x);
                            const double nxw = __shfl_up_sync(0xffffffffu, nx1, 1);
                            // backward differences with replicate on the normalised fields (:371-373)
                            const double kx0 = (a >= 1) ? nx0 - nxw : 0.0;
                            const double kx1 = nx1 - nx0;
                            const double ky0 = (i >= 1) ? ny0 - nyp0 : 0.0;
                            const double ky1 = (i >= 1) ? ny1 - nyp1 : 0.0;
                            const double kap0 = kx0 + ky0, kap1 = kx1 + ky1;
                            const size_t offo = (size_t)(i - G.row_lo + HALO) * G.pitch + a;
                            if (MODE == MODE_KAPPA) {
                                if (lane >= 1) {
                                    if (a + 1 < w)
                                        *reinterpret_cast<double2 *>(uout + offo) = make_double2(kap0, kap1);
                                    else if (a < w)
                                        uout[offo] = kap0;
                                }
                            } else {
                                double I0[NCH], I1[NCH];
#pragma unroll
                                for (int c = 0; c < NCH; ++c) {
                                    I0[c] = u8_to_double(Ib[c] & 0xffu);
                                    I1[c] = u8_to_double(Ib[c] >> 8);
                                }
                                double du0, du1;
                                if (STRICT) {
                                    double s0 = 0.0, s1 = 0.0;  // u_diff, :965, :977-979 serial in k
#pragma unroll
                                    for (int c = 0; c < NCH; ++c) {
                                        double t = __dadd_rn(I0[c], -c1[c]);
                                        const double vi0 = __dmul_rn(__dmul_rn(t, t), A.lambda1[c]);
                                        t = __dadd_rn(I0[c], -c2[c]);
                                        const double vo0 = __dmul_rn(__dmul_rn(t, t), A.lambda2[c]);
                                        s0 = __dadd_rn(s0, __dadd_rn(-vi0, vo0));
                                        t = __dadd_rn(I1[c], -c1[c]);
                                        const double vi1 = __dmul_rn(__dmul_rn(t, t), A.lambda1[c]);
                                        t = __dadd_rn(I1[c], -c2[c]);
                                        const double vo1 = __dmul_rn(__dmul_rn(t, t), A.lambda2[c]);
                                        s1 = __dadd_rn(s1, __dadd_rn(-vi1, vo1));
                                    }
                                    // :985 as one addWeighted, then delta (:204-210), multiply (:992)
                                    const double d0 =
                                        __dadd_rn(__dadd_rn(__dmul_rn(kap0, A.alpha), __dmul_rn(s0, A.beta)), A.gamma);
                                    const double d1 =
                                        __dadd_rn(__dadd_rn(__dmul_rn(kap1, A.alpha), __dmul_rn(s1, A.beta)), A.gamma);
                                    const double e2 = __dmul_rn(eps, eps);
                                    const double de0 =
                                        __ddiv_rn(eps, __dmul_rn(CVB_PI, __dadd_rn(e2, __dmul_rn(C.x, C.x))));
                                    const double de1 =
                                        __ddiv_rn(eps, __dmul_rn(CVB_PI, __dadd_rn(e2, __dmul_rn(C.y, C.y))));
                                    du0 = __dmul_rn(d0, de0);
                                    du1 = __dmul_rn(d1, de1);
                                } else {
                                    double t0 = q0, t1 = q0;
#pragma unroll
                                    for (int c = 0; c < NCH; ++c) {
                                        t0 = fma(fma(cA[c], I0[c], cB[c]), I0[c], t0);
                                        t1 = fma(fma(cA[c], I1[c], cB[c]), I1[c], t1);
                                    }
                                    t0 = fma(kap0, alphap, t0);
                                    t1 = fma(kap1, alphap, t1);
                                    du0 = t0 * fast_rcp(fma(C.x, C.x, eps2));
                                    du1 = t1 * fast_rcp(fma(C.y, C.y, eps2));
                                }
                                const double un0 = __dadd_rn(C.x, du0), un1 = __dadd_rn(C.y, du1);  // :994
                                const bool v0 = lane >= 1 && a < w, v1 = lane >= 1 && a + 1 < w;
                                if (v1)
                                    *reinterpret_cast<double2 *>(uout + offo) = make_double2(un0, un1);
                                else if (v0)
                                    uout[offo] = un0;

                                // sums of the UPDATED level set (next step's c1/c2) and of du^2 (:993)
                                double a0 = atan_over_pi(un0 * inv_eps, s_tab);
                                double a1 = atan_over_pi(un1 * inv_eps, s_tab);
                                a0 = v0 ? a0 : 0.0;
                                a1 = v1 ? a1 : 0.0;
                                du0 = v0 ? du0 : 0.0;
                                du1 = v1 ? du1 : 0.0;
                                acc[ACC_A] += a0;
                                acc[ACC_A] += a1;
#pragma unroll
                                for (int c = 0; c < NCH; ++c) {
                                    acc[ACC_IA + c] = fma(I0[c], a0, acc[ACC_IA + c]);
                                    acc[ACC_IA + c] = fma(I1[c], a1, acc[ACC_IA + c]);
                                }
                                acc[ACC_SQ] = fma(du0, du0, acc[ACC_SQ]);
                                acc[ACC_SQ] = fma(du1, du1, acc[ACC_SQ]);
                            }
                        }
                        nyp0 = ny0;
                        nyp1 = ny1;
                    }
                    N = C;
                    C = S;
                    e2c = e2s;
                }
            }
        }
    }
    if (MODE == MODE_STEP) {
        // image top / bottom: keep copies of the border rows in the halo rows of the buffer just written
        if (cs < w && colok && lane >= 1 && (ra == 0 || rb == h)) replicate_border_rows(uout, G, ra, rb, a);
        // P2P multi-GPU: the slab's first / last HALO rows are the neighbours' halo rows
        bool pushed = false;
        if (A.cv.p2p && cs < w && colok && (ra < G.row_lo + HALO || rb > G.row_hi - HALO))
            pushed = push_boundary_rows(A, uout, par ^ 1, ra, rb, a, lane);
        finish_tile<NCH, false>(A, img, cur.seg, cb, G.ncb_csv, acc, 0, pushed);  // the last (or only) segment of this CTA
    }
}

// Sums of the CURRENT level set and of the image: sum a, sum I_k*a, sum I_k, sum mean_k(I)^2.
// Gives the first step's c1/c2 (src/main.cpp:973-974) and the stop condition (:949-960).
template <int NCH>
__global__ void __launch_bounds__(CTA_THREADS, 16) csv_init_kernel(const __grid_constant__ CsvArgs A, int final_mode) {
    const Geom &G = A.g;
    __shared__ double s_tab[ATAN_TAB_N];
    const int lane = threadIdx.x;
    constexpr int warp = 0;
    int bid = blockIdx.x;
    const int cb = bid % G.ncb_csv;
    bid /= G.ncb_csv;
    const int seg = bid % G.nseg;
    const int img = bid / G.nseg;
    CsvState *st = A.state + img;
#pragma unroll
    for (int q = 0; q < ATAN_TAB_N; q += 32) s_tab[q + lane] = A.atan_tab[q + lane];
    __syncwarp();
    const int par = (final_mode == 1) ? 0 : (st->steps_done & 1);
    const double *__restrict__ uin = A.u[par] + (size_t)img * G.plane_elems;
    const uint8_t *__restrict__ im = A.img + (size_t)img * G.nch * G.plane_elems;
    const int gseg = G.seg0 + seg;
    const int ra = max(gseg * G.seg_rows, G.row_lo);
    const int rb = min((gseg + 1) * G.seg_rows, G.row_hi);
    const int cs = cb * CSV_CB + warp * CSV_STRIP_OWN;
    const int a = cs - 2 + 2 * lane;
    const int w = G.w;
    const double inv_eps = A.inv_eps;
    const double inv_n = 1.0 / (double)NCH;  // Mat /= N multiplies by 1/N (src/main.cpp:958)
    double acc[NACC];
#pragma unroll
    for (int v = 0; v < NACC; ++v) acc[v] = 0.0;
    if (lane >= 1 && a < w) {
        const bool v1 = a + 1 < w;
        constexpr int RB = 4;  // rows per batch: all loads of a batch are issued before its arithmetic (the pass is pure
                               // streaming; one dependent load per row left it latency-bound at a third of the HBM rate)
        for (int i0 = ra; i0 < rb; i0 += RB) {
            double2 Ub[RB];
            unsigned int Bb[RB][NCH];
#pragma unroll
            for (int k = 0; k < RB; ++k) {
                const int i = min(i0 + k, rb - 1);  // rows past the segment re-read its last row and are not summed
                const size_t off = (size_t)(i - G.row_lo + HALO) * G.pitch + a;
                Ub[k] = __ldg(reinterpret_cast<const double2 *>(uin + off));
#pragma unroll
                for (int c = 0; c < NCH; ++c)
                    Bb[k][c] = __ldg(reinterpret_cast<const unsigned short *>(im + (size_t)c * G.plane_elems + off));
            }
#pragma unroll
            for (int k = 0; k < RB; ++k) {
                if (i0 + k >= rb) break;
                const double2 U = Ub[k];
                double a0 = atan_over_pi(U.x * inv_eps, s_tab);
                double a1 = atan_over_pi(U.y * inv_eps, s_tab);
                a1 = v1 ? a1 : 0.0;
                acc[ACC_A] += a0;
                acc[ACC_A] += a1;
                double m0 = 0.0, m1 = 0.0;
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    const unsigned int b = Bb[k][c];
                    const double I0 = u8_to_double(b & 0xffu);
                    const double I1 = v1 ? u8_to_double(b >> 8) : 0.0;
                    acc[ACC_IA + c] = fma(I0, a0, acc[ACC_IA + c]);
                    acc[ACC_IA + c] = fma(I1, a1, acc[ACC_IA + c]);
                    acc[ACC_I + c] += I0;
                    acc[ACC_I + c] += I1;
                    m0 += I0;
                    m1 += I1;
                }
                m0 *= inv_n;
                m1 *= inv_n;
                acc[ACC_SQ] = fma(m0, m0, acc[ACC_SQ]);
                acc[ACC_SQ] = fma(m1, m1, acc[ACC_SQ]);
            }
        }
    }
    finish_tile<NCH, true>(A, img, seg, cb, G.ncb_csv, acc, final_mode);
}

// Multi-rank path: after the all-gather of the group sums, one thread per image.
__global__ void csv_finalize_kernel(const __grid_constant__ CsvArgs A, int mode) {
    if (A.cv.p2p) {  // P2P: one warp waits for every rank's group sums of the newest reduction and folds them
        csv_wait_fold(A);
        return;
    }
    // NCCL path: the host all-gathered the group sums; one warp per image folds them
    const int img = blockIdx.x;
    if (img >= A.g.count) return;
    if (mode == 0 && A.state[img].done) return;
    csv_fold(A, img, mode, 0u);
}

// ---- elementwise helpers ------------------------------------------------------------------------------
// ParallelPixelFunction with f = regularized_delta (src/ParallelPixelFunction.cpp:12-17, main.cpp:204-210,989)
__global__ void delta_map_kernel(double *data, size_t n, double eps) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const double e2 = __dmul_rn(eps, eps);
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += stride) {
        const double x = data[q];
        data[q] = __ddiv_rn(eps, __dmul_rn(CVB_PI, __dadd_rn(e2, __dmul_rn(x, x))));
    }
}

// separate()'s mask: float32(u) > 0, optionally inverted (src/main.cpp:395-400).  Pitched in, pitched out.
__global__ void mask_kernel(const double *u, uint8_t *mask, int rows, int w, int pitch, int invert) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= w) return;
    for (int i = blockIdx.y; i < rows; i += gridDim.y) {  // gridDim.y is capped at 65535: taller planes loop
        const size_t q = (size_t)i * pitch + j;
        const uint8_t m = (__double2float_rn(u[q]) > 0.0f) ? 1 : 0;
        mask[q] = invert ? (uint8_t)(1 - m) : m;
    }
}

// the same mask, 8 pixels per byte (MSB first), rows padded to whole bytes; u is fp64 or fp32
// rule 0: separate()'s float32(u) > 0 (src/main.cpp:395-400); rule 1: the video contour's saturate_cast<uchar>(u) > 0, i.e.
// cvRound(u) >= 1 (src/VideoWriterManager.cpp:65-68)
__global__ void mask_packed_kernel(const void *u, int f32, uint8_t *bits, int rows, int w, int pitch, int invert, int rule) {
    const int jb = blockIdx.x * blockDim.x + threadIdx.x;  // byte within the row
    const int wb = (w + 7) / 8;
    if (jb >= wb) return;
    for (int i = blockIdx.y; i < rows; i += gridDim.y) {
        unsigned int b = 0;
        for (int k = 0; k < 8; ++k) {
            const int j = jb * 8 + k;
            unsigned int m = 0;
            if (j < w) {
                if (rule == 1) {
                    const double x = f32 ? (double)reinterpret_cast<const float *>(u)[(size_t)i * pitch + j]
                                         : reinterpret_cast<const double *>(u)[(size_t)i * pitch + j];
                    m = (__double2int_rn(x) >= 1) ? 1u : 0u;  // round half to even, saturation keeps the sign
                } else {
                    const float v = f32 ? reinterpret_cast<const float *>(u)[(size_t)i * pitch + j]
                                        : __double2float_rn(reinterpret_cast<const double *>(u)[(size_t)i * pitch + j]);
                    m = (v > 0.0f) ? 1u : 0u;
                }
                if (invert) m ^= 1u;
            }
            b |= m << (7 - k);
        }
        bits[(size_t)i * wb + jb] = (uint8_t)b;
    }
}

// every image of a batch at once: image m's current level set is in buffer (steps_done & 1) of its own state
__global__ void mask_packed_batch_kernel(const char *u0, const char *u1, const CsvState *state, int f32, uint8_t *bits, int rows,
                                         int w, int pitch, size_t plane_bytes, int invert) {
    const int jb = blockIdx.x * blockDim.x + threadIdx.x;
    const int wb = (w + 7) / 8;
    if (jb >= wb) return;
    const int img = blockIdx.z;
    const size_t esz = f32 ? 4 : 8;
    const char *u = ((state[img].steps_done & 1) ? u1 : u0) + (size_t)img * plane_bytes + (size_t)HALO * pitch * esz;
    uint8_t *out = bits + (size_t)img * rows * wb;
    for (int i = blockIdx.y; i < rows; i += gridDim.y) {
        unsigned int b = 0;
        for (int k = 0; k < 8; ++k) {
            const int j = jb * 8 + k;
            unsigned int m = 0;
            if (j < w) {
                const float v = f32 ? reinterpret_cast<const float *>(u)[(size_t)i * pitch + j]
                                    : __double2float_rn(reinterpret_cast<const double *>(u)[(size_t)i * pitch + j]);
                m = (v > 0.0f) ? 1u : 0u;
                if (invert) m ^= 1u;
            }
            b |= m << (7 - k);
        }
        out[(size_t)i * wb + jb] = (uint8_t)b;
    }
}

// u0(i,j) = si[i] * sj[j] with host-computed sign vectors (bit parity with glibc sin)
__global__ void checkerboard_kernel(double *u, const signed char *si, const signed char *sj, int row_lo, int rows,
                                    int w, int pitch) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= w) return;
    for (int i = blockIdx.y; i < rows; i += gridDim.y)
        u[(size_t)(i + HALO) * pitch + j] = (double)((int)si[row_lo + i] * (int)sj[j]);
}

// Fill the border halo rows of freshly written planes (upload, initialisers): the HALO rows above row 0 := row 0 when the job
// owns the image top, the HALO rows below row h-1 := row h-1 when it owns the bottom.  One thread per 16 bytes of a row.
__global__ void replicate_halo_kernel(uint8_t *base, size_t plane_bytes, size_t row_bytes, int rows, int top, int bottom) {
    uint8_t *pl = base + (size_t)blockIdx.y * plane_bytes;
    const size_t x = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 16;
    if (x >= row_bytes) return;
    if (top) {
        const uint4 v = *reinterpret_cast<const uint4 *>(pl + (size_t)HALO * row_bytes + x);
        for (int k = 0; k < HALO; ++k) *reinterpret_cast<uint4 *>(pl + (size_t)k * row_bytes + x) = v;
    }
    if (bottom) {
        const uint4 v = *reinterpret_cast<const uint4 *>(pl + (size_t)(HALO + rows - 1) * row_bytes + x);
        for (int k = 0; k < HALO; ++k) *reinterpret_cast<uint4 *>(pl + (size_t)(HALO + rows + k) * row_bytes + x) = v;
    }
}
cudaError_t launch_replicate_halo(void *base, size_t plane_bytes, size_t row_bytes, int nplanes, int rows, int top, int bottom,
                                  cudaStream_t s) {
    if ((!top && !bottom) || nplanes <= 0) return cudaSuccess;
    dim3 grid((unsigned int)((row_bytes / 16 + 127) / 128), nplanes);
    replicate_halo_kernel<<<grid, 128, 0, s>>>(reinterpret_cast<uint8_t *>(base), plane_bytes, row_bytes, rows, top, bottom);
    return cudaGetLastError();
}

// ---- host launchers -----------------------------------------------------------------------------------
template <int NCH>
static cudaError_t launch_step_n(const CsvArgs &A, bool strict, int mode, cudaStream_t s) {
    const Geom &G = A.g;
    const int mult = (mode == MODE_STEP && !strict) ? G.seg_mult : 1;
    const unsigned int grid = (unsigned int)((size_t)G.count * ceil_div(G.nseg, mult) * G.ncb_csv);
    if (mode == MODE_KAPPA) {
        if (strict)
            csv_step_kernel<NCH, true, MODE_KAPPA><<<grid, CTA_THREADS, 0, s>>>(A);
        else
            csv_step_kernel<NCH, false, MODE_KAPPA><<<grid, CTA_THREADS, 0, s>>>(A);
    } else {
        if (strict)
            csv_step_kernel<NCH, true, MODE_STEP><<<grid, CTA_THREADS, 0, s>>>(A);
        else {
            // 16 resident one-warp CTAs x (row ring + atan table) need ~210 KB of shared memory per SM
            const cudaError_t carve = prefer_max_shared(csv_step_kernel<NCH, false, MODE_STEP>);
            if (carve != cudaSuccess) return carve;
            if (!use_pdl(A.multi_rank != 0)) {
                csv_step_kernel<NCH, false, MODE_STEP><<<grid, CTA_THREADS, 0, s>>>(A);
                return cudaGetLastError();
            }
            // Programmatic dependent launch: the CTAs of step n+1 may become resident while the tail of step n (last
            // wave, group reductions) is still running; they load the atan table and then block in
            // griddepcontrol.wait until step n has completed and flushed.
            cudaLaunchConfig_t cfg;
            memset(&cfg, 0, sizeof cfg);
            cfg.gridDim = dim3(grid);
            cfg.blockDim = dim3(CTA_THREADS);
            cfg.stream = s;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            return cudaLaunchKernelEx(&cfg, csv_step_kernel<NCH, false, MODE_STEP>, A);
        }
    }
    return cudaGetLastError();
}

cudaError_t launch_csv_step(const CsvArgs &A, bool strict, cudaStream_t s) {
    return A.g.nch == 1 ? launch_step_n<1>(A, strict, MODE_STEP, s) : launch_step_n<3>(A, strict, MODE_STEP, s);
}
cudaError_t launch_csv_kappa(const CsvArgs &A, bool strict, cudaStream_t s) {
    return launch_step_n<1>(A, strict, MODE_KAPPA, s);
}
cudaError_t launch_csv_init(const CsvArgs &A, int final_mode, cudaStream_t s) {
    const Geom &G = A.g;
    const unsigned int grid = (unsigned int)((size_t)G.count * G.nseg * G.ncb_csv);
    if (G.nch == 1)
        csv_init_kernel<1><<<grid, CTA_THREADS, 0, s>>>(A, final_mode);
    else
        csv_init_kernel<3><<<grid, CTA_THREADS, 0, s>>>(A, final_mode);
    return cudaGetLastError();
}
cudaError_t launch_csv_finalize(const CsvArgs &A, int mode, cudaStream_t s) {
    csv_finalize_kernel<<<A.g.count, 32, 0, s>>>(A, mode);
    return cudaGetLastError();
}
cudaError_t launch_delta_map(double *data, size_t n, double eps, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    const unsigned int grid = (unsigned int)std::min<size_t>((n + 255) / 256, 148 * 16);
    delta_map_kernel<<<grid, 256, 0, s>>>(data, n, eps);
    return cudaGetLastError();
}
cudaError_t launch_mask(const double *u, uint8_t *mask, int rows, int w, int pitch, int invert, cudaStream_t s) {
    dim3 grid((w + 255) / 256, std::min(rows, 65535));
    mask_kernel<<<grid, 256, 0, s>>>(u, mask, rows, w, pitch, invert);
    return cudaGetLastError();
}
cudaError_t launch_mask_packed(const void *u, int f32, uint8_t *bits, int rows, int w, int pitch, int invert, cudaStream_t s,
                               int rule) {
    dim3 grid(((w + 7) / 8 + 127) / 128, std::min(rows, 65535));
    mask_packed_kernel<<<grid, 128, 0, s>>>(u, f32, bits, rows, w, pitch, invert, rule);
    return cudaGetLastError();
}
cudaError_t launch_mask_packed_batch(const void *u0, const void *u1, const CsvState *state, int f32, uint8_t *bits, int count,
                                     int rows, int w, int pitch, size_t plane_bytes, int invert, cudaStream_t s) {
    dim3 grid(((w + 7) / 8 + 127) / 128, std::min(rows, 65535), count);
    mask_packed_batch_kernel<<<grid, 128, 0, s>>>(reinterpret_cast<const char *>(u0), reinterpret_cast<const char *>(u1), state,
                                                  f32, bits, rows, w, pitch, plane_bytes, invert);
    return cudaGetLastError();
}
cudaError_t launch_checkerboard(double *u, const signed char *si, const signed char *sj, int row_lo, int rows, int w,
                                int pitch, cudaStream_t s) {
    dim3 grid((w + 255) / 256, std::min(rows, 65535));
    checkerboard_kernel<<<grid, 256, 0, s>>>(u, si, sj, row_lo, rows, w, pitch);
    return cudaGetLastError();
}

}  // namespace cvb
