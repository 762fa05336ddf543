// Perona-Malik anisotropic diffusion kernels for sm_100a.
//
// pm_step_kernel fuses one iteration of the reference's diffusion loop (src/main.cpp:498-552): the two
// Sobel passes (:503-504), the edge-stopping function g (:513-522) and the 4-neighbour flux update
// (:524-548) in a single pass over HBM -- read I 8 B, write I' 8 B per channel-pixel (the reference makes
// ~9 full-array passes).  The first step reads the uint8 image directly, the last one rounds to uint8
// (saturate_cast, :551), so no separate conversion passes exist.
//
// Mapping: as in csv_kernels.cu a warp (= one CTA) marches down a strip of 64 columns, two per lane.  Rows i..i+2
// of I, the separable Sobel row sums and two rows of g live in registers.  The stencil has radius 2, so lanes 0 and
// 31 are halo lanes: a strip owns 60 columns.  fp64 planes stream through a cp.async shared-memory ring
// (pm_rows_ring: east / west neighbours of I are read from the ring, g and the west flux come from the neighbouring
// lanes by shuffle); the uint8-input first step and the strict path march rows through registers (pm_rows_fast,
// pm_rows_generic).  pm2_step_kernel (pm2_rows_ring) runs TWO diffusion steps per launch -- temporal blocking: the
// intermediate rows never leave the SM -- and is what every pair of fp64 -> fp64 steps goes through.
#include <string.h>

#include "async_copy.cuh"
#include "comm.cuh"
#include "common.cuh"
#include "kernels.h"
#include "math.cuh"

namespace cvb {

#ifndef PM_UNROLL
#define PM_UNROLL 2
#endif
constexpr int kPmUnroll = PM_UNROLL;  // unroll factor of the fast-path row loop

#ifndef PM_D
#define PM_D 2
#endif
#ifndef PM_PF
#define PM_PF 8
#endif
#ifndef PM_MIN_CTAS
#define PM_MIN_CTAS 16
#endif

template <typename T>
__device__ __forceinline__ double2 pm_load2(const T *p);
template <>
__device__ __forceinline__ double2 pm_load2<double>(const double *p) {
    return __ldg(reinterpret_cast<const double2 *>(p));
}
template <>
__device__ __forceinline__ double2 pm_load2<uint8_t>(const uint8_t *p) {  // :495-496 uint8 -> fp64
    const unsigned int b = __ldg(reinterpret_cast<const unsigned short *>(p));
    return make_double2(u8_to_double(b & 0xffu), u8_to_double(b >> 8));
}
// saturate_cast<uchar>(double): cvRound (round-half-even) then clamp, src/main.cpp:551
__device__ __forceinline__ unsigned int sat_u8(double v) {
    const int r = __double2int_rn(v);
    return (unsigned int)min(max(r, 0), 255);
}
__device__ __forceinline__ void pm_store(double *p, double x, double y, bool two) {
    if (two)
        *reinterpret_cast<double2 *>(p) = make_double2(x, y);
    else
        *p = x;
}
__device__ __forceinline__ void pm_store(uint8_t *p, double x, double y, bool two) {
    if (two)
        *reinterpret_cast<unsigned short *>(p) = (unsigned short)(sat_u8(x) | (sat_u8(y) << 8));
    else
        *p = (uint8_t)sat_u8(x);
}

// g = 1 / (1 + (gx^2 + gy^2) / K^2), src/main.cpp:518-521
template <bool STRICT>
__device__ __forceinline__ double edge_stop(double gx, double gy, double K, double inv_k2) {
    if (STRICT) {
        const double mag = __dadd_rn(__dmul_rn(gx, gx), __dmul_rn(gy, gy));
        return __ddiv_rn(1.0, __dadd_rn(1.0, __ddiv_rn(mag, __dmul_rn(K, K))));
    } else {
        const double mag = fma(gx, gx, __dmul_rn(gy, gy));
        return fast_rcp(fma(mag, inv_k2, 1.0));
    }
}
// I0 + L*((gS+g0)(IS-I0) + (gE+g0)(IE-I0) + (gN+g0)(IN-I0) + (gW+g0)(IW-I0))/4, src/main.cpp:544-547
template <bool STRICT>
__device__ __forceinline__ double pm_update(double I0, double IS, double IE, double IN, double IW, double g0,
                                            double gS, double gE, double gN, double gW, double L, double lq) {
    if (STRICT) {
        double s = __dmul_rn(__dadd_rn(gS, g0), __dadd_rn(IS, -I0));
        s = __dadd_rn(s, __dmul_rn(__dadd_rn(gE, g0), __dadd_rn(IE, -I0)));
        s = __dadd_rn(s, __dmul_rn(__dadd_rn(gN, g0), __dadd_rn(IN, -I0)));
        s = __dadd_rn(s, __dmul_rn(__dadd_rn(gW, g0), __dadd_rn(IW, -I0)));
        return __dadd_rn(I0, __ddiv_rn(__dmul_rn(L, s), 4.0));
    } else {
        double s = (gS + g0) * (IS - I0);
        s = fma(gE + g0, IE - I0, s);
        s = fma(gN + g0, IN - I0, s);
        s = fma(gW + g0, IW - I0, s);
        return fma(s, lq, I0);
    }
}


// ---- interior fast path ---------------------------------------------------------------------------------------
// For CTAs whose stencils never touch an image border.  The update is written in flux form: with
//   Fy(i+1/2,j) = (g(i,j) + g(i+1,j)) * (I(i+1,j) - I(i,j)),   Fx(i,j+1/2) = (g(i,j) + g(i,j+1)) * (I(i,j+1) - I(i,j))
// the reference's four terms are  Fy(i+1/2) - Fy(i-1/2) + Fx(j+1/2) - Fx(j-1/2)  (each product bit-identical to the
// reference's; only the order of the final additions differs), so every flux is computed once: the south flux is
// carried to the next row and the west flux comes from the neighbouring lane.  All row queues are two deep, so a
// 2x unrolled loop needs no register moves.  Per iteration: row i+2 arrives, g(i+1) is finished, row i is written.
// EDGE: the strip touches the left or right image border: loads are predicated, g = 1 in the border columns (:516),
// the fluxes across the border are zero (clamped neighbours, :529-530), columns beyond w are not stored.
template <typename TIN, typename TOUT, bool EDGE>
__device__ __forceinline__ void pm_rows_fast(const TIN *__restrict__ in, TOUT *__restrict__ out, const Geom &G, int ra,
                                             int rb, int a, int lane, double inv_k2, double lq) {
    const int w = G.w;
    const bool colok = !EDGE || (a >= 0 && a < G.pitch);
    const bool bc0 = EDGE && (a == 0 || a == w - 1), bc1 = EDGE && a + 1 == w - 1;  // border columns
    const bool nofxw = EDGE && a == 0, nofx0 = EDGE && a == w - 1;
    auto ldr = [&](const TIN *p) { return colok ? pm_load2<TIN>(p) : make_double2(0.0, 0.0); };
    auto fixg = [&](double2 &g) {
        if (EDGE) {
            g.x = bc0 ? 1.0 : g.x;
            g.y = bc1 ? 1.0 : g.y;
        }
    };
    const size_t pitch = (size_t)G.pitch;
    const TIN *pin = in + (size_t)(ra - 2 - G.row_lo + HALO) * pitch + a;  // row ra-2
    TOUT *po = out + (size_t)(ra - G.row_lo + HALO) * pitch + a;           // row ra
    const int n = rb - ra;

    auto sobel_rows = [&](const double2 &X, double2 &rd, double2 &rs) {  // separable Sobel, row pass
        const double Wn = __shfl_up_sync(0xffffffffu, X.y, 1);
        const double E2 = __shfl_down_sync(0xffffffffu, X.x, 1);
        rd.x = X.y - Wn;
        rs.x = fma(2.0, X.x, Wn) + X.y;
        rd.y = E2 - X.x;
        rs.y = fma(2.0, X.y, X.x) + E2;
    };
    auto edge = [&](double gx, double gy) { return fast_rcp(fma(fma(gx, gx, __dmul_rn(gy, gy)), inv_k2, 1.0)); };  // :518-521

    // prologue: rows ra-2 .. ra+1 give g(ra-1), g(ra) and the flux Fy(ra-1/2)
    const double2 X0 = ldr(pin), X1 = ldr(pin + pitch), X2 = ldr(pin + 2 * pitch),
                  X3 = ldr(pin + 3 * pitch);
    double2 rd0, rs0, rd1, rs1, rd2, rs2, rd3, rs3;
    sobel_rows(X0, rd0, rs0);
    sobel_rows(X1, rd1, rs1);
    sobel_rows(X2, rd2, rs2);
    sobel_rows(X3, rd3, rs3);
    double2 gP, gC;  // g(ra-1), g(ra)
    gP.x = edge(fma(2.0, rd1.x, rd0.x) + rd2.x, rs2.x - rs0.x);
    gP.y = edge(fma(2.0, rd1.y, rd0.y) + rd2.y, rs2.y - rs0.y);
    gC.x = edge(fma(2.0, rd2.x, rd1.x) + rd3.x, rs3.x - rs1.x);
    gC.y = edge(fma(2.0, rd2.y, rd1.y) + rd3.y, rs3.y - rs1.y);
    if (ra == 0 || ra == G.h - 1) gC = make_double2(1.0, 1.0);  // g = 1 on the image border rows (:516)
    fixg(gP);
    fixg(gC);
    double fy0 = __dmul_rn(gP.x + gC.x, X2.x - X1.x), fy1 = __dmul_rn(gP.y + gC.y, X2.y - X1.y);  // Fy(ra-1/2); 0 at the image top
    double2 IC = X2, IS = X3;                                                  // rows i, i+1
    double2 P = make_double2(fma(2.0, rd3.x, rd2.x), fma(2.0, rd3.y, rd2.y));  // rd(i) + 2 rd(i+1)
    double2 rdB = rd3;                                                         // rd(i+1)
    double2 rsA = rs2, rsB = rs3;                                              // rs(i), rs(i+1)
    pin += 4 * pitch;                                                          // row ra+2
    double2 q0 = ldr(pin), q1 = make_double2(0.0, 0.0);
    if (n > 1) q1 = ldr(pin + pitch);
    pin += 2 * pitch;  // row ra+4: next row to fetch
#pragma unroll kPmUnroll
    for (int r = 0; r < n; ++r) {
        const double2 X = q0;  // row i+2
        q0 = q1;
        if (r + 2 < n) q1 = ldr(pin);
        if (PM_PF > 0 && r + PM_PF < n) prefetch_l2(pin + (size_t)(PM_PF - 2) * pitch);
        pin += pitch;
        // g(i+1) from the Sobel sums of rows i, i+1, i+2 (:503-504, :513-522)
        double2 rdC, rsC;
        sobel_rows(X, rdC, rsC);
        double2 gS;
        gS.x = edge(P.x + rdC.x, rsC.x - rsA.x);
        gS.y = edge(P.y + rdC.y, rsC.y - rsA.y);
        if (ra + r + 1 == G.h - 1) gS = make_double2(1.0, 1.0);  // g = 1 on the last image row (:516)
        fixg(gS);
        // fluxes of row i and the update (:524-548); at the image top/bottom the halo rows hold copies of the border
        // rows (clamped neighbours, :527-528), so the fluxes across the border are exactly zero
        const double fs0 = __dmul_rn(gC.x + gS.x, IS.x - IC.x), fs1 = __dmul_rn(gC.y + gS.y, IS.y - IC.y);  // Fy(i+1/2)
        const double Ie = __shfl_down_sync(0xffffffffu, IC.x, 1);
        const double ge = __shfl_down_sync(0xffffffffu, gC.x, 1);
        // Fx(a+1/2) = g0 * d0 enters both pixels and nothing else: it is folded into two explicit FMAs
        double g0 = gC.x + gC.y;
        double d0 = IC.y - IC.x;
        double fx1 = __dmul_rn(gC.y + ge, Ie - IC.y);      // Fx(a+3/2)
        if (EDGE) {
            g0 = nofx0 ? 0.0 : g0;  // no flux across the right border; g of the column beyond it may be anything (NaN included)
            d0 = nofx0 ? 0.0 : d0;
            fx1 = bc1 ? 0.0 : fx1;
        }
        double fxw = __shfl_up_sync(0xffffffffu, fx1, 1);  // Fx(a-1/2)
        if (EDGE) fxw = nofxw ? 0.0 : fxw;
        const double o0 = fma((fs0 - fy0) + fma(g0, d0, -fxw), lq, IC.x);
        const double o1 = fma((fs1 - fy1) + fma(-g0, d0, fx1), lq, IC.y);
        if (EDGE) {
            if (lane >= 1 && lane <= 30 && a < w) pm_store(po, o0, o1, a + 1 < w);
        } else if (lane >= 1 && lane <= 30) {
            pm_store(po, o0, o1, true);
        }
        po += pitch;
        // next row
        fy0 = fs0;
        fy1 = fs1;
        P.x = fma(2.0, rdC.x, rdB.x);
        P.y = fma(2.0, rdC.y, rdB.y);
        rdB = rdC;
        rsA = rsB;
        rsB = rsC;
        IC = IS;
        IS = X;
        gC = gS;
    }
}

// ---- fast path with an asynchronous shared-memory row ring (fp64 input planes) ---------------------------
// As csv_rows_ring: rows travel HBM -> shared memory with cp.async, ~8 rows ahead of their use, two rows per commit
// group; the east / west neighbours of a row are read from the ring (6 of the 10 shuffles per row go away) and no
// registers are held by loads in flight.  Ring row k holds image row ra-2+k in slot k % PM_RING_NS:
//   [16 + 16*j, +16)  chunk j = 0..31: columns cs-2+2j, cs-1+2j (lane j's two columns)
// Everything that is carried from row to row has period 2 or 3, and the loop is unrolled 6x: no register moves; with
// PM_RING_NS = 12 the slot offsets are compile-time constants on top of a base that alternates between 0 and 6 slots.
// The arithmetic is that of pm_rows_fast, operation for operation.
constexpr int PM_RING_NS = 12;
constexpr int PM_RING_SLOT = 16 + 32 * 16 + 16;
constexpr int PM_RING_BYTES = PM_RING_NS * PM_RING_SLOT;
static_assert(PM_RING_NS + 2 <= TAIL_ROWS + HALO, "tail padding too small for the PM ring");

// LASTSEG: the segment ends at the image bottom (the only place where g = 1 must be forced inside the loop, :516).
template <typename TOUT, bool EDGE, bool LASTSEG>
__device__ __forceinline__ void pm_rows_ring(const double *__restrict__ in, TOUT *__restrict__ out, const Geom &G,
                                             unsigned char *ring, int ra, int rb, int a, int lane, double inv_k2, double lq) {
    const int w = G.w;
    const bool colok = !EDGE || (a >= 0 && a < G.pitch);
    const bool bc0 = EDGE && (a == 0 || a == w - 1), bc1 = EDGE && a + 1 == w - 1;  // border columns
    const bool nofxw = EDGE && a == 0, nofx0 = EDGE && a == w - 1;
    auto fixg = [&](double2 &g) {
        if (EDGE) {
            g.x = bc0 ? 1.0 : g.x;
            g.y = bc1 ? 1.0 : g.y;
        }
    };
    const size_t pitch = (size_t)G.pitch;
    const double *pin = in + (size_t)(ra - 2 - G.row_lo + HALO) * pitch + a;  // row ra-2 = ring row 0
    TOUT *po = out + (size_t)(ra - G.row_lo + HALO) * pitch + a;              // row ra
    const int n = rb - ra;
    const unsigned int ring_s = (unsigned int)__cvta_generic_to_shared(ring) + 16 + 16 * lane;
    const unsigned char *my = ring + 16 + 16 * lane;  // + slot: own chunk; west neighbour at -8, east at +16

    auto issue = [&](unsigned int slot_off) {  // the next ring row
        if (colok) cp_async16(ring_s + slot_off, pin);
        pin += pitch;
    };
    struct Row {
        double2 X;
        double Wn, E2;
    };
    auto fetch = [&](unsigned int slot_off) {
        Row r;
        r.X = *reinterpret_cast<const double2 *>(my + slot_off);
        r.Wn = *reinterpret_cast<const double *>(my + slot_off - 8);
        r.E2 = *reinterpret_cast<const double *>(my + slot_off + 16);
        return r;
    };
    auto sobel_rows = [&](const Row &r, double2 &rd, double2 &rs) {  // separable Sobel, row pass
        rd.x = r.X.y - r.Wn;
        rs.x = fma(2.0, r.X.x, r.Wn) + r.X.y;
        rd.y = r.E2 - r.X.x;
        rs.y = fma(2.0, r.X.y, r.X.x) + r.E2;
    };
    auto edge = [&](double gx, double gy) { return fast_rcp(fma(fma(gx, gx, __dmul_rn(gy, gy)), inv_k2, 1.0)); };  // :518-521

    // ring rows 0 .. NS-1, two per group
#pragma unroll
    for (int k = 0; k < PM_RING_NS; k += 2) {
        issue(k * PM_RING_SLOT);
        issue((k + 1) * PM_RING_SLOT);
        cp_async_commit();
    }
    cp_async_wait<PM_RING_NS / 2 - 2>();  // ring rows 0..3 have landed
    __syncwarp();
    // prologue: rows ra-2 .. ra+1 give g(ra-1), g(ra) and the flux Fy(ra-1/2)
    const Row Q0 = fetch(0), Q1 = fetch(PM_RING_SLOT), Q2 = fetch(2 * PM_RING_SLOT), Q3 = fetch(3 * PM_RING_SLOT);
    double2 rd0, rs0, rd1, rs1, rd2, rs2, rd3, rs3;
    sobel_rows(Q0, rd0, rs0);
    sobel_rows(Q1, rd1, rs1);
    sobel_rows(Q2, rd2, rs2);
    sobel_rows(Q3, rd3, rs3);
    double2 gP, gC;  // g(ra-1), g(ra)
    gP.x = edge(fma(2.0, rd1.x, rd0.x) + rd2.x, rs2.x - rs0.x);
    gP.y = edge(fma(2.0, rd1.y, rd0.y) + rd2.y, rs2.y - rs0.y);
    gC.x = edge(fma(2.0, rd2.x, rd1.x) + rd3.x, rs3.x - rs1.x);
    gC.y = edge(fma(2.0, rd2.y, rd1.y) + rd3.y, rs3.y - rs1.y);
    if (ra == 0 || ra == G.h - 1) gC = make_double2(1.0, 1.0);  // g = 1 on the image border rows (:516)
    fixg(gP);
    fixg(gC);
    double fy0 = __dmul_rn(gP.x + gC.x, Q2.X.x - Q1.X.x), fy1 = __dmul_rn(gP.y + gC.y, Q2.X.y - Q1.X.y);  // Fy(ra-1/2)
    double2 IC = Q2.X, IS = Q3.X;                                              // rows i, i+1
    double ICe = Q2.E2, ISe = Q3.E2;                                           // their east neighbours (column a+2)
    double2 P = make_double2(fma(2.0, rd3.x, rd2.x), fma(2.0, rd3.y, rd2.y));  // rd(i) + 2 rd(i+1)
    double2 rdB = rd3;                                                         // rd(i+1)
    double2 rsA = rs2, rsB = rs3;                                              // rs(i), rs(i+1)
    int r = 0;

    // one output row i = ra + r: ring row r+4 (row i+2) arrives from slot s_x
    auto row = [&](unsigned int s_x) {
        const Row Q = fetch(s_x);
        // g(i+1) from the Sobel sums of rows i, i+1, i+2 (:503-504, :513-522)
        double2 rdC, rsC;
        sobel_rows(Q, rdC, rsC);
        double2 gS;
        gS.x = edge(P.x + rdC.x, rsC.x - rsA.x);
        gS.y = edge(P.y + rdC.y, rsC.y - rsA.y);
        if (LASTSEG && ra + r + 1 == G.h - 1) gS = make_double2(1.0, 1.0);  // g = 1 on the last image row (:516)
        fixg(gS);
        // fluxes of row i and the update (:524-548); at the image top/bottom the halo rows hold copies of the border
        // rows (clamped neighbours, :527-528), so the fluxes across the border are exactly zero
        const double fs0 = __dmul_rn(gC.x + gS.x, IS.x - IC.x), fs1 = __dmul_rn(gC.y + gS.y, IS.y - IC.y);  // Fy(i+1/2)
        const double ge = __shfl_down_sync(0xffffffffu, gC.x, 1);
        double g0 = gC.x + gC.y;                           // Fx(a+1/2) = g0 * d0, folded into the two FMAs below
        double d0 = IC.y - IC.x;
        double fx1 = __dmul_rn(gC.y + ge, ICe - IC.y);     // Fx(a+3/2)
        if (EDGE) {
            g0 = nofx0 ? 0.0 : g0;  // no flux across the right border; g of the column beyond it may be anything (NaN included)
            d0 = nofx0 ? 0.0 : d0;
            fx1 = bc1 ? 0.0 : fx1;
        }
        double fxw = __shfl_up_sync(0xffffffffu, fx1, 1);  // Fx(a-1/2)
        if (EDGE) fxw = nofxw ? 0.0 : fxw;
        const double o0 = fma((fs0 - fy0) + fma(g0, d0, -fxw), lq, IC.x);
        const double o1 = fma((fs1 - fy1) + fma(-g0, d0, fx1), lq, IC.y);
        if (EDGE) {
            if (lane >= 1 && lane <= 30 && a < w) pm_store(po, o0, o1, a + 1 < w);
        } else if (lane >= 1 && lane <= 30) {
            pm_store(po, o0, o1, true);
        }
        po += pitch;
        ++r;
        // next row
        fy0 = fs0;
        fy1 = fs1;
        P.x = fma(2.0, rdC.x, rdB.x);
        P.y = fma(2.0, rdC.y, rdB.y);
        rdB = rdC;
        rsA = rsB;
        rsB = rsC;
        IC = IS;
        ICe = ISe;
        IS = Q.X;
        ISe = Q.E2;
        gC = gS;
    };
    // two rows: refill the two slots freed longest ago (ring rows r, r+1 -> r+NS, r+NS+1), wait for ring rows r+4, r+5
    auto pair = [&](unsigned int s_w0, unsigned int s_x0, unsigned int s_x1) {
        issue(s_w0);
        issue(s_w0 + PM_RING_SLOT);
        cp_async_commit();
        cp_async_wait<PM_RING_NS / 2 - 2>();
        __syncwarp();
        row(s_x0);
        row(s_x1);
    };
    constexpr unsigned int S = PM_RING_SLOT, H = 6 * PM_RING_SLOT;
    unsigned int tog = 0;  // offset of the slot of ring row r: 0 or 6 slots
#pragma unroll 1
    while (r + 6 <= n) {
        const unsigned int t2 = H - tog;
        pair(tog, tog + 4 * S, tog + 5 * S);
        pair(tog + 2 * S, t2, t2 + S);
        pair(tog + 4 * S, t2 + 2 * S, t2 + 3 * S);
        tog = t2;
    }
    cp_async_wait<0>();  // the rows of the tail (<= 5) have all been requested
    __syncwarp();
#pragma unroll 1
    while (r < n) row((unsigned int)((r + 4) % PM_RING_NS) * S);
}

// ---- TWO diffusion steps per launch (temporal blocking) --------------------------------------------------------
// The PM step is HBM-bound (16 B per channel-pixel against ~43 FP64 operations); two consecutive steps fused into one
// pass read and write every plane once instead of twice.  A warp marches two stencils down its strip, the second one
// two rows behind the first:
//   stage A = pm_rows_ring, operation for operation: I rows from the cp.async input ring -> J = PM(I), rows ra-2 .. rb+1;
//             instead of going to HBM a J row is put into a small shared-memory ring (its east / west neighbours come
//             back from there, as those of I come from the input ring);
//   stage B = the same arithmetic on the J rows -> out = PM(J), rows ra .. rb-1.
// The stencil of two steps has radius 4: lanes 0, 1, 30 and 31 are halo lanes (a strip owns 56 columns), stage A needs
// I rows ra-4 .. rb+3 (HALO = 4 rows around every slab).  Bit-identical to two single launches: every J value is the
// double the first launch would have stored, every operation of stage B the one the second launch would execute.
// (For that to hold whatever the compiler does, no multiplication that feeds an addition is left to FMA contraction in
// ANY of the fast row functions of this file: flux products are __dmul_rn, everything else is an explicit fma or an add.)
// Image borders: g = 1 on the border rows and columns and zero flux across them in BOTH stages (:516, :527-530).  The
// J rows above row 0 / below row h-1 that stage A makes out of the replicated halo rows are NOT copies of J(0) / J(h-1),
// as the clamped neighbours of the second step require; stage B therefore takes the flux across the top border as
// zero and substitutes J(h-1) for the rows below the image (BROWS instantiation only).
constexpr int PM2_STRIP_OWN = 56;
constexpr int PM2_JNS = 6;  // J ring slots: the prologue of stage B needs 4 rows; 6 keeps every slot offset constant
constexpr int PM2_SMEM = (PM_RING_NS + PM2_JNS) * PM_RING_SLOT;
#ifndef PM2_MIN_CTAS
#define PM2_MIN_CTAS 16
#endif

template <bool EDGE, bool BROWS>
__device__ __forceinline__ void pm2_rows_ring(const double *__restrict__ in, double *__restrict__ out, const Geom &G,
                                              unsigned char *ring, unsigned char *jring, int ra, int rb, int a, int lane,
                                              double inv_k2, double lq) {
    const int w = G.w, h = G.h;
    const bool colok = !EDGE || (a >= 0 && a < G.pitch);
    const bool bc0 = EDGE && (a == 0 || a == w - 1), bc1 = EDGE && a + 1 == w - 1;  // border columns
    const bool nofxw = EDGE && a == 0, nofx0 = EDGE && a == w - 1;
    const bool own = lane >= 2 && lane <= 29;
    auto fixg = [&](double2 &g) {
        if (EDGE) {
            g.x = bc0 ? 1.0 : g.x;
            g.y = bc1 ? 1.0 : g.y;
        }
    };
    const size_t pitch = (size_t)G.pitch;
    const double *pin = in + (size_t)(ra - 4 - G.row_lo + HALO) * pitch + a;  // row ra-4 = ring row 0
    double *po = out + (size_t)(ra - G.row_lo + HALO) * pitch + a;            // row ra
    const int n = rb - ra;
    const unsigned int ring_s = (unsigned int)__cvta_generic_to_shared(ring) + 16 + 16 * lane;
    const unsigned char *my = ring + 16 + 16 * lane;  // + slot: own chunk; west neighbour at -8, east at +16
    unsigned char *jmy = jring + 16 + 16 * lane;

    auto issue = [&](unsigned int slot_off) {  // the next ring row
        if (colok) cp_async16(ring_s + slot_off, pin);
        pin += pitch;
    };
    struct Row {
        double2 X;
        double Wn, E2;
    };
    auto fetch = [&](const unsigned char *base, unsigned int slot_off) {
        Row r;
        r.X = *reinterpret_cast<const double2 *>(base + slot_off);
        r.Wn = *reinterpret_cast<const double *>(base + slot_off - 8);
        r.E2 = *reinterpret_cast<const double *>(base + slot_off + 16);
        return r;
    };
    auto sobel_rows = [&](const Row &r, double2 &rd, double2 &rs) {  // separable Sobel, row pass
        rd.x = r.X.y - r.Wn;
        rs.x = fma(2.0, r.X.x, r.Wn) + r.X.y;
        rd.y = r.E2 - r.X.x;
        rs.y = fma(2.0, r.X.y, r.X.x) + r.E2;
    };
    auto edge = [&](double gx, double gy) { return fast_rcp(fma(fma(gx, gx, __dmul_rn(gy, gy)), inv_k2, 1.0)); };  // :518-521

    // what one diffusion step carries from row to row (the locals of pm_rows_ring)
    struct Carry {
        double2 IC, IS;    // rows i, i+1
        double ICe, ISe;   // their east neighbours (column a+2)
        double2 P;         // rd(i) + 2 rd(i+1)
        double2 rdB;       // rd(i+1)
        double2 rsA, rsB;  // rs(i), rs(i+1)
        double2 gC;        // g(i)
        double fy0, fy1;   // Fy(i-1/2)
    };
    // rows i-2 .. i+1 give g(i-1), g(i) and the flux Fy(i-1/2); i = first output row of the stage
    auto prologue = [&](Carry &c, const Row &Q0, const Row &Q1, const Row &Q2, const Row &Q3, int i, bool zero_top_flux) {
        double2 rd0, rs0, rd1, rs1, rd2, rs2, rd3, rs3;
        sobel_rows(Q0, rd0, rs0);
        sobel_rows(Q1, rd1, rs1);
        sobel_rows(Q2, rd2, rs2);
        sobel_rows(Q3, rd3, rs3);
        double2 gP;
        gP.x = edge(fma(2.0, rd1.x, rd0.x) + rd2.x, rs2.x - rs0.x);
        gP.y = edge(fma(2.0, rd1.y, rd0.y) + rd2.y, rs2.y - rs0.y);
        c.gC.x = edge(fma(2.0, rd2.x, rd1.x) + rd3.x, rs3.x - rs1.x);
        c.gC.y = edge(fma(2.0, rd2.y, rd1.y) + rd3.y, rs3.y - rs1.y);
        if (i == 0 || i == h - 1) c.gC = make_double2(1.0, 1.0);  // g = 1 on the image border rows (:516)
        fixg(gP);
        fixg(c.gC);
        c.fy0 = __dmul_rn(gP.x + c.gC.x, Q2.X.x - Q1.X.x);
        c.fy1 = __dmul_rn(gP.y + c.gC.y, Q2.X.y - Q1.X.y);
        if (BROWS && zero_top_flux && i == 0) c.fy0 = c.fy1 = 0.0;  // stage B: J(-1) := J(0)
        c.IC = Q2.X;
        c.IS = Q3.X;
        c.ICe = Q2.E2;
        c.ISe = Q3.E2;
        c.P = make_double2(fma(2.0, rd3.x, rd2.x), fma(2.0, rd3.y, rd2.y));
        c.rdB = rd3;
        c.rsA = rs2;
        c.rsB = rs3;
    };
    // one output row i of a stage: Q = row i+2
    auto step = [&](Carry &c, const Row &Q, int i, double &o0, double &o1) {
        double2 rdC, rsC;
        sobel_rows(Q, rdC, rsC);
        double2 gS;
        gS.x = edge(c.P.x + rdC.x, rsC.x - c.rsA.x);
        gS.y = edge(c.P.y + rdC.y, rsC.y - c.rsA.y);
        if (BROWS && (i + 1 == 0 || i + 1 == h - 1)) gS = make_double2(1.0, 1.0);  // g = 1 on the border rows (:516)
        fixg(gS);
        const double fs0 = __dmul_rn(c.gC.x + gS.x, c.IS.x - c.IC.x), fs1 = __dmul_rn(c.gC.y + gS.y, c.IS.y - c.IC.y);  // Fy(i+1/2)
        const double ge = __shfl_down_sync(0xffffffffu, c.gC.x, 1);
        double g0 = c.gC.x + c.gC.y;                          // Fx(a+1/2) = g0 * d0, folded into the two FMAs below
        double d0 = c.IC.y - c.IC.x;
        double fx1 = __dmul_rn(c.gC.y + ge, c.ICe - c.IC.y);  // Fx(a+3/2)
        if (EDGE) {
            g0 = nofx0 ? 0.0 : g0;  // no flux across the right border; g of the column beyond it may be anything (NaN included)
            d0 = nofx0 ? 0.0 : d0;
            fx1 = bc1 ? 0.0 : fx1;
        }
        double fxw = __shfl_up_sync(0xffffffffu, fx1, 1);  // Fx(a-1/2)
        if (EDGE) fxw = nofxw ? 0.0 : fxw;
        o0 = fma((fs0 - c.fy0) + fma(g0, d0, -fxw), lq, c.IC.x);
        o1 = fma((fs1 - c.fy1) + fma(-g0, d0, fx1), lq, c.IC.y);
        c.fy0 = fs0;
        c.fy1 = fs1;
        c.P.x = fma(2.0, rdC.x, c.rdB.x);
        c.P.y = fma(2.0, rdC.y, c.rdB.y);
        c.rdB = rdC;
        c.rsA = c.rsB;
        c.rsB = rsC;
        c.IC = c.IS;
        c.ICe = c.ISe;
        c.IS = Q.X;
        c.ISe = Q.E2;
        c.gC = gS;
    };

    constexpr unsigned int S = PM_RING_SLOT, H = 6 * PM_RING_SLOT;
    // ring rows 0 .. NS-1, two per group
#pragma unroll
    for (int k = 0; k < PM_RING_NS; k += 2) {
        issue(k * S);
        issue((k + 1) * S);
        cp_async_commit();
    }
    cp_async_wait<PM_RING_NS / 2 - 2>();  // ring rows 0..3 have landed
    __syncwarp();
    Carry A, B;
    prologue(A, fetch(my, 0), fetch(my, S), fetch(my, 2 * S), fetch(my, 3 * S), ra - 2, false);
    int rowA = ra - 2;  // J row stage A produces next; it consumes ring row (rowA - ra) + 6
    // stage A alone: J rows ra-2 .. ra+1 into J slots 0..3 (ring rows 4..7; refill ring slots 0..3 with rows 12..15)
    auto a_row = [&](unsigned int s_x, unsigned int s_j) {
        double o0, o1;
        step(A, fetch(my, s_x), rowA, o0, o1);
        *reinterpret_cast<double2 *>(jmy + s_j) = make_double2(o0, o1);
        ++rowA;
    };
#pragma unroll
    for (int k = 0; k < 4; k += 2) {
        issue(k * S);
        issue((k + 1) * S);
        cp_async_commit();
        cp_async_wait<PM_RING_NS / 2 - 2>();
        __syncwarp();
        a_row((4 + k) * S, k * S);
        a_row((5 + k) * S, (k + 1) * S);
    }
    __syncwarp();
    {
        const Row J2 = fetch(jmy, 2 * S);
        Row J3 = fetch(jmy, 3 * S);
        if (BROWS && ra + 1 >= h) J3.X = J2.X;  // one-row image: J(1) := J(0)
        prologue(B, fetch(jmy, 0), fetch(jmy, S), J2, J3, ra, true);
    }
    int r = 0;
    // one output row ra + r: stage A makes J row ra+2+r from ring row r+8 (slot s_x) into J slot s_j, stage B consumes it
    auto row = [&](unsigned int s_x, unsigned int s_j) {
        a_row(s_x, s_j);
        __syncwarp();
        Row Q = fetch(jmy, s_j);
        if (BROWS && ra + r + 2 >= h) Q.X = B.IS;  // below the image: J(h), J(h+1) := J(h-1) (clamped neighbours, :527-528)
        double o0, o1;
        step(B, Q, ra + r, o0, o1);
        if (EDGE) {
            if (own && a < w) pm_store(po, o0, o1, a + 1 < w);
        } else if (own) {
            pm_store(po, o0, o1, true);
        }
        po += pitch;
        ++r;
    };
    // two rows: refill the two ring slots freed longest ago (ring rows r+4, r+5 -> r+16, r+17), wait for ring rows r+8, r+9
    auto pair = [&](unsigned int s_w0, unsigned int s_x0, unsigned int s_j0) {
        issue(s_w0);
        issue(s_w0 + S);
        cp_async_commit();
        cp_async_wait<PM_RING_NS / 2 - 2>();
        __syncwarp();
        row(s_x0, s_j0);
        row(s_x0 + S, s_j0 + S);
    };
    unsigned int tog = 0;  // (r mod 12) slots: 0 or 6
#pragma unroll 1
    while (r + 6 <= n) {
        const unsigned int t2 = H - tog;
        pair(tog + 4 * S, t2 + 2 * S, 4 * S);
        pair(t2, t2 + 4 * S, 0);
        pair(t2 + 2 * S, tog, 2 * S);
        tog = t2;
    }
    cp_async_wait<0>();  // the rows of the tail (<= 5) have all been requested
    __syncwarp();
#pragma unroll 1
    while (r < n) row((unsigned int)((r + 8) % PM_RING_NS) * S, (unsigned int)((r + 4) % PM2_JNS) * S);
}

template <typename TIN, typename TOUT, bool STRICT>
__device__ __forceinline__ void pm_rows_generic(const TIN *__restrict__ in, TOUT *__restrict__ out, const Geom &G, int ra,
                                                int rb, int a, int lane, bool colok, double K, double L, double inv_k2,
                                                double lq) {
    const int w = G.w, h = G.h;

    const int nk = rb - ra + 4;  // streamed rows ra-2 .. rb+1
    auto row_off = [&](int k) -> size_t {
        int gr = ra - 2 + k;
        gr = min(max(gr, 0), h - 1);  // clamped neighbours in i (:527-528)
        return (size_t)(gr - G.row_lo + HALO) * G.pitch;
    };
    double2 pq[PM_D];
    auto issue = [&](int k, int j) {
        pq[j] = make_double2(0.0, 0.0);
        if (k < nk && colok) pq[j] = pm_load2<TIN>(in + row_off(k) + a);
    };
#pragma unroll
    for (int j = 0; j < PM_D; ++j) issue(j, j);

    const double2 z2 = make_double2(0.0, 0.0);
    double2 IN = z2, IC = z2, IS = z2;     // rows r-3, r-2, r-1 of I
    double2 rdA = z2, rdB = z2, rsA = z2, rsB = z2;  // Sobel row sums of rows r-2, r-1
    double2 gN = z2, gC = z2;              // g of rows r-3, r-2
    const bool bc0 = (a == 0) || (a == w - 1), bc1 = (a + 1 == w - 1);  // border columns (:516)
#pragma unroll 1
    for (int k0 = 0; k0 < nk; k0 += PM_D) {
#pragma unroll
        for (int j = 0; j < PM_D; ++j) {
            const int k = k0 + j;
            if (k < nk) {
                const double2 X = pq[j];  // row r = ra-2+k
                issue(k + PM_D, j);
                if (k + PM_PF < nk && colok) prefetch_l2(in + row_off(k + PM_PF) + a);
                // separable 3x3 Sobel, row pass (summation order of cv2: (l + 2c) + r, r - l)
                const double Wn = __shfl_up_sync(0xffffffffu, X.y, 1);
                const double E2 = __shfl_down_sync(0xffffffffu, X.x, 1);
                double2 rdC, rsC;
                rdC.x = X.y - Wn;
                rsC.x = __dadd_rn(fma(2.0, X.x, Wn), X.y);
                rdC.y = E2 - X.x;
                rsC.y = __dadd_rn(fma(2.0, X.y, X.x), E2);
                double2 gS = z2;
                if (k >= 2) {
                    // g of row r-1: column pass 2*m + (t + b), b - t, then :513-522
                    const int ig = ra - 3 + k;
                    const double gx0 = fma(2.0, rdB.x, __dadd_rn(rdA.x, rdC.x));
                    const double gx1 = fma(2.0, rdB.y, __dadd_rn(rdA.y, rdC.y));
                    const double gy0 = rsC.x - rsA.x, gy1 = rsC.y - rsA.y;
                    gS.x = edge_stop<STRICT>(gx0, gy0, K, inv_k2);
                    gS.y = edge_stop<STRICT>(gx1, gy1, K, inv_k2);
                    const bool br = (ig == 0) || (ig == h - 1);
                    if (br || bc0) gS.x = 1.0;
                    if (br || bc1) gS.y = 1.0;
                }
                if (k >= 4) {
                    const int i = ra - 4 + k;  // output row: I rows IN, IC, IS = i-1, i, i+1
                    const double Wc = __shfl_up_sync(0xffffffffu, IC.y, 1);
                    const double Ec = __shfl_down_sync(0xffffffffu, IC.x, 1);
                    const double gW = __shfl_up_sync(0xffffffffu, gC.y, 1);
                    const double gE = __shfl_down_sync(0xffffffffu, gC.x, 1);
                    // clamped neighbours in j (:529-530): a clamped neighbour is the pixel itself
                    const bool w0 = a >= 1, e0 = a + 1 < w, e1 = a + 2 < w;
                    const double o0 = pm_update<STRICT>(IC.x, IS.x, e0 ? IC.y : IC.x, IN.x, w0 ? Wc : IC.x, gC.x, gS.x,
                                                        e0 ? gC.y : gC.x, gN.x, w0 ? gW : gC.x, L, lq);
                    const double o1 = pm_update<STRICT>(IC.y, IS.y, e1 ? Ec : IC.y, IN.y, IC.x, gC.y, gS.y,
                                                        e1 ? gE : gC.y, gN.y, gC.x, L, lq);
                    if (lane >= 1 && lane <= 30 && a < w)
                        pm_store(out + (size_t)(i - G.row_lo + HALO) * G.pitch + a, o0, o1, a + 1 < w);
                }
                IN = IC;
                IC = IS;
                IS = X;
                rdA = rdB;
                rdB = rdC;
                rsA = rsB;
                rsB = rsC;
                gN = gC;
                gC = gS;
            }
        }
    }
}

// Clamped neighbours in i for the fast path: the halo rows above row 0 / below row h-1 of the plane just written get
// copies of those rows (each lane copies the columns it owns, read back from L2).
template <typename TOUT>
__device__ __noinline__ void pm_replicate_border(TOUT *out, const Geom &G, int ra, int rb, int a) {
    const bool two = a + 1 < G.w;
    if (ra == 0) {
        const TOUT *src = out + (size_t)(0 - G.row_lo + HALO) * G.pitch + a;
        const TOUT v0 = __ldcg(src), v1 = two ? __ldcg(src + 1) : v0;
        for (int k = 0; k < HALO; ++k) {
            out[(size_t)k * G.pitch + a] = v0;
            if (two) out[(size_t)k * G.pitch + a + 1] = v1;
        }
    }
    if (rb == G.h) {
        const size_t last = (size_t)(G.h - 1 - G.row_lo + HALO);
        const TOUT *src = out + last * G.pitch + a;
        const TOUT v0 = __ldcg(src), v1 = two ? __ldcg(src + 1) : v0;
        for (int k = 1; k <= HALO; ++k) {
            out[(last + k) * G.pitch + a] = v0;
            if (two) out[(last + k) * G.pitch + a + 1] = v1;
        }
    }
}

// P2P multi-GPU: the slab's first / last HALO rows are the neighbours' halo rows.  Read back what this warp just
// wrote and store it into the neighbour's output buffer; the LAST boundary CTA of the launch then raises the
// neighbour's flag (st.release.sys after system fences), which pm_wait_kernel polls before the next launch.
// lane_lo .. lane_hi: the lanes that own columns; ncb / seg_rows / nseg: the tiling of the launching kernel.
template <typename TOUT>
__device__ __noinline__ void pm_push_boundary(const PmArgs &A, const TOUT *out, int plane, int ra, int rb, int a, int lane,
                                              int lane_lo, int lane_hi, int ncb, int seg_rows, int nseg) {
    const Geom &G = A.g;
    const bool top = ra < G.row_lo + HALO && A.cv.rank > 0, bot = rb > G.row_hi - HALO && A.cv.rank < A.cv.nranks - 1;
    if (!top && !bot) return;
    if (sizeof(TOUT) == 8 && A.out_buf >= 0 && lane >= lane_lo && lane <= lane_hi && a < G.w) {
        const double *o = reinterpret_cast<const double *>(out);
        for (int i = ra; i < rb; ++i) {
            const bool t = top && i < G.row_lo + HALO, b = bot && i >= G.row_hi - HALO;
            if (!t && !b) continue;
            const double *src = o + (size_t)(i - G.row_lo + HALO) * G.pitch + a;
            const bool two = a + 1 < G.w;
            const double v0 = __ldcg(src), v1 = two ? __ldcg(src + 1) : 0.0;
            if (t) {
                double *d = A.cv.up_pm[A.out_buf] + (size_t)plane * (size_t)(A.cv.up_rows + 2 * HALO) * G.pitch +
                            (size_t)(HALO + A.cv.up_rows + (i - G.row_lo)) * G.pitch + a;
                pm_store(d, v0, v1, two);
            }
            if (b) {
                double *d = A.cv.dn_pm[A.out_buf] + (size_t)plane * (size_t)(A.cv.dn_rows + 2 * HALO) * G.pitch +
                            (size_t)(i - (G.row_hi - HALO)) * G.pitch + a;
                pm_store(d, v0, v1, two);
            }
        }
    }
    __threadfence_system();
    __syncwarp();
    if (lane == 0) {
        // boundary CTAs per side: the first HALO rows lie in one segment (segments have >= 4 rows), the last HALO rows
        // in two when the slab's last segment is a single row
        const unsigned int nb = (unsigned int)(ncb * G.count * G.nch);
        const int last_rows = (G.row_hi - G.row_lo) - (nseg - 1) * seg_rows;
        const unsigned int nb_dn = (nseg > 1 && last_rows < HALO) ? 2u * nb : nb;
        if (top) {
            __threadfence();
            if (atomicAdd(&A.cv.box->pm_ticket_up, 1u) == nb - 1u) {
                A.cv.box->pm_ticket_up = 0u;
                __threadfence_system();
                st_release_sys(&A.cv.peer_box[A.cv.rank - 1]->pm_from_below, A.cv.pm_seq);
            }
        }
        if (bot) {
            __threadfence();
            if (atomicAdd(&A.cv.box->pm_ticket_dn, 1u) == nb_dn - 1u) {
                A.cv.box->pm_ticket_dn = 0u;
                __threadfence_system();
                st_release_sys(&A.cv.peer_box[A.cv.rank + 1]->pm_from_above, A.cv.pm_seq);
            }
        }
    }
}

// One warp, between two PM launches of a P2P slab run: wait until both neighbours have pushed launch `need`.
__global__ void pm_wait_kernel(CommBox *box, unsigned int need, int has_up, int has_dn) {
    if (threadIdx.x == 0 && has_up) spin_until(&box->pm_from_above, need, box);
    if (threadIdx.x == 1 && has_dn) spin_until(&box->pm_from_below, need, box);
}

template <typename TIN, typename TOUT, bool STRICT>
__global__ void __launch_bounds__(CTA_THREADS, PM_MIN_CTAS) pm_step_kernel(const __grid_constant__ PmArgs A) {
    const Geom &G = A.g;
    const int lane = threadIdx.x;
    constexpr int warp = 0;
    int bid = blockIdx.x;
    const int cb = bid % G.ncb_pm;
    bid /= G.ncb_pm;
    const int seg = bid % G.pm_nseg;
    const int plane = bid / G.pm_nseg;  // image * nch + channel: channels diffuse independently (:489)
    if (!STRICT) {  // launched with programmatic stream serialization: wait for the previous launch, release the next
        asm volatile("griddepcontrol.wait;" ::: "memory");
        asm volatile("griddepcontrol.launch_dependents;");
    }
    const TIN *__restrict__ in = reinterpret_cast<const TIN *>(A.in) + (size_t)plane * G.plane_elems;
    TOUT *__restrict__ out = reinterpret_cast<TOUT *>(A.out) + (size_t)plane * G.plane_elems;
    const int ra = G.row_lo + seg * G.pm_seg_rows;
    const int rb = min(ra + G.pm_seg_rows, G.row_hi);
    const int cs = cb * PM_CB + warp * PM_STRIP_OWN;
    if (cs >= G.w) return;
    const int a = cs - 2 + 2 * lane;
    const int w = G.w, h = G.h;
    const bool colok = a >= 0 && a < G.pitch;
    const double K = A.K, L = A.L;
    const double inv_k2 = A.inv_k2, lq = L * 0.25;
    // CTAs whose stencils stay inside the image take the fast path (all but the outermost ring)
    const bool interior = !STRICT && cb >= 1 && (cb + 1) * PM_CB + 2 <= w;  // image top/bottom included (replicated halo rows)
    constexpr bool RING = !STRICT && sizeof(TIN) == 8;  // fp64 input planes stream through the shared-memory ring
    __shared__ __align__(16) unsigned char s_ring[RING ? PM_RING_BYTES : 16];
    if (RING) {
        const double *ind = reinterpret_cast<const double *>(in);
        if (rb >= h - 1) {  // rows h-2 and h-1 need g(h-1) = 1
            if (interior)
                pm_rows_ring<TOUT, false, true>(ind, out, G, s_ring, ra, rb, a, lane, inv_k2, lq);
            else
                pm_rows_ring<TOUT, true, true>(ind, out, G, s_ring, ra, rb, a, lane, inv_k2, lq);
        } else if (interior) {
            pm_rows_ring<TOUT, false, false>(ind, out, G, s_ring, ra, rb, a, lane, inv_k2, lq);
        } else {
            pm_rows_ring<TOUT, true, false>(ind, out, G, s_ring, ra, rb, a, lane, inv_k2, lq);
        }
    } else if (interior) {
        pm_rows_fast<TIN, TOUT, false>(in, out, G, ra, rb, a, lane, inv_k2, lq);
    } else if (!STRICT) {
        pm_rows_fast<TIN, TOUT, true>(in, out, G, ra, rb, a, lane, inv_k2, lq);
    } else {
        pm_rows_generic<TIN, TOUT, STRICT>(in, out, G, ra, rb, a, lane, colok, K, L, inv_k2, lq);
    }
    if ((ra == 0 || rb == h) && lane >= 1 && lane <= 30 && a < w) pm_replicate_border<TOUT>(out, G, ra, rb, a);
    if (A.cv.p2p) pm_push_boundary<TOUT>(A, out, plane, ra, rb, a, lane, 1, 30, G.ncb_pm, G.pm_seg_rows, G.pm_nseg);
}

template <int DUMMY>
__global__ void __launch_bounds__(CTA_THREADS, PM2_MIN_CTAS) pm2_step_kernel(const __grid_constant__ PmArgs A) {
    const Geom &G = A.g;
    const int lane = threadIdx.x;
    int bid = blockIdx.x;
    const int cb = bid % G.ncb_pm2;
    bid /= G.ncb_pm2;
    const int seg = bid % G.pm2_nseg;
    const int plane = bid / G.pm2_nseg;
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;");
    const double *__restrict__ in = reinterpret_cast<const double *>(A.in) + (size_t)plane * G.plane_elems;
    double *__restrict__ out = reinterpret_cast<double *>(A.out) + (size_t)plane * G.plane_elems;
    const int ra = G.row_lo + seg * G.pm2_seg_rows;
    const int rb = min(ra + G.pm2_seg_rows, G.row_hi);
    const int cs = cb * PM2_STRIP_OWN;
    if (cs >= G.w) return;
    const int a = cs - 4 + 2 * lane;
    const int w = G.w, h = G.h;
    const double inv_k2 = A.inv_k2, lq = A.L * 0.25;
    const bool interior = cb >= 1 && (cb + 1) * PM2_STRIP_OWN + 4 <= w;
    // a stage forces g(0) or g(h-1) inside its row loop: stage A computes g of rows ra-1 .. rb+2, stage B of ra+1 .. rb
    const bool brows = ra <= 2 || rb + 3 >= h;
    __shared__ __align__(16) unsigned char s_ring[PM2_SMEM];
    unsigned char *jring = s_ring + PM_RING_NS * PM_RING_SLOT;
    if (brows) {
        if (interior)
            pm2_rows_ring<false, true>(in, out, G, s_ring, jring, ra, rb, a, lane, inv_k2, lq);
        else
            pm2_rows_ring<true, true>(in, out, G, s_ring, jring, ra, rb, a, lane, inv_k2, lq);
    } else if (interior) {
        pm2_rows_ring<false, false>(in, out, G, s_ring, jring, ra, rb, a, lane, inv_k2, lq);
    } else {
        pm2_rows_ring<true, false>(in, out, G, s_ring, jring, ra, rb, a, lane, inv_k2, lq);
    }
    if ((ra == 0 || rb == h) && lane >= 2 && lane <= 29 && a < w) pm_replicate_border<double>(out, G, ra, rb, a);
    if (A.cv.p2p) pm_push_boundary<double>(A, out, plane, ra, rb, a, lane, 2, 29, G.ncb_pm2, G.pm2_seg_rows, G.pm2_nseg);
}


__global__ void pm_quantise_kernel(const double *in, uint8_t *out, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += stride) out[q] = (uint8_t)sat_u8(in[q]);
}

template <typename TIN, typename TOUT>
static cudaError_t launch_pm_t(const PmArgs &A, bool strict, cudaStream_t s) {
    const Geom &G = A.g;
    const unsigned int grid = (unsigned int)((size_t)G.count * G.nch * G.pm_nseg * G.ncb_pm);
    if (strict)
        pm_step_kernel<TIN, TOUT, true><<<grid, CTA_THREADS, 0, s>>>(A);
    else {
        // programmatic dependent launch (see csv_kernels.cu): step n+1's CTAs become resident during the tail of step n
        const cudaError_t carve = prefer_max_shared(pm_step_kernel<TIN, TOUT, false>);
        if (carve != cudaSuccess) return carve;
        if (!use_pdl(A.cv.nranks > 1)) {
            pm_step_kernel<TIN, TOUT, false><<<grid, CTA_THREADS, 0, s>>>(A);
            return cudaGetLastError();
        }
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
        return cudaLaunchKernelEx(&cfg, pm_step_kernel<TIN, TOUT, false>, A);
    }
    return cudaGetLastError();
}

cudaError_t launch_pm_step(const PmArgs &A, bool in_u8, bool out_u8, bool strict, cudaStream_t s) {
    if (in_u8 && out_u8) return launch_pm_t<uint8_t, uint8_t>(A, strict, s);
    if (in_u8) return launch_pm_t<uint8_t, double>(A, strict, s);
    if (out_u8) return launch_pm_t<double, uint8_t>(A, strict, s);
    return launch_pm_t<double, double>(A, strict, s);
}

// two fused steps, fp64 planes in and out
cudaError_t launch_pm2_step(const PmArgs &A, cudaStream_t s) {
    const Geom &G = A.g;
    const unsigned int grid = (unsigned int)((size_t)G.count * G.nch * G.pm2_nseg * G.ncb_pm2);
    const cudaError_t carve = prefer_max_shared(pm2_step_kernel<0>);
    if (carve != cudaSuccess) return carve;
    if (!use_pdl(A.cv.nranks > 1)) {
        pm2_step_kernel<0><<<grid, CTA_THREADS, 0, s>>>(A);
        return cudaGetLastError();
    }
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
    return cudaLaunchKernelEx(&cfg, pm2_step_kernel<0>, A);
}

cudaError_t launch_pm_wait(CommBox *box, unsigned int need, int has_up, int has_dn, cudaStream_t s) {
    pm_wait_kernel<<<1, 32, 0, s>>>(box, need, has_up, has_dn);
    return cudaGetLastError();
}

cudaError_t launch_pm_quantise(const double *in, uint8_t *out, size_t n, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    const unsigned int grid = (unsigned int)std::min<size_t>((n + 255) / 256, 148 * 16);
    pm_quantise_kernel<<<grid, 256, 0, s>>>(in, out, n);
    return cudaGetLastError();
}

}  // namespace cvb
