// fp32 variant of the two solvers (CVB_PRECISION_F32): level set and PM state stored and computed in fp32, the
// reductions leave the warp in fp64 and use the same deterministic tree as the fp64 path (reduce.cuh).
// The north star asks for this variant to be REPORTED SEPARATELY: fp32 cannot meet the fp64 level-set tolerance (the
// evolution amplifies rounding differences once |u| >> 1); it is judged on the segmentation mask.
// Same mapping as the fp64 kernels (one warp per CTA, 64-column strips, two columns per lane, halo lanes, row
// recurrences in registers, flux form for PM); algorithmic traffic 8+N B per pixel-iteration (CSV) and 8 B per
// channel-pixel-iteration (PM).
#include "async_copy.cuh"
#include "common.cuh"
#include "kernels.h"
#include "math.cuh"
#include "reduce.cuh"

namespace cvb {

__device__ __forceinline__ float u8_to_float(unsigned int v) { return __int_as_float(0x4B000000u | v) - 8388608.0f; }
__device__ __forceinline__ float2 ldf2(const float *p, bool ok) {
    return ok ? __ldg(reinterpret_cast<const float2 *>(p)) : make_float2(0.0f, 0.0f);
}
// up / sqrt(up^2 + (d/2)^2 + eta^2): MUFU.RSQ (2 ulp) -- as good as fp32 storage of u deserves
__device__ __forceinline__ float normal_f32(float up, float d) {
    float s = fmaf(up, up, 1e-16f);
    s = fmaf(d * d, 0.25f, s);
    return up * rsqrtf(s);
}
// 1/x for x >= eps^2 > 0: MUFU.RCP (1 ulp)
__device__ __forceinline__ float rcp_f32(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// atan(x)/pi, branch-free: r = min(|x|, 1/|x|) in [0, 1], odd minimax polynomial of degree 15 (max error 4e-8 of atan),
// pi/2 - p for |x| > 1, sign restored.  ~18 FP32 instructions; libm's atanf is 2-3x that with a division and branches.
__device__ __forceinline__ float atan_over_pi_f32(float x) {
    const float t = fabsf(x);
    const bool big = t > 1.0f;
    const float r = big ? rcp_f32(t) : t;
    const float w = r * r;
    float p = fmaf(w, -0.004054565913975239f, 0.021862953901290894f);
    p = fmaf(p, w, -0.0559123232960701f);
    p = fmaf(p, w, 0.0964219719171524f);
    p = fmaf(p, w, -0.1390862911939621f);
    p = fmaf(p, w, 0.19946566224098206f);
    p = fmaf(p, w, -0.33329859375953674f);
    p = fmaf(p, w, 0.9999993443489075f);
    p *= r;                                                // atan(r), |error| < 4e-8
    const float a = (big ? 1.57079637f - p : p) * (float)CVB_INV_PI;
    return copysignf(a, x);
}

// ---- CSV step ---------------------------------------------------------------------------------------------------
template <int NCH, bool EDGE>
__device__ __forceinline__ void csv_rows_f32(const float *__restrict__ uin, float *__restrict__ uout,
                                             const uint8_t *__restrict__ im, const Geom &G, const float (&cA)[NCH],
                                             const float (&cB)[NCH], float q0c, float alphap, float eps2, float inv_eps, int ra,
                                             int rb, int a, int lane, double (&acc)[NACC]) {
    const int w = G.w;
    const size_t pitch = (size_t)G.pitch, pe = (size_t)G.plane_elems;
    const bool colok = !EDGE || (a >= 0 && a < G.pitch);
    const bool first = EDGE && a == 0, last0 = EDGE && a == w - 1, last1 = EDGE && a + 1 == w - 1;
    const bool v0 = lane >= 1 && (!EDGE || a < w), v1 = lane >= 1 && (!EDGE || a + 1 < w);
    const bool l31 = lane == 31 && (!EDGE || a + 2 < G.pitch);
    const float *pu = uin + (size_t)(ra - 2 - G.row_lo + HALO) * pitch + a;
    float *po = uout + (size_t)(ra - G.row_lo + HALO) * pitch + a;
    const uint8_t *pi = im + (size_t)(ra - G.row_lo + HALO) * pitch + a;
    auto ldi = [&](const uint8_t *p) -> unsigned int { return colok ? __ldg(reinterpret_cast<const unsigned short *>(p)) : 0u; };

    const float2 R0 = ldf2(pu, colok), R1 = ldf2(pu + pitch, colok);
    float2 C = ldf2(pu + 2 * pitch, colok);
    float e2c = l31 ? __ldg(pu + 2 * pitch + 2) : 0.0f;
    float dN0 = C.x - R1.x, dN1 = C.y - R1.y;
    float nyp0 = normal_f32(dN0, dN0 + (R1.x - R0.x)), nyp1 = normal_f32(dN1, dN1 + (R1.y - R0.y));
    pu += 3 * pitch;
    float2 q0 = ldf2(pu, colok), q1 = make_float2(0.0f, 0.0f);
    if (ra == 0) {  // image top: ny(-1) := ny(0) (src/main.cpp:372); the halo rows hold copies of row 0
        nyp0 = normal_f32(q0.x - C.x, (q0.x - C.x) + dN0);
        nyp1 = normal_f32(q0.y - C.y, (q0.y - C.y) + dN1);
    }
    float f0 = l31 ? __ldg(pu + 2) : 0.0f, f1 = 0.0f;
    unsigned int j0[NCH], j1[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        j0[c] = ldi(pi + c * pe);
        j1[c] = 0u;
    }
    const int n = rb - ra;
    if (n > 1) {
        q1 = ldf2(pu + pitch, colok);
        if (l31) f1 = __ldg(pu + pitch + 2);
#pragma unroll
        for (int c = 0; c < NCH; ++c) j1[c] = ldi(pi + pitch + c * pe);
    }
    pu += 2 * pitch;
    pi += 2 * pitch;
    float accA = 0.0f, accS = 0.0f, accI[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) accI[c] = 0.0f;
    double dA = 0.0, dS = 0.0, dI[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) dI[c] = 0.0;

#pragma unroll 2
    for (int r = 0; r < n; ++r) {
        const float2 S = q0;
        const float e2s = f0;
        unsigned int Ib[NCH];
#pragma unroll
        for (int c = 0; c < NCH; ++c) Ib[c] = j0[c];
        q0 = q1;
        f0 = f1;
#pragma unroll
        for (int c = 0; c < NCH; ++c) j0[c] = j1[c];
        if (r + 2 < n) {
            q1 = ldf2(pu, colok);
            if (l31) f1 = __ldg(pu + 2);
#pragma unroll
            for (int c = 0; c < NCH; ++c) j1[c] = ldi(pi + c * pe);
        }
        if (r + 8 < n) prefetch_l2(pu + 6 * pitch);
        pu += pitch;
        pi += pitch;
        // curvature (:342-375)
        const float upy0 = S.x - C.x, upy1 = S.y - C.y;
        const float ny0 = normal_f32(upy0, upy0 + dN0), ny1 = normal_f32(upy1, upy1 + dN1);
        float Wn = __shfl_up_sync(0xffffffffu, C.y, 1);
        float E2 = __shfl_down_sync(0xffffffffu, C.x, 1);
        E2 = (lane == 31) ? e2c : E2;
        float E0 = C.y;
        if (EDGE) {
            Wn = first ? C.x : Wn;
            E0 = last0 ? C.x : C.y;
            E2 = last1 ? C.y : E2;
        }
        const float nx0 = normal_f32(E0 - C.x, E0 - Wn), nx1 = normal_f32(E2 - C.y, E2 - C.x);
        const float nxw = __shfl_up_sync(0xffffffffu, nx1, 1);
        float kx0 = nx0 - nxw;
        if (EDGE) kx0 = first ? 0.0f : kx0;
        const float kap0 = kx0 + (ny0 - nyp0), kap1 = (nx1 - nx0) + (ny1 - nyp1);
        // data term + combine (:968-985), delta (:988-992), update (:994)
        float I0[NCH], I1[NCH], t0 = q0c, t1 = q0c;
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            I0[c] = u8_to_float(Ib[c] & 0xffu);
            I1[c] = u8_to_float(Ib[c] >> 8);
            t0 = fmaf(fmaf(cA[c], I0[c], cB[c]), I0[c], t0);
            t1 = fmaf(fmaf(cA[c], I1[c], cB[c]), I1[c], t1);
        }
        t0 = fmaf(kap0, alphap, t0);
        t1 = fmaf(kap1, alphap, t1);
        const float du0 = t0 * rcp_f32(fmaf(C.x, C.x, eps2)), du1 = t1 * rcp_f32(fmaf(C.y, C.y, eps2));
        const float un0 = C.x + du0, un1 = C.y + du1;
        if (v1)
            *reinterpret_cast<float2 *>(po) = make_float2(un0, un1);
        else if (v0)
            *po = un0;
        po += pitch;
        // sums of the updated level set and of du^2
        float a0 = atan_over_pi_f32(un0 * inv_eps), a1 = atan_over_pi_f32(un1 * inv_eps);
        float dq0 = du0, dq1 = du1;
        a0 = v0 ? a0 : 0.0f;
        a1 = v1 ? a1 : 0.0f;
        dq0 = v0 ? dq0 : 0.0f;
        dq1 = v1 ? dq1 : 0.0f;
        accA += a0 + a1;
#pragma unroll
        for (int c = 0; c < NCH; ++c) accI[c] = fmaf(I1[c], a1, fmaf(I0[c], a0, accI[c]));
        accS = fmaf(dq1, dq1, fmaf(dq0, dq0, accS));
        if ((r & 15) == 15) {  // keep the fp32 running sums short: fold into fp64 every 16 rows
            dA += (double)accA;
            dS += (double)accS;
            accA = accS = 0.0f;
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                dI[c] += (double)accI[c];
                accI[c] = 0.0f;
            }
        }
        dN0 = upy0;
        dN1 = upy1;
        nyp0 = ny0;
        nyp1 = ny1;
        C = S;
        e2c = e2s;
    }
    acc[ACC_A] = dA + (double)accA;
    acc[ACC_SQ] = dS + (double)accS;
#pragma unroll
    for (int c = 0; c < NCH; ++c) acc[ACC_IA + c] = dI[c] + (double)accI[c];
}

// ---- the same row loop fed by a cp.async shared-memory ring (as csv_rows_ring of the fp64 kernel) ------------------
// The register-prefetch loop above is latency-bound for the reason the fp64 kernels were before their rings: within 128
// registers ptxas sinks the prefetch loads next to their first use.  Rows travel HBM -> shared memory RINGF_NS - 1 rows
// ahead; the east / west neighbours of a row are read from the ring instead of being shuffled.
// Slot of row k (RINGF_SLOT bytes at (k % RINGF_NS) * RINGF_SLOT):
//   [8 + 8*j, +8)           chunk j = 0..32 of the u row: columns cs-2+2j, cs-1+2j (fp32); chunk 32 = lane 31's east neighbour
//   [RINGF_IMG + 80*c, +80) image row of channel c from the 16-byte aligned column (cs-2) & ~15
constexpr int RINGF_NS = 8;
constexpr int RINGF_U = 8;
constexpr int RINGF_IMG = 288;
constexpr int RINGF_SLOT = 544;
constexpr int RINGF_BYTES = RINGF_NS * RINGF_SLOT;
static_assert(RINGF_U + 33 * 8 <= RINGF_IMG && RINGF_IMG + 80 * MAX_CH <= RINGF_SLOT, "fp32 ring slot layout");
static_assert(RINGF_NS - 1 <= TAIL_ROWS, "tail padding too small for the fp32 ring");

template <int NCH, bool EDGE>
__device__ __forceinline__ void csv_rows_ring_f32(const float *__restrict__ uin, float *__restrict__ uout,
                                                  const uint8_t *__restrict__ im, const Geom &G, const float (&cA)[NCH],
                                                  const float (&cB)[NCH], float q0c, float alphap, float eps2, float inv_eps,
                                                  unsigned char *ring, int ra, int rb, int cs, int lane, double (&acc)[NACC]) {
    const int w = G.w;
    const int a = cs - 2 + 2 * lane;
    const size_t pitch = (size_t)G.pitch, pe = (size_t)G.plane_elems;
    const bool first = EDGE && a == 0, last0 = EDGE && a == w - 1, last1 = EDGE && a + 1 == w - 1;
    const bool v0 = lane >= 1 && (!EDGE || a < w), v1 = lane >= 1 && (!EDGE || a + 1 < w);
    const bool ok1 = !EDGE || (a >= 0 && a < G.pitch);
    // copies of one row: every lane its own 8-byte chunk of u; lanes 0 .. 5*NCH-1 a 16-byte chunk of the image strips;
    // lane 15 the 33rd u chunk
    const int s_al = (cs - 2) & ~15, dsh = (cs - 2) - s_al;
    const float *src1 = uin + (size_t)(ra - G.row_lo + HALO) * pitch + a;  // row ra, advanced by pitch per issued row
    const uint8_t *src2 = nullptr;
    unsigned int dst2 = 0;
    bool ok2 = false;
    if (lane < 5 * NCH) {
        const int c = lane / 5, q = lane % 5, col = s_al + 16 * q;
        src2 = im + (size_t)c * pe + (size_t)(ra - G.row_lo + HALO) * pitch + col;
        dst2 = RINGF_IMG + 80 * c + 16 * q;
        ok2 = !EDGE || (col >= 0 && col < G.pitch);
    }
    const bool ok3 = lane == 15 && (!EDGE || cs - 2 + 64 < G.pitch);
    const float *src3 = uin + (size_t)(ra - G.row_lo + HALO) * pitch + (cs - 2 + 64);
    const unsigned int ring_s = (unsigned int)__cvta_generic_to_shared(ring);
    auto issue = [&](unsigned int slot_off) {
        if (ok1) cp_async8(ring_s + slot_off + RINGF_U + 8 * lane, src1);
        if (ok2) cp_async16(ring_s + slot_off + dst2, src2);
        if (ok3) cp_async8(ring_s + slot_off + RINGF_U + 8 * 32, src3);
        cp_async_commit();
        src1 += pitch;
        src2 += pitch;
        src3 += pitch;
    };
#pragma unroll
    for (int k = 0; k < RINGF_NS - 1; ++k) issue(k * RINGF_SLOT);
    // rows ra-2, ra-1 (own columns only) straight from global memory
    const float *pr = uin + (size_t)(ra - 2 - G.row_lo + HALO) * pitch + a;
    const float2 R0 = ldf2(pr, ok1), R1 = ldf2(pr + pitch, ok1);
    const unsigned char *my8 = ring + RINGF_U + 8 * lane;  // + slot: own chunk; west neighbour at -4, east at +8
    const unsigned char *my2 = ring + RINGF_IMG + dsh + 2 * lane;  // + slot + 80 c: own two image bytes
    cp_async_wait<RINGF_NS - 3>();  // rows ra and ra+1 have landed
    __syncwarp();
    float2 C = *reinterpret_cast<const float2 *>(my8);
    float CW = *reinterpret_cast<const float *>(my8 - 4);
    float CE = *reinterpret_cast<const float *>(my8 + 8);
    float dN0 = C.x - R1.x, dN1 = C.y - R1.y;
    float nyp0 = normal_f32(dN0, dN0 + (R1.x - R0.x)), nyp1 = normal_f32(dN1, dN1 + (R1.y - R0.y));
    if (ra == 0) {  // image top: ny(-1) := ny(0) (src/main.cpp:372)
        const float2 q = *reinterpret_cast<const float2 *>(my8 + RINGF_SLOT);
        nyp0 = normal_f32(q.x - C.x, (q.x - C.x) + dN0);
        nyp1 = normal_f32(q.y - C.y, (q.y - C.y) + dN1);
    }
    float accA = 0.0f, accS = 0.0f, accI[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) accI[c] = 0.0f;
    double dA = 0.0, dS = 0.0, dI[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) dI[c] = 0.0;
    float *po = uout + (size_t)(ra - G.row_lo + HALO) * pitch + a;
    const int n = rb - ra;

    // one row: s_wr = slot of row i-1 (free, receives row i+NS-1), s_img = slot of row i, s_u = slot of row i+1
    auto row = [&](unsigned int s_wr, unsigned int s_img, unsigned int s_u) {
        issue(s_wr);
        cp_async_wait<RINGF_NS - 2>();  // row i+1 has landed
        __syncwarp();
        const float2 S = *reinterpret_cast<const float2 *>(my8 + s_u);
        const float SW = *reinterpret_cast<const float *>(my8 + s_u - 4);
        const float SE = *reinterpret_cast<const float *>(my8 + s_u + 8);
        unsigned int Ib[NCH];
#pragma unroll
        for (int c = 0; c < NCH; ++c) Ib[c] = *reinterpret_cast<const unsigned short *>(my2 + s_img + 80 * c);
        // curvature (:342-375)
        const float upy0 = S.x - C.x, upy1 = S.y - C.y;
        const float ny0 = normal_f32(upy0, upy0 + dN0), ny1 = normal_f32(upy1, upy1 + dN1);
        float Wn = CW, E2 = CE, E0 = C.y;
        if (EDGE) {
            Wn = first ? C.x : Wn;
            E0 = last0 ? C.x : C.y;
            E2 = last1 ? C.y : E2;
        }
        const float nx0 = normal_f32(E0 - C.x, E0 - Wn), nx1 = normal_f32(E2 - C.y, E2 - C.x);
        const float nxw = __shfl_up_sync(0xffffffffu, nx1, 1);
        float kx0 = nx0 - nxw;
        if (EDGE) kx0 = first ? 0.0f : kx0;
        const float kap0 = kx0 + (ny0 - nyp0), kap1 = (nx1 - nx0) + (ny1 - nyp1);
        // data term + combine (:968-985), delta (:988-992), update (:994)
        float I0[NCH], I1[NCH], t0 = q0c, t1 = q0c;
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            I0[c] = u8_to_float(Ib[c] & 0xffu);
            I1[c] = u8_to_float(Ib[c] >> 8);
            t0 = fmaf(fmaf(cA[c], I0[c], cB[c]), I0[c], t0);
            t1 = fmaf(fmaf(cA[c], I1[c], cB[c]), I1[c], t1);
        }
        t0 = fmaf(kap0, alphap, t0);
        t1 = fmaf(kap1, alphap, t1);
        const float du0 = t0 * rcp_f32(fmaf(C.x, C.x, eps2)), du1 = t1 * rcp_f32(fmaf(C.y, C.y, eps2));
        const float un0 = C.x + du0, un1 = C.y + du1;
        if (EDGE) {
            if (v1)
                *reinterpret_cast<float2 *>(po) = make_float2(un0, un1);
            else if (v0)
                *po = un0;
        } else if (lane) {
            *reinterpret_cast<float2 *>(po) = make_float2(un0, un1);
        }
        po += pitch;
        // sums of the updated level set and of du^2
        float a0 = atan_over_pi_f32(un0 * inv_eps), a1 = atan_over_pi_f32(un1 * inv_eps);
        float dq0 = du0, dq1 = du1;
        if (EDGE) {
            a0 = v0 ? a0 : 0.0f;
            a1 = v1 ? a1 : 0.0f;
            dq0 = v0 ? dq0 : 0.0f;
            dq1 = v1 ? dq1 : 0.0f;
        }
        accA += a0 + a1;
#pragma unroll
        for (int c = 0; c < NCH; ++c) accI[c] = fmaf(I1[c], a1, fmaf(I0[c], a0, accI[c]));
        accS = fmaf(dq1, dq1, fmaf(dq0, dq0, accS));
        dN0 = upy0;
        dN1 = upy1;
        nyp0 = ny0;
        nyp1 = ny1;
        C = S;
        CW = SW;
        CE = SE;
    };
    auto fold = [&]() {  // keep the fp32 running sums short: into fp64 every 16 rows
        dA += (double)accA;
        dS += (double)accS;
        accA = accS = 0.0f;
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            dI[c] += (double)accI[c];
            accI[c] = 0.0f;
        }
    };
    int r = 0;
    unsigned int tog = 0;  // offset of slot 0 or slot 4: the slot of the first row of a group of four
#pragma unroll 1
    for (; r + 4 <= n; r += 4) {
        const unsigned int t2 = tog ^ (4 * RINGF_SLOT);
        row(t2 + 3 * RINGF_SLOT, tog, tog + RINGF_SLOT);
        row(tog, tog + RINGF_SLOT, tog + 2 * RINGF_SLOT);
        row(tog + RINGF_SLOT, tog + 2 * RINGF_SLOT, tog + 3 * RINGF_SLOT);
        row(tog + 2 * RINGF_SLOT, tog + 3 * RINGF_SLOT, t2);
        tog = t2;
        if ((r & 12) == 12) fold();
    }
#pragma unroll 1
    for (; r < n; ++r) {
        const unsigned int k = (unsigned int)r;
        row(((k + RINGF_NS - 1) % RINGF_NS) * RINGF_SLOT, (k % RINGF_NS) * RINGF_SLOT, ((k + 1) % RINGF_NS) * RINGF_SLOT);
    }
    cp_async_wait<0>();  // nothing may land in the ring after the CTA has gone
    acc[ACC_A] = dA + (double)accA;
    acc[ACC_SQ] = dS + (double)accS;
#pragma unroll
    for (int c = 0; c < NCH; ++c) acc[ACC_IA + c] = dI[c] + (double)accI[c];
}

__device__ __noinline__ void replicate_border_rows_f32(float *uout, const Geom &G, int ra, int rb, int a) {
    if (ra == 0) {
        const float2 v = __ldcg(reinterpret_cast<const float2 *>(uout + (size_t)(0 - G.row_lo + HALO) * G.pitch + a));
        for (int k = 0; k < HALO; ++k) *reinterpret_cast<float2 *>(uout + (size_t)k * G.pitch + a) = v;
    }
    if (rb == G.h) {
        const size_t last = (size_t)(G.h - 1 - G.row_lo + HALO);
        const float2 v = __ldcg(reinterpret_cast<const float2 *>(uout + last * G.pitch + a));
        for (int k = 1; k <= HALO; ++k) *reinterpret_cast<float2 *>(uout + (last + k) * G.pitch + a) = v;
    }
}

template <int NCH>
__global__ void __launch_bounds__(CTA_THREADS, 16) csv_step_f32_kernel(const __grid_constant__ CsvArgs A) {
    const Geom &G = A.g;
    const int lane = threadIdx.x;
    int bid = blockIdx.x;
    const int cb = bid % G.ncb_csv;
    bid /= G.ncb_csv;
    const int seg = bid % G.nseg;
    const int img = bid / G.nseg;
    CsvState *st = A.state + img;
    const int2 ds = *reinterpret_cast<const int2 *>(&st->done);
    if (ds.x) return;
    const int par = A.par;
    const float *uin = reinterpret_cast<const float *>(A.u[par]) + (size_t)img * G.plane_elems;
    float *uout = reinterpret_cast<float *>(A.u[par ^ 1]) + (size_t)img * G.plane_elems;
    const uint8_t *im = A.img + (size_t)img * G.nch * G.plane_elems;
    const int gseg = G.seg0 + seg;
    const int ra = max(gseg * G.seg_rows, G.row_lo), rb = min((gseg + 1) * G.seg_rows, G.row_hi);
    const int cs = cb * CSV_CB;
    const int a = cs - 2 + 2 * lane;
    // per-step coefficients in fp64, then rounded once
    const double kd = A.eps * CVB_INV_PI, bk = A.beta * kd;
    double q0 = A.gamma * kd;
    float cA[NCH], cB[NCH];
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
        const double c1 = st->c1[k], c2 = st->c2[k], l1 = A.lambda1[k], l2 = A.lambda2[k];
        cA[k] = (float)(bk * (l2 - l1));
        cB[k] = (float)(2.0 * bk * (l1 * c1 - l2 * c2));
        q0 += bk * (l2 * c2 * c2 - l1 * c1 * c1);
    }
    double acc[NACC];
#pragma unroll
    for (int v = 0; v < NACC; ++v) acc[v] = 0.0;
    const bool interior = cb > 0 && (cb + 1) * CSV_CB < G.w;
    __shared__ __align__(16) unsigned char s_ring[RINGF_BYTES];
#ifdef CVB_F32_NO_RING  // the register-prefetch loop of round 1 (comparison builds)
    if (interior)
        csv_rows_f32<NCH, false>(uin, uout, im, G, cA, cB, (float)q0, (float)(A.alpha * kd), (float)(A.eps * A.eps),
                                 (float)A.inv_eps, ra, rb, a, lane, acc);
    else
        csv_rows_f32<NCH, true>(uin, uout, im, G, cA, cB, (float)q0, (float)(A.alpha * kd), (float)(A.eps * A.eps),
                                (float)A.inv_eps, ra, rb, a, lane, acc);
#else
    if (interior)
        csv_rows_ring_f32<NCH, false>(uin, uout, im, G, cA, cB, (float)q0, (float)(A.alpha * kd), (float)(A.eps * A.eps),
                                      (float)A.inv_eps, s_ring, ra, rb, cs, lane, acc);
    else
        csv_rows_ring_f32<NCH, true>(uin, uout, im, G, cA, cB, (float)q0, (float)(A.alpha * kd), (float)(A.eps * A.eps),
                                     (float)A.inv_eps, s_ring, ra, rb, cs, lane, acc);
#endif
    if (lane == 0) {
#pragma unroll
        for (int v = 0; v < NACC; ++v) acc[v] = 0.0;  // halo lane
    }
    if (a >= 0 && a < G.pitch && lane >= 1 && (ra == 0 || rb == G.h)) replicate_border_rows_f32(uout, G, ra, rb, a);
    finish_tile<NCH, false>(A, img, seg, cb, G.ncb_csv, acc, 0);
}

template <int NCH>
__global__ void __launch_bounds__(CTA_THREADS, 16) csv_init_f32_kernel(const __grid_constant__ CsvArgs A, int final_mode) {
    const Geom &G = A.g;
    const int lane = threadIdx.x;
    int bid = blockIdx.x;
    const int cb = bid % G.ncb_csv;
    bid /= G.ncb_csv;
    const int seg = bid % G.nseg;
    const int img = bid / G.nseg;
    CsvState *st = A.state + img;
    const int par = (final_mode == 1) ? 0 : (st->steps_done & 1);
    const float *uin = reinterpret_cast<const float *>(A.u[par]) + (size_t)img * G.plane_elems;
    const uint8_t *im = A.img + (size_t)img * G.nch * G.plane_elems;
    const int gseg = G.seg0 + seg;
    const int ra = max(gseg * G.seg_rows, G.row_lo), rb = min((gseg + 1) * G.seg_rows, G.row_hi);
    const int a = cb * CSV_CB - 2 + 2 * lane;
    const float inv_eps = (float)A.inv_eps;
    const double inv_n = 1.0 / (double)NCH;
    double acc[NACC];
#pragma unroll
    for (int v = 0; v < NACC; ++v) acc[v] = 0.0;
    if (lane >= 1 && a < G.w) {
        const bool v1 = a + 1 < G.w;
        for (int i = ra; i < rb; ++i) {
            const size_t off = (size_t)(i - G.row_lo + HALO) * G.pitch + a;
            const float2 U = __ldg(reinterpret_cast<const float2 *>(uin + off));
            const double a0 = (double)atan_over_pi_f32(U.x * inv_eps);
            const double a1 = v1 ? (double)atan_over_pi_f32(U.y * inv_eps) : 0.0;
            acc[ACC_A] += a0 + a1;
            double m0 = 0.0, m1 = 0.0;
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                const unsigned int b = __ldg(reinterpret_cast<const unsigned short *>(im + (size_t)c * G.plane_elems + off));
                const double I0 = u8_to_double(b & 0xffu), I1 = v1 ? u8_to_double(b >> 8) : 0.0;
                acc[ACC_IA + c] = fma(I1, a1, fma(I0, a0, acc[ACC_IA + c]));
                acc[ACC_I + c] += I0 + I1;
                m0 += I0;
                m1 += I1;
            }
            m0 *= inv_n;
            m1 *= inv_n;
            acc[ACC_SQ] = fma(m1, m1, fma(m0, m0, acc[ACC_SQ]));
        }
    }
    finish_tile<NCH, true>(A, img, seg, cb, G.ncb_csv, acc, final_mode);
}

// ---- PM step ----------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ float2 pmf_load(const T *p);
template <>
__device__ __forceinline__ float2 pmf_load<float>(const float *p) {
    return __ldg(reinterpret_cast<const float2 *>(p));
}
template <>
__device__ __forceinline__ float2 pmf_load<uint8_t>(const uint8_t *p) {
    const unsigned int b = __ldg(reinterpret_cast<const unsigned short *>(p));
    return make_float2(u8_to_float(b & 0xffu), u8_to_float(b >> 8));
}
__device__ __forceinline__ unsigned int sat_u8f(float v) { return (unsigned int)min(max(__float2int_rn(v), 0), 255); }
__device__ __forceinline__ void pmf_store(float *p, float x, float y, bool two) {
    if (two)
        *reinterpret_cast<float2 *>(p) = make_float2(x, y);
    else
        *p = x;
}
__device__ __forceinline__ void pmf_store(uint8_t *p, float x, float y, bool two) {
    if (two)
        *reinterpret_cast<unsigned short *>(p) = (unsigned short)(sat_u8f(x) | (sat_u8f(y) << 8));
    else
        *p = (uint8_t)sat_u8f(x);
}

template <typename TIN, typename TOUT>
__global__ void __launch_bounds__(CTA_THREADS, 16) pm_step_f32_kernel(const __grid_constant__ PmArgs A) {
    const Geom &G = A.g;
    const int lane = threadIdx.x;
    int bid = blockIdx.x;
    const int cb = bid % G.ncb_pm;
    bid /= G.ncb_pm;
    const int seg = bid % G.pm_nseg;
    const int plane = bid / G.pm_nseg;
    const TIN *in = reinterpret_cast<const TIN *>(A.in) + (size_t)plane * G.plane_elems;
    TOUT *out = reinterpret_cast<TOUT *>(A.out) + (size_t)plane * G.plane_elems;
    const int ra = G.row_lo + seg * G.pm_seg_rows, rb = min(ra + G.pm_seg_rows, G.row_hi);
    const int a = cb * PM_CB - 2 + 2 * lane;
    const int w = G.w, h = G.h;
    const size_t pitch = (size_t)G.pitch;
    const bool colok = a >= 0 && a < G.pitch;
    const bool bc0 = a == 0 || a == w - 1, bc1 = a + 1 == w - 1;
    const bool nofxw = a == 0, nofx0 = a == w - 1;
    const float inv_k2 = (float)A.inv_k2, lq = (float)(A.L * 0.25);
    auto ldr = [&](const TIN *p) { return colok ? pmf_load<TIN>(p) : make_float2(0.0f, 0.0f); };
    auto sobel_rows = [&](const float2 &X, float2 &rd, float2 &rs) {
        const float Wn = __shfl_up_sync(0xffffffffu, X.y, 1), E2 = __shfl_down_sync(0xffffffffu, X.x, 1);
        rd.x = X.y - Wn;
        rs.x = fmaf(2.0f, X.x, Wn) + X.y;
        rd.y = E2 - X.x;
        rs.y = fmaf(2.0f, X.y, X.x) + E2;
    };
    auto edge = [&](float gx, float gy) { return rcp_f32(fmaf(fmaf(gx, gx, gy * gy), inv_k2, 1.0f)); };
    auto fixg = [&](float2 &g) {
        g.x = bc0 ? 1.0f : g.x;
        g.y = bc1 ? 1.0f : g.y;
    };
    const TIN *pin = in + (size_t)(ra - 2 - G.row_lo + HALO) * pitch + a;
    TOUT *po = out + (size_t)(ra - G.row_lo + HALO) * pitch + a;
    const int n = rb - ra;
    const float2 X0 = ldr(pin), X1 = ldr(pin + pitch), X2 = ldr(pin + 2 * pitch), X3 = ldr(pin + 3 * pitch);
    float2 rd0, rs0, rd1, rs1, rd2, rs2, rd3, rs3;
    sobel_rows(X0, rd0, rs0);
    sobel_rows(X1, rd1, rs1);
    sobel_rows(X2, rd2, rs2);
    sobel_rows(X3, rd3, rs3);
    float2 gP, gC;
    gP.x = edge((rd0.x + 2.0f * rd1.x) + rd2.x, rs2.x - rs0.x);
    gP.y = edge((rd0.y + 2.0f * rd1.y) + rd2.y, rs2.y - rs0.y);
    gC.x = edge((rd1.x + 2.0f * rd2.x) + rd3.x, rs3.x - rs1.x);
    gC.y = edge((rd1.y + 2.0f * rd2.y) + rd3.y, rs3.y - rs1.y);
    if (ra == 0 || ra == h - 1) gC = make_float2(1.0f, 1.0f);
    fixg(gP);
    fixg(gC);
    float fy0 = (gP.x + gC.x) * (X2.x - X1.x), fy1 = (gP.y + gC.y) * (X2.y - X1.y);
    float2 IC = X2, IS = X3;
    float2 P = make_float2(fmaf(2.0f, rd3.x, rd2.x), fmaf(2.0f, rd3.y, rd2.y));
    float2 rdB = rd3, rsA = rs2, rsB = rs3;
    pin += 4 * pitch;
    float2 q0 = ldr(pin), q1 = make_float2(0.0f, 0.0f);
    if (n > 1) q1 = ldr(pin + pitch);
    pin += 2 * pitch;
#pragma unroll 2
    for (int r = 0; r < n; ++r) {
        const float2 X = q0;
        q0 = q1;
        if (r + 2 < n) q1 = ldr(pin);
        if (r + 8 < n && colok) prefetch_l2(pin + 6 * pitch);
        pin += pitch;
        float2 rdC, rsC, gS;
        sobel_rows(X, rdC, rsC);
        gS.x = edge(P.x + rdC.x, rsC.x - rsA.x);
        gS.y = edge(P.y + rdC.y, rsC.y - rsA.y);
        if (ra + r + 1 == h - 1) gS = make_float2(1.0f, 1.0f);
        fixg(gS);
        const float fs0 = (gC.x + gS.x) * (IS.x - IC.x), fs1 = (gC.y + gS.y) * (IS.y - IC.y);
        const float Ie = __shfl_down_sync(0xffffffffu, IC.x, 1), ge = __shfl_down_sync(0xffffffffu, gC.x, 1);
        float fx0 = (gC.x + gC.y) * (IC.y - IC.x), fx1 = (gC.y + ge) * (Ie - IC.y);
        fx0 = nofx0 ? 0.0f : fx0;
        fx1 = bc1 ? 0.0f : fx1;
        float fxw = __shfl_up_sync(0xffffffffu, fx1, 1);
        fxw = nofxw ? 0.0f : fxw;
        const float o0 = fmaf((fs0 - fy0) + (fx0 - fxw), lq, IC.x), o1 = fmaf((fs1 - fy1) + (fx1 - fx0), lq, IC.y);
        if (lane >= 1 && lane <= 30 && a < w) pmf_store(po, o0, o1, a + 1 < w);
        po += pitch;
        fy0 = fs0;
        fy1 = fs1;
        P.x = fmaf(2.0f, rdC.x, rdB.x);
        P.y = fmaf(2.0f, rdC.y, rdB.y);
        rdB = rdC;
        rsA = rsB;
        rsB = rsC;
        IC = IS;
        IS = X;
        gC = gS;
    }
    // clamped neighbours in i: copies of the border rows in the halo rows of the plane just written
    if ((ra == 0 || rb == h) && lane >= 1 && lane <= 30 && a < w) {
        const bool two = a + 1 < w;
        if (ra == 0) {
            const TOUT *src = out + (size_t)(0 - G.row_lo + HALO) * pitch + a;
            const TOUT s0 = __ldcg(src), s1 = two ? __ldcg(src + 1) : s0;
            for (int k = 0; k < HALO; ++k) {
                out[(size_t)k * pitch + a] = s0;
                if (two) out[(size_t)k * pitch + a + 1] = s1;
            }
        }
        if (rb == h) {
            const size_t last = (size_t)(h - 1 - G.row_lo + HALO);
            const TOUT *src = out + last * pitch + a;
            const TOUT s0 = __ldcg(src), s1 = two ? __ldcg(src + 1) : s0;
            for (int k = 1; k <= HALO; ++k) {
                out[(last + k) * pitch + a] = s0;
                if (two) out[(last + k) * pitch + a + 1] = s1;
            }
        }
    }
}

// ---- small kernels ------------------------------------------------------------------------------------------------
__global__ void convert_d2f_kernel(const double *in, float *out, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += stride) out[q] = (float)in[q];
}
__global__ void convert_f2d_kernel(const float *in, double *out, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += stride) out[q] = (double)in[q];
}
__global__ void mask_f32_kernel(const float *u, uint8_t *mask, int rows, int w, int pitch, int invert) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= w) return;
    for (int i = blockIdx.y; i < rows; i += gridDim.y) {
        const size_t q = (size_t)i * pitch + j;
        const uint8_t m = (u[q] > 0.0f) ? 1 : 0;  // separate(): float32(u) > 0 (src/main.cpp:395-400)
        mask[q] = invert ? (uint8_t)(1 - m) : m;
    }
}
__global__ void checkerboard_f32_kernel(float *u, const signed char *si, const signed char *sj, int row_lo, int rows, int w,
                                        int pitch) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= w) return;
    for (int i = blockIdx.y; i < rows; i += gridDim.y)
        u[(size_t)(i + HALO) * pitch + j] = (float)((int)si[row_lo + i] * (int)sj[j]);
}
__global__ void quantise_f32_kernel(const float *in, uint8_t *out, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += stride) out[q] = (uint8_t)sat_u8f(in[q]);
}

// ---- host launchers -------------------------------------------------------------------------------------------------
cudaError_t launch_csv_step_f32(const CsvArgs &A, cudaStream_t s) {
    const Geom &G = A.g;
    const unsigned int grid = (unsigned int)((size_t)G.count * G.nseg * G.ncb_csv);
    if (G.nch == 1)
        csv_step_f32_kernel<1><<<grid, CTA_THREADS, 0, s>>>(A);
    else
        csv_step_f32_kernel<3><<<grid, CTA_THREADS, 0, s>>>(A);
    return cudaGetLastError();
}
cudaError_t launch_csv_init_f32(const CsvArgs &A, int final_mode, cudaStream_t s) {
    const Geom &G = A.g;
    const unsigned int grid = (unsigned int)((size_t)G.count * G.nseg * G.ncb_csv);
    if (G.nch == 1)
        csv_init_f32_kernel<1><<<grid, CTA_THREADS, 0, s>>>(A, final_mode);
    else
        csv_init_f32_kernel<3><<<grid, CTA_THREADS, 0, s>>>(A, final_mode);
    return cudaGetLastError();
}
cudaError_t launch_pm_step_f32(const PmArgs &A, bool in_u8, bool out_u8, cudaStream_t s) {
    const Geom &G = A.g;
    const unsigned int grid = (unsigned int)((size_t)G.count * G.nch * G.pm_nseg * G.ncb_pm);
    if (in_u8 && out_u8)
        pm_step_f32_kernel<uint8_t, uint8_t><<<grid, CTA_THREADS, 0, s>>>(A);
    else if (in_u8)
        pm_step_f32_kernel<uint8_t, float><<<grid, CTA_THREADS, 0, s>>>(A);
    else if (out_u8)
        pm_step_f32_kernel<float, uint8_t><<<grid, CTA_THREADS, 0, s>>>(A);
    else
        pm_step_f32_kernel<float, float><<<grid, CTA_THREADS, 0, s>>>(A);
    return cudaGetLastError();
}
static unsigned int flat_grid(size_t n) { return (unsigned int)std::min<size_t>((n + 255) / 256, 148 * 16); }
cudaError_t launch_convert_d2f(const double *in, float *out, size_t n, cudaStream_t s) {
    if (n) convert_d2f_kernel<<<flat_grid(n), 256, 0, s>>>(in, out, n);
    return cudaGetLastError();
}
cudaError_t launch_convert_f2d(const float *in, double *out, size_t n, cudaStream_t s) {
    if (n) convert_f2d_kernel<<<flat_grid(n), 256, 0, s>>>(in, out, n);
    return cudaGetLastError();
}
cudaError_t launch_mask_f32(const float *u, uint8_t *mask, int rows, int w, int pitch, int invert, cudaStream_t s) {
    dim3 grid((w + 255) / 256, std::min(rows, 65535));
    mask_f32_kernel<<<grid, 256, 0, s>>>(u, mask, rows, w, pitch, invert);
    return cudaGetLastError();
}
cudaError_t launch_checkerboard_f32(float *u, const signed char *si, const signed char *sj, int row_lo, int rows, int w,
                                    int pitch, cudaStream_t s) {
    dim3 grid((w + 255) / 256, std::min(rows, 65535));
    checkerboard_f32_kernel<<<grid, 256, 0, s>>>(u, si, sj, row_lo, rows, w, pitch);
    return cudaGetLastError();
}
cudaError_t launch_quantise_f32(const float *in, uint8_t *out, size_t n, cudaStream_t s) {
    if (n) quantise_f32_kernel<<<flat_grid(n), 256, 0, s>>>(in, out, n);
    return cudaGetLastError();
}

}  // namespace cvb
