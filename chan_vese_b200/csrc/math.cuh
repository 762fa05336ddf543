// Device math for the fp64 solvers.
//
// The fused CSV step is bound by the FP64 pipe (64 lanes/clk/SM on B200), not by HBM, once the HBM
// traffic is down to 16+N bytes/pixel, so the production ("fast") math counts FP64 instructions:
//   * reciprocal / rsqrt = MUFU seed (rcp.approx.ftz.f64 / rsqrt.approx.ftz.f64) + ONE cubic
//     Newton step (3 resp. 5 DFMA-class ops, ~1 ulp), no IEEE division or sqrt sequences;
//   * atan(x)/pi by a 192-interval table-driven argument reduction with one reciprocal and a degree-4
//     polynomial (16 FP64 ops, <= 4 ulp, branch-free; the CUDA libm atan is ~2x that);
//   * uint8 -> double through the 2^52 magic constant (1 DADD instead of a quarter-rate I2F.F64).
// "Strict" math reproduces the oracle's operation order with IEEE div/sqrt and no FMA contraction
// (the reference build has no FMA and forbids reassociation, Makefile:16): a test mode.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cvb {

#define CVB_PI 3.14159265358979323846
#define CVB_INV_PI 0.31830988618379067154

__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

__device__ __forceinline__ double rcp_seed(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y;
}
__device__ __forceinline__ double rsqrt_seed(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y;
}
// 1/x for normal positive x: seed (>= 20 bits) + cubic step, 3 FP64 ops
__device__ __forceinline__ double fast_rcp(double x) {
    const double y = rcp_seed(x);
    double e = fma(-x, y, 1.0);
    e = fma(e, e, e);
    return fma(y, e, y);
}
// 1/sqrt(x) for normal positive x: seed + cubic step, 5 FP64 ops
__device__ __forceinline__ double fast_rsqrt(double x) {
    const double y = rsqrt_seed(x);
    const double t = x * y;
    const double e = fma(-t, y, 1.0);
    const double p = fma(e, 0.375, 0.5);
    const double ye = y * e;
    return fma(ye, p, y);
}

// uint8 -> double, exact: bits(2^52 + v) - 2^52
__device__ __forceinline__ double u8_to_double(unsigned int v) {
    return __hiloint2double(0x43300000, (int)v) - 4503599627370496.0;
}

// ---- atan(x)/pi ------------------------------------------------------------------------------------
// t = |x| clamped to [2^-4, 2^44): c = centre of t's quarter-octave (exponent and top two mantissa bits kept, third
// bit set); atan t = atan c + atan z with z = (t-c)/(1+t*c), |z| <= 0.0704 (for t < 2^-4 the first centre is
// shared: the ABSOLUTE error stays ~1e-17, which is what the sums of a = H - 1/2 need; for t >= 2^44 the last centre
// is shared: z < 6e-14).  No branches, no selects: tab[q] = atan(c_q)/pi, q = 0..191, lives in shared memory.
constexpr int ATAN_NQ = 192;
constexpr int ATAN_TAB_N = ATAN_NQ;
constexpr int ATAN_Q0 = 1019 * 4;                        // (biased exponent of 2^-4) * 4
constexpr int ATAN_HC_MIN = 0x3fb00000;                  // hi word of 2^-4
constexpr int ATAN_HC_MAX = ((1023 + 43) << 20) | 0xfffff;  // hi word just below 2^44

__device__ __forceinline__ void atan_reduce(double x, const double *tab, int &hx, double &num, double &den, double &hi) {
    hx = __double2hiint(x);
    const int ht = hx & 0x7fffffff;
    const double t = __hiloint2double(ht, __double2loint(x));
    const int hc = min(max(ht, ATAN_HC_MIN), ATAN_HC_MAX);
    const double c = __hiloint2double((hc & 0xfffc0000) | 0x00020000, 0);
    hi = tab[(hc >> 18) - ATAN_Q0];
    num = t - c;
    den = fma(t, c, 1.0);
}
// the same against a table of {atan(c_q)/pi, c_q} pairs (one LDS.128, no integer construction of c) and with |x|
// as an operand modifier of the two FP64 instructions
__device__ __forceinline__ void atan_reduce(double x, const double2 *tab, int &hx, double &num, double &den, double &hi) {
    hx = __double2hiint(x);
    const int hc = min(max(hx & 0x7fffffff, ATAN_HC_MIN), ATAN_HC_MAX);
    const double2 e = tab[(hc >> 18) - ATAN_Q0];
    const double t = fabs(x);
    hi = e.x;
    num = t - e.y;
    den = fma(t, e.y, 1.0);
}
__device__ __forceinline__ double atan_centre(int q) {  // c_q, q = 0..ATAN_NQ-1
    return __hiloint2double(((ATAN_Q0 + q) << 18) | 0x00020000, 0);
}
__device__ __forceinline__ double atan_finish(double z, double hi, int hx) {
    const double w = z * z;
    // (atan z / z - 1) / w on w in [0, 0.0704^2]: degree-4 least-squares fit on Chebyshev nodes, max error 4.4e-16
    // (a 2e-18 relative error of atan z); one FMA shorter than the Taylor series
    double p = fma(w, -0.089962697680458735129, 0.11110701662127189912);
    p = fma(p, w, -0.14285713561821073646);
    p = fma(p, w, 0.19999999999551851066);
    p = fma(p, w, -0.33333333333333288932);
    const double zw = z * w;
    const double at = fma(zw, p, z);           // atan z
    const double r = fma(at, CVB_INV_PI, hi);  // atan(t)/pi
    return __hiloint2double(__double2hiint(r) ^ (hx & 0x80000000), __double2loint(r));
}
template <typename TAB>
__device__ __forceinline__ double atan_over_pi(double x, const TAB *tab /* shared memory */) {
    int hx;
    double num, den, hi;
    atan_reduce(x, tab, hx, num, den, hi);
    return atan_finish(num * fast_rcp(den), hi, hx);
}
// two arguments sharing ONE reciprocal: 1/d0 = d1/(d0*d1), 1/d1 = d0/(d0*d1)  (d in [1, 2^90): no overflow)
template <typename TAB>
__device__ __forceinline__ void atan_over_pi2(double x0, double x1, const TAB *tab, double &a0, double &a1) {
    int hx0, hx1;
    double n0, d0, h0, n1, d1, h1;
    atan_reduce(x0, tab, hx0, n0, d0, h0);
    atan_reduce(x1, tab, hx1, n1, d1, h1);
    const double r = fast_rcp(d0 * d1);
    a0 = atan_finish(n0 * (r * d1), h0, hx0);
    a1 = atan_finish(n1 * (r * d0), h1, hx1);
}

// ---- curvature normal component n = up / sqrt(up^2 + uc^2 + eta^2), src/main.cpp:365-368 --------------
// d = E - W (so uc = d/2, uc^2 = d^2/4 exactly)
template <bool STRICT>
__device__ __forceinline__ double normal_component(double up, double d) {
    if (STRICT) {
        const double uc = __dmul_rn(0.5, d);
        const double s = __dadd_rn(__dadd_rn(__dmul_rn(up, up), __dmul_rn(uc, uc)), 1e-8 * 1e-8);
        return __ddiv_rn(up, __dsqrt_rn(s));
    } else {
        double s = fma(up, up, 1e-16);
        s = fma(d * d, 0.25, s);
        // up * rsqrt(s) with the cubic Newton step applied to the product (no 3-register FMA: they cost 3 cycles)
        const double y = rsqrt_seed(s);
        const double t = s * y;
        const double e = fma(-t, y, 1.0);
        const double pe = fma(e, 0.375, 0.5) * e;
        const double n0 = up * y;
        return fma(n0, pe, n0);
    }
}

}  // namespace cvb
