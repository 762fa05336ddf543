/*
 * ORACLE -- test infrastructure, NOT product code.
 *
 * Plain-C restatement of the hot path of ktht/chan_vese (reference at /root/reference, C++14):
 * the Perona-Malik loop and the Chan-Sandberg-Vese time-step loop of src/main.cpp, plus the helpers
 * they call.  Every function cites the reference lines it follows.
 *
 * The reference delegates array arithmetic to OpenCV 2.4.8 (README.md:12; un-vendored, headers absent
 * in this image).  The OpenCV calls on the path are restated here from their published semantics
 * and checked bit-for-bit against the Python binding cv2 4.13 in tests/test_oracle.py:
 *   cv::filter2D (correlation, centre anchor, BORDER_REPLICATE)      src/main.cpp:351-354,371-372
 *   cv::Sobel ksize 3 (BORDER_REFLECT_101; summation order of cv2 4.13: row pass then column pass)
 *                                                                   src/main.cpp:503-504
 *   Mat::convertTo(CV_8U) = saturate_cast<uchar>(cvRound(x)), round-half-even   src/main.cpp:551
 *   MatExpr folding of dt*(mu*kappa - nu + u_diff/N) into one addWeighted       src/main.cpp:985
 *   cv::circle thickness 1 (midpoint circle)                        src/InteractiveDataCirc.cpp:22
 *   cv::threshold(float32(u), 0, 1, THRESH_BINARY)                  src/main.cpp:397-399
 * cv::norm(NORM_L2) is restated as a serial sum (OpenCV's own order is SIMD-dispatch dependent and
 * only decides ties of the stop test).
 *
 * PARITY UNPINNED by the reference: it has no tests, golden vectors or fixtures.
 * The pins are (i) cv2 4.13 run on the same call sites (oracle/cv2_oracle.py, tests/golden/) and
 * (ii) the reference's own source lines cited below.
 *
 * Race-free reading of the reference's data races: intensity_avg is zero-initialised
 * and channels are accumulated serially in k = 0..N-1 order.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off: the reference Makefile sets no -march, so
 * its build has no FMA and forbids reassociation, Makefile:16).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this library.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define CVO_PI 3.14159265358979323846 /* boost::math::constants::pi<double>() */

typedef struct {
    double mu, nu, dt, eps;
    double lambda1[3];
    double lambda2[3];
} cvo_params;

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* src/main.cpp:188-194 */
double cvo_heaviside(double x, double eps) { return (1 + 2 / CVO_PI * atan(x / eps)) / 2; }

/* src/main.cpp:204-210 (std::pow(.,2) is folded to x*x by gcc) */
double cvo_delta(double x, double eps) { return eps / (CVO_PI * (eps * eps + x * x)); }

/* src/main.cpp:221-233: boost::math::sign(sin(pi*i/5) * sin(pi*j/5)) */
void cvo_levelset_checkerboard(int h, int w, double *u) {
    for (int i = 0; i < h; ++i)
        for (int j = 0; j < w; ++j) {
            const double p = sin(CVO_PI * i / 5) * sin(CVO_PI * j / 5);
            u[(size_t)i * w + j] = (p > 0) - (p < 0);
        }
}

/* src/InteractiveDataRect.cpp:20-27: zeros, u(roi) = 1 (roi clipped to the image like mouse_on does) */
void cvo_levelset_rect(int h, int w, int x, int y, int rw, int rh, double *u) {
    memset(u, 0, sizeof(double) * (size_t)h * w);
    for (int i = y; i < y + rh; ++i)
        for (int j = x; j < x + rw; ++j)
            if (i >= 0 && i < h && j >= 0 && j < w) u[(size_t)i * w + j] = 1.0;
}

/* src/InteractiveDataCirc.cpp:18-25: cv::circle(u, P1, radius, 1): thickness 1, LINE_8, shift 0 ->
 * OpenCV's midpoint circle; points outside the image are clipped. */
static void plot(double *u, int h, int w, int x, int y) {
    if (x >= 0 && x < w && y >= 0 && y < h) u[(size_t)y * w + x] = 1.0;
}
void cvo_levelset_circ(int h, int w, int cx, int cy, int radius, double *u) {
    memset(u, 0, sizeof(double) * (size_t)h * w);
    int err = 0, dx = radius, dy = 0, plus = 1, minus = (radius << 1) - 1;
    while (dx >= dy) {
        const int y11 = cy - dy, y12 = cy + dy, y21 = cy - dx, y22 = cy + dx;
        const int x11 = cx - dx, x12 = cx + dx, x21 = cx - dy, x22 = cx + dy;
        plot(u, h, w, x11, y11); plot(u, h, w, x11, y12);
        plot(u, h, w, x12, y11); plot(u, h, w, x12, y12);
        plot(u, h, w, x21, y21); plot(u, h, w, x21, y22);
        plot(u, h, w, x22, y21); plot(u, h, w, x22, y22);
        dy++;
        err += plus;
        plus += 2;
        const int mask = (err <= 0) - 1;
        err -= minus & mask;
        dx += mask;
        minus -= mask & 2;
    }
}

/* src/main.cpp:255-281: c = sum I*g(u) / sum g(u), g = H or 1-H, serial row-major accumulation */
double cvo_region_variance(const uint8_t *img, const double *u, int h, int w, int inside, double eps) {
    double nom = 0.0, denom = 0.0;
    for (int i = 0; i < h; ++i)
        for (int j = 0; j < w; ++j) {
            double hv = cvo_heaviside(u[(size_t)i * w + j], eps);
            if (!inside) hv = 1 - hv;
            nom += img[(size_t)i * w + j] * hv;
            denom += hv;
        }
    return nom / denom;
}

/* src/main.cpp:299-312: lambda * (double(I) - c)^2 */
void cvo_variance_penalty(const uint8_t *ch, int h, int w, double c, double lambda, double *out) {
    for (size_t p = 0; p < (size_t)h * w; ++p) {
        double t = (double)ch[p];
        t -= c;
        t = t * t;
        t *= lambda;
        out[p] = t;
    }
}

/* src/main.cpp:342-375.  filter2D with BORDER_REPLICATE on u and again on the normalised fields. */
void cvo_curvature(const double *u, int h, int w, double *kappa) {
    const double eta = 1E-8;
    const double eta2 = eta * eta;
    double *nx = (double *)malloc(sizeof(double) * (size_t)h * w);
    double *ny = (double *)malloc(sizeof(double) * (size_t)h * w);
    for (int i = 0; i < h; ++i)
        for (int j = 0; j < w; ++j) {
            const int jn = clampi(j + 1, 0, w - 1), jp = clampi(j - 1, 0, w - 1);
            const int in = clampi(i + 1, 0, h - 1), ip = clampi(i - 1, 0, h - 1);
            const double u0 = u[(size_t)i * w + j];
            const double upx = -u0 + u[(size_t)i * w + jn];                           /* fwd_x :351 */
            const double upy = -u0 + u[(size_t)in * w + j];                           /* fwd_y :352 */
            const double ucx = -0.5 * u[(size_t)i * w + jp] + 0.5 * u[(size_t)i * w + jn]; /* :353 */
            const double ucy = -0.5 * u[(size_t)ip * w + j] + 0.5 * u[(size_t)in * w + j]; /* :354 */
            nx[(size_t)i * w + j] = upx / sqrt(upx * upx + ucx * ucx + eta2);         /* :365-366 */
            ny[(size_t)i * w + j] = upy / sqrt(upy * upy + ucy * ucy + eta2);         /* :367-368 */
        }
    for (int i = 0; i < h; ++i)
        for (int j = 0; j < w; ++j) {
            const int jp = clampi(j - 1, 0, w - 1), ip = clampi(i - 1, 0, h - 1);
            const double kx = -nx[(size_t)i * w + jp] + nx[(size_t)i * w + j];        /* bwd_x :371 */
            const double ky = -ny[(size_t)ip * w + j] + ny[(size_t)i * w + j];        /* bwd_y :372 */
            kappa[(size_t)i * w + j] = kx + ky;                                       /* :373 */
        }
    free(nx);
    free(ny);
}

/* cv::Sobel(I, d, CV_64F, {1,0}|{0,1}, 3), default BORDER_REFLECT_101; summation order of cv2 4.13:
 * row pass first ([-1 0 1] -> R-L ; [1 2 1] -> (L + 2C) + R), column pass second
 * ([1 2 1] -> 2*m + (t + b) ; [-1 0 1] -> b - t).  src/main.cpp:503-504 */
static inline int refl101(int p, int n) {
    if (n == 1) return 0;
    while (p < 0 || p >= n) p = p < 0 ? -p : 2 * (n - 1) - p;
    return p;
}
void cvo_sobel(const double *I, int h, int w, double *dx, double *dy) {
    double *rd = (double *)malloc(sizeof(double) * (size_t)h * w); /* row difference  */
    double *rs = (double *)malloc(sizeof(double) * (size_t)h * w); /* row smoothing   */
    for (int i = 0; i < h; ++i)
        for (int j = 0; j < w; ++j) {
            const double l = I[(size_t)i * w + refl101(j - 1, w)];
            const double c = I[(size_t)i * w + j];
            const double r = I[(size_t)i * w + refl101(j + 1, w)];
            rd[(size_t)i * w + j] = r - l;
            rs[(size_t)i * w + j] = (l + 2 * c) + r;
        }
    for (int i = 0; i < h; ++i) {
        const int it = refl101(i - 1, h), ib = refl101(i + 1, h);
        for (int j = 0; j < w; ++j) {
            dx[(size_t)i * w + j] = 2 * rd[(size_t)i * w + j] + (rd[(size_t)it * w + j] + rd[(size_t)ib * w + j]);
            dy[(size_t)i * w + j] = rs[(size_t)ib * w + j] - rs[(size_t)it * w + j];
        }
    }
    free(rd);
    free(rs);
}

/* src/main.cpp:498: for (double t = 0; t < T; t += L) -- the step count is decided by fp accumulation */
int cvo_pm_num_steps(double L, double T) {
    int n = 0;
    for (double t = 0; t < T; t += L) ++n;
    return n;
}

/* One Perona-Malik step on one fp64 plane, src/main.cpp:500-548 */
void cvo_pm_step(const double *Ip, int h, int w, double K, double L, double *Ic) {
    double *g = (double *)malloc(sizeof(double) * (size_t)h * w);
    double *dx = (double *)malloc(sizeof(double) * (size_t)h * w);
    double *dy = (double *)malloc(sizeof(double) * (size_t)h * w);
    cvo_sobel(Ip, h, w, dx, dy);
    for (int i = 0; i < h; ++i)
        for (int j = 0; j < w; ++j) { /* :513-522; pow(x,-1) is folded to 1/x by gcc */
            const double gx = dx[(size_t)i * w + j], gy = dy[(size_t)i * w + j];
            g[(size_t)i * w + j] = (i == 0 || i == h - 1 || j == 0 || j == w - 1)
                                       ? 1
                                       : 1.0 / (1.0 + (gx * gx + gy * gy) / (K * K));
        }
    for (int i = 0; i < h; ++i)
        for (int j = 0; j < w; ++j) { /* :524-548 */
            const int in = i == h - 1 ? i : i + 1, ip = i == 0 ? i : i - 1;
            const int jn = j == w - 1 ? j : j + 1, jp = j == 0 ? j : j - 1;
            const double Is = Ip[(size_t)in * w + j], Ie = Ip[(size_t)i * w + jn];
            const double In = Ip[(size_t)ip * w + j], Iw = Ip[(size_t)i * w + jp];
            const double I0 = Ip[(size_t)i * w + j];
            const double cs = g[(size_t)in * w + j], ce = g[(size_t)i * w + jn];
            const double cn = g[(size_t)ip * w + j], cw = g[(size_t)i * w + jp];
            const double c0 = g[(size_t)i * w + j];
            Ic[(size_t)i * w + j] = I0 + L * ((cs + c0) * (Is - I0) + (ce + c0) * (Ie - I0) +
                                              (cn + c0) * (In - I0) + (cw + c0) * (Iw - I0)) / 4;
        }
    free(g);
    free(dx);
    free(dy);
}

/* saturate_cast<uchar>(double): cvRound (lrint, round-half-even) then clamp, src/main.cpp:551 */
uint8_t cvo_saturate_u8(double v) {
    const long r = lrint(v);
    return (uint8_t)(r < 0 ? 0 : (r > 255 ? 255 : r));
}

/* nsteps PM steps on one fp64 plane in place (the fp64 state the reference keeps across steps) */
void cvo_pm_evolve(double *I, int h, int w, double K, double L, int nsteps) {
    double *tmp = (double *)malloc(sizeof(double) * (size_t)h * w);
    for (int s = 0; s < nsteps; ++s) {
        cvo_pm_step(I, h, w, K, L, tmp);
        memcpy(I, tmp, sizeof(double) * (size_t)h * w); /* :550 */
    }
    free(tmp);
}

/* src/main.cpp:478-560: N planar uint8 channels in, N planar uint8 channels out; returns steps run.
 * With zero steps the reference returns an unset I_res; this restatement returns the input. */
int cvo_perona_malik(const uint8_t *const *in, int n, int h, int w, double K, double L, double T,
                     uint8_t *const *out) {
    const int nsteps = cvo_pm_num_steps(L, T);
    double *I = (double *)malloc(sizeof(double) * (size_t)h * w);
    for (int k = 0; k < n; ++k) {
        for (size_t p = 0; p < (size_t)h * w; ++p) I[p] = (double)in[k][p]; /* :495-496 */
        cvo_pm_evolve(I, h, w, K, L, nsteps);
        for (size_t p = 0; p < (size_t)h * w; ++p) out[k][p] = nsteps ? cvo_saturate_u8(I[p]) : in[k][p];
    }
    free(I);
    return nsteps;
}

/* src/main.cpp:949-960: tol * || (sum_k I_k) * (1/N) ||_2 (Mat /= s multiplies by 1/s) */
double cvo_stop_condition(const uint8_t *const *ch, int n, int h, int w, double tol) {
    const double inv = 1.0 / n;
    double s = 0.0;
    for (size_t p = 0; p < (size_t)h * w; ++p) {
        double a = 0.0;
        for (int k = 0; k < n; ++k) a += (double)ch[k][p];
        a *= inv;
        s += a * a;
    }
    return tol * sqrt(s);
}

/* One time step, src/main.cpp:965-994.  u is updated in place; returns ||du||_2.
 * c1/c2 (length n) receive the region means used by this step when non-NULL.
 * If c1_in/c2_in are non-NULL they are used instead of computing the means (test hook). */
double cvo_csv_step_ex(const uint8_t *const *ch, int n, int h, int w, double *u, const cvo_params *p,
                       const double *c1_in, const double *c2_in, double *c1_out, double *c2_out) {
    const size_t np = (size_t)h * w;
    double *u_diff = (double *)calloc(np, sizeof(double)); /* :965 */
    double *vi = (double *)malloc(sizeof(double) * np);
    double *vo = (double *)malloc(sizeof(double) * np);
    double *kappa = (double *)malloc(sizeof(double) * np);
    for (int k = 0; k < n; ++k) { /* :968-980, serial */
        const double c1 = c1_in ? c1_in[k] : cvo_region_variance(ch[k], u, h, w, 1, p->eps); /* :973 */
        const double c2 = c2_in ? c2_in[k] : cvo_region_variance(ch[k], u, h, w, 0, p->eps); /* :974 */
        if (c1_out) c1_out[k] = c1;
        if (c2_out) c2_out[k] = c2;
        cvo_variance_penalty(ch[k], h, w, c1, p->lambda1[k], vi); /* :977 */
        cvo_variance_penalty(ch[k], h, w, c2, p->lambda2[k], vo); /* :978 */
        for (size_t q = 0; q < np; ++q) u_diff[q] = u_diff[q] + (-vi[q] + vo[q]); /* :979 */
    }
    cvo_curvature(u, h, w, kappa); /* :982 */
    /* :985 as OpenCV's MatExpr executes it: addWeighted(kappa, mu*dt, u_diff, (1/N)*dt, -nu*dt) */
    const double alpha = p->mu * p->dt, beta = (1.0 / n) * p->dt, gamma = (-p->nu) * p->dt;
    double s = 0.0;
    for (size_t q = 0; q < np; ++q) {
        double d = kappa[q] * alpha + u_diff[q] * beta + gamma;
        d = d * cvo_delta(u[q], p->eps); /* :988-992 */
        s += d * d;                      /* :993 */
        u_diff[q] = d;
    }
    for (size_t q = 0; q < np; ++q) u[q] += u_diff[q]; /* :994 */
    free(u_diff);
    free(vi);
    free(vo);
    free(kappa);
    return sqrt(s);
}

double cvo_csv_step(const uint8_t *const *ch, int n, int h, int w, double *u, const cvo_params *p,
                    double *c1_out, double *c2_out) {
    return cvo_csv_step_ex(ch, n, h, w, u, p, NULL, NULL, c1_out, c2_out);
}

/* src/main.cpp:949-1001: returns steps done (the breaking step's update IS applied, :994,:1000) */
int cvo_csv_run(const uint8_t *const *ch, int n, int h, int w, double *u, const cvo_params *p, double tol,
                int max_steps, double *last_norm) {
    const double stop = cvo_stop_condition(ch, n, h, w, tol);
    int t, done = 0;
    double norm = NAN;
    for (t = 1; t <= max_steps; ++t) {
        norm = cvo_csv_step(ch, n, h, w, u, p, NULL, NULL);
        done = t;
        if (norm <= stop) break;
    }
    if (last_norm) *last_norm = norm;
    return done;
}

/* src/main.cpp:395-400: float32(u) > 0 -> 1, optionally inverted */
void cvo_mask(const double *u, int h, int w, int invert, uint8_t *mask) {
    for (size_t q = 0; q < (size_t)h * w; ++q) {
        const uint8_t m = ((float)u[q] > 0.0f) ? 1 : 0;
        mask[q] = invert ? (uint8_t)(1 - m) : m;
    }
}

/* ParallelPixelFunction with f = regularized_delta, src/ParallelPixelFunction.cpp:12-17 + main.cpp:989 */
void cvo_delta_map(double *data, size_t n, double eps) {
    for (size_t q = 0; q < n; ++q) data[q] = cvo_delta(data[q], eps);
}
