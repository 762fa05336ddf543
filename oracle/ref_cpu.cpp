/*
 * ORACLE / CPU BASELINE -- test and measurement infrastructure, NOT product code.
 *
 * "ref_cpu": a dependency-free C++14/OpenMP port of the reference's hot path that KEEPS the
 * reference's pass structure and thread structure, so that timing it on the GPU box's host cores is a
 * fair stand-in for the reference binary (which cannot be built here: it needs OpenCV 2.4-era C++
 * headers and Boost, neither installed).  bench.py reports it as
 * cpu_baseline.kind = "port".
 *
 * What is kept from /root/reference/src/main.cpp:
 *   - perona_malik (:478-560): channel-parallel OpenMP (num_threads(nof_channels)); per step three
 *     fresh fp64 planes, two Sobel passes, a zero fill, the g loop, the update loop, a copy and a
 *     uint8 conversion;
 *   - CSV loop (:963-1001): channel-parallel OpenMP; region_variance called twice per channel with a
 *     std::function Heaviside (:255-281); variance_penalty as 5 whole-array passes (:299-312);
 *     curvature as 4 filter passes + a 3-thread normalisation loop + 2 filter passes + add (:342-375);
 *     the combine pass (:985); clone + per-pixel std::function delta through a flat-index functor
 *     with i / w, i % w addressing (ParallelPixelFunction.cpp:12-17); multiply; norm; add.
 * What differs: the two data races are removed -- the shared `+=` is done under an
 * ordered section so that the result equals the serial k = 0..N-1 order of oracle/cv_oracle.c, and
 * intensity_avg is zero-initialised.  OpenCV's filter/Sobel inner loops are plain scalar loops here.
 *
 * Compiled with the reference's flags (Makefile:6,16):
 *   -std=c++14 -fopenmp -DNUM_THREADS=3 -O3 -fno-unsafe-math-optimizations -fno-associative-math
 */
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifndef NUM_THREADS
#define NUM_THREADS 3
#endif

namespace {

using Plane = std::vector<double>;
const double kPi = 3.14159265358979323846;

double regularized_heaviside(double x, double eps) { return (1 + 2 / kPi * std::atan(x / eps)) / 2; }
double regularized_delta(double x, double eps) { return eps / (kPi * (std::pow(eps, 2) + std::pow(x, 2))); }

inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
inline int refl101(int p, int n) {
    if (n == 1) return 0;
    while (p < 0 || p >= n) p = p < 0 ? -p : 2 * (n - 1) - p;
    return p;
}

// 1x3 / 3x1 correlation with BORDER_REPLICATE (what cv::filter2D does for ChanVese::Kernel::*)
void filter3(const Plane &src, Plane &dst, int h, int w, const double k[3], bool vertical) {
    Plane out(src.size());
    for (int i = 0; i < h; ++i)
        for (int j = 0; j < w; ++j) {
            double s = 0;
            bool first = true;
            for (int t = 0; t < 3; ++t) {
                if (k[t] == 0) continue;  // OpenCV keeps only the non-zero taps
                const int ii = vertical ? clampi(i + t - 1, 0, h - 1) : i;
                const int jj = vertical ? j : clampi(j + t - 1, 0, w - 1);
                const double v = k[t] * src[(size_t)ii * w + jj];
                s = first ? v : s + v;
                first = false;
            }
            out[(size_t)i * w + j] = s;
        }
    dst.swap(out);
}

// cv::Sobel ksize 3, BORDER_REFLECT_101: separable, row pass then column pass
void sobel(const Plane &I, Plane &d, int h, int w, bool xdir) {
    Plane row(I.size());
    for (int i = 0; i < h; ++i)
        for (int j = 0; j < w; ++j) {
            const double l = I[(size_t)i * w + refl101(j - 1, w)], c = I[(size_t)i * w + j],
                         r = I[(size_t)i * w + refl101(j + 1, w)];
            row[(size_t)i * w + j] = xdir ? r - l : (l + 2 * c) + r;
        }
    for (int i = 0; i < h; ++i) {
        const int it = refl101(i - 1, h), ib = refl101(i + 1, h);
        for (int j = 0; j < w; ++j) {
            const double t = row[(size_t)it * w + j], m = row[(size_t)i * w + j], b = row[(size_t)ib * w + j];
            d[(size_t)i * w + j] = xdir ? 2 * m + (t + b) : b - t;
        }
    }
}

double region_variance(const uint8_t *img, const Plane &u, int h, int w, bool inside,
                       std::function<double(double)> heaviside) {
    double nom = 0.0, denom = 0.0;
    const std::function<double(double)> H =
        inside ? heaviside : std::function<double(double)>([&heaviside](double x) -> double { return 1 - heaviside(x); });
    const double *u_ptr = u.data();
    for (int i = 0; i < h; ++i)
        for (int j = 0; j < w; ++j) {
            const double hv = H(u_ptr[i * w + j]);
            nom += img[i * w + j] * hv;
            denom += hv;
        }
    return nom / denom;
}

Plane variance_penalty(const uint8_t *ch, int h, int w, double c, double lambda) {
    Plane t((size_t)h * w, 0.0);                                           // zeros
    for (size_t p = 0; p < t.size(); ++p) t[p] = (double)ch[p];            // convertTo
    for (size_t p = 0; p < t.size(); ++p) t[p] -= c;                       // -= c
    for (size_t p = 0; p < t.size(); ++p) t[p] = t[p] * t[p];              // pow 2
    for (size_t p = 0; p < t.size(); ++p) t[p] *= lambda;                  // *= lambda
    return t;
}

Plane curvature(const Plane &u, int h, int w) {
    const double eta = 1E-8;
    const double eta2 = std::pow(eta, 2);
    static const double fwd[3] = {0, -1, 1}, bwd[3] = {-1, 1, 0}, ctr[3] = {-0.5, 0, 0.5};
    Plane upx, upy, ucx, ucy;
    filter3(u, upx, h, w, fwd, false);
    filter3(u, upy, h, w, fwd, true);
    filter3(u, ucx, h, w, ctr, false);
    filter3(u, ucy, h, w, ctr, true);
#pragma omp parallel for num_threads(NUM_THREADS)
    for (int i = 0; i < h; ++i)
        for (int j = 0; j < w; ++j) {
            const size_t q = (size_t)i * w + j;
            upx[q] = upx[q] / std::sqrt(std::pow(upx[q], 2) + std::pow(ucx[q], 2) + eta2);
            upy[q] = upy[q] / std::sqrt(std::pow(upy[q], 2) + std::pow(ucy[q], 2) + eta2);
        }
    filter3(upx, upx, h, w, bwd, false);
    filter3(upy, upy, h, w, bwd, true);
    for (size_t q = 0; q < upx.size(); ++q) upx[q] += upy[q];
    return upx;
}

struct PixelFunction {  // ParallelPixelFunction: flat index range, i / w and i % w addressing
    double *data;
    int w;
    std::function<double(double)> func;
    void operator()(long start, long end) const {
        for (long i = start; i != end; ++i) data[(i / w) * w + (i % w)] = func(data[(i / w) * w + (i % w)]);
    }
};

}  // namespace

extern "C" {

struct refcpu_params {
    double mu, nu, dt, eps;
    double lambda1[3];
    double lambda2[3];
};

int refcpu_perona_malik(const uint8_t *const *in, int n, int h, int w, double K, double L, double T,
                        uint8_t *const *out) {
    int steps_done = 0;
#pragma omp parallel for num_threads(n)
    for (int k = 0; k < n; ++k) {
        const size_t np = (size_t)h * w;
        Plane I_prev(np), I_curr(np);
        std::vector<uint8_t> I_res(in[k], in[k] + np);
        for (size_t p = 0; p < np; ++p) I_prev[p] = (double)in[k][p];
        int steps = 0;
        for (double t = 0; t < T; t += L) {
            Plane g(np), dx(np), dy(np);
            sobel(I_prev, dx, h, w, true);
            sobel(I_prev, dy, h, w, false);
            I_curr.assign(np, 0.0);
            for (int i = 0; i < h; ++i)
                for (int j = 0; j < w; ++j) {
                    const double gx = dx[(size_t)i * w + j], gy = dy[(size_t)i * w + j];
                    g[(size_t)i * w + j] = (i == 0 || i == h - 1 || j == 0 || j == w - 1)
                                               ? 1
                                               : std::pow(1.0 + (std::pow(gx, 2) + std::pow(gy, 2)) / (std::pow(K, 2)), -1);
                }
            for (int i = 0; i < h; ++i)
                for (int j = 0; j < w; ++j) {
                    const int in_ = i == h - 1 ? i : i + 1, ip = i == 0 ? i : i - 1;
                    const int jn = j == w - 1 ? j : j + 1, jp = j == 0 ? j : j - 1;
                    const double Is = I_prev[(size_t)in_ * w + j], Ie = I_prev[(size_t)i * w + jn];
                    const double In = I_prev[(size_t)ip * w + j], Iw = I_prev[(size_t)i * w + jp];
                    const double I0 = I_prev[(size_t)i * w + j];
                    const double cs = g[(size_t)in_ * w + j], ce = g[(size_t)i * w + jn];
                    const double cn = g[(size_t)ip * w + j], cw = g[(size_t)i * w + jp];
                    const double c0 = g[(size_t)i * w + j];
                    I_curr[(size_t)i * w + j] = I0 + L * ((cs + c0) * (Is - I0) + (ce + c0) * (Ie - I0) +
                                                          (cn + c0) * (In - I0) + (cw + c0) * (Iw - I0)) / 4;
                }
            I_prev = I_curr;
            for (size_t p = 0; p < np; ++p) {
                const long r = std::lrint(I_prev[p]);
                I_res[p] = (uint8_t)(r < 0 ? 0 : (r > 255 ? 255 : r));
            }
            ++steps;
        }
        std::memcpy(out[k], I_res.data(), np);
        if (k == 0) steps_done = steps;
    }
    return steps_done;
}

double refcpu_stop_condition(const uint8_t *const *ch, int n, int h, int w, double tol) {
    const size_t np = (size_t)h * w;
    Plane avg(np, 0.0);
    for (int k = 0; k < n; ++k) {
        Plane c(np);
        for (size_t p = 0; p < np; ++p) c[p] = (double)ch[k][p];
        for (size_t p = 0; p < np; ++p) avg[p] += c[p];
    }
    const double inv = 1.0 / n;
    for (size_t p = 0; p < np; ++p) avg[p] *= inv;
    double s = 0;
    for (size_t p = 0; p < np; ++p) s += avg[p] * avg[p];
    return tol * std::sqrt(s);
}

int refcpu_csv_run(const uint8_t *const *ch, int n, int h, int w, double *u_io, const refcpu_params *prm,
                   double tol, int max_steps, double *last_norm) {
    const size_t np = (size_t)h * w;
    Plane u(u_io, u_io + np);
    const double eps = prm->eps;
    const auto heaviside = std::bind(regularized_heaviside, std::placeholders::_1, eps);
    const auto delta = std::bind(regularized_delta, std::placeholders::_1, eps);
    const double stop_cond = refcpu_stop_condition(ch, n, h, w, tol);
    int done = 0;
    double norm = NAN;
    for (int t = 1; t <= max_steps; ++t) {
        Plane u_diff(np, 0.0);
#pragma omp parallel for ordered num_threads(n) schedule(static, 1)
        for (int k = 0; k < n; ++k) {
            const double c1 = region_variance(ch[k], u, h, w, true, heaviside);
            const double c2 = region_variance(ch[k], u, h, w, false, heaviside);
            const Plane vin = variance_penalty(ch[k], h, w, c1, prm->lambda1[k]);
            const Plane vout = variance_penalty(ch[k], h, w, c2, prm->lambda2[k]);
            Plane tmp(np);
            for (size_t p = 0; p < np; ++p) tmp[p] = -vin[p] + vout[p];
#pragma omp ordered
            for (size_t p = 0; p < np; ++p) u_diff[p] += tmp[p];
        }
        const Plane kappa = curvature(u, h, w);
        const double alpha = prm->mu * prm->dt, beta = (1.0 / n) * prm->dt, gamma = (-prm->nu) * prm->dt;
        for (size_t p = 0; p < np; ++p) u_diff[p] = kappa[p] * alpha + u_diff[p] * beta + gamma;
        Plane u_cp(u);
        const PixelFunction body{u_cp.data(), w, delta};
        {
            const long total = (long)np;
#pragma omp parallel
            {
                // cv::parallel_for_ splits the flat range into stripes over the pool's threads
                int nt = 1, id = 0;
#ifdef _OPENMP
                nt = omp_get_num_threads();
                id = omp_get_thread_num();
#endif
                const long lo = total * id / nt, hi = total * (id + 1) / nt;
                body(lo, hi);
            }
        }
        for (size_t p = 0; p < np; ++p) u_diff[p] = u_diff[p] * u_cp[p];
        double s = 0;
        for (size_t p = 0; p < np; ++p) s += u_diff[p] * u_diff[p];
        norm = std::sqrt(s);
        for (size_t p = 0; p < np; ++p) u[p] += u_diff[p];
        done = t;
        if (norm <= stop_cond) break;
    }
    std::memcpy(u_io, u.data(), np * sizeof(double));
    if (last_norm) *last_norm = norm;
    return done;
}

}  // extern "C"
