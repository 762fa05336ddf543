// C ABI of chan_vese_b200 (include/chan_vese_b200.h): contexts, resident jobs (sessions / batches),
// one-shot calls on host buffers, the multi-GPU row-slab plumbing (NCCL, loaded at run time).
// All compute is in csv_kernels.cu / pm_kernels.cu; there is no CPU fallback anywhere in this file.
#include <dlfcn.h>
#include <limits.h>
#include <math.h>
#include <stdarg.h>
#include <stddef.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>  // header-only NVTX v3: ranges cost nothing unless a profiler is attached

#include "../../include/chan_vese_b200.h"
#include "common.cuh"
#include "kernels.h"

// NVTX range around a solver phase (tracing): cvb.pm / cvb.csv / cvb.upload+pm show up on the
// timeline of nsys / ncu --nvtx next to the kernels they launch.
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

using namespace cvb;

// ---- NCCL, bound at run time so that the library loads (and the host helpers work) without it -----------
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclUint8 = 1, ncclFloat64 = 8 };
struct NcclApi {
    void *handle = nullptr;
    int (*GetUniqueId)(ncclUniqueId *) = nullptr;
    int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;
static bool load_nccl(std::string &err) {
    if (g_nccl.handle) return true;
    const char *names[] = {getenv("CVB_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void *h = nullptr;
    for (const char *n : names) {
        if (!n || !*n) continue;
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) {
        err = std::string("cannot load NCCL: ") + dlerror();
        return false;
    }
#define CVB_SYM(field, name)                                         \
    *(void **)(&g_nccl.field) = dlsym(h, name);                      \
    if (!g_nccl.field) {                                             \
        err = std::string("NCCL symbol missing: ") + name;           \
        return false;                                                \
    }
    CVB_SYM(GetUniqueId, "ncclGetUniqueId")
    CVB_SYM(CommInitRank, "ncclCommInitRank")
    CVB_SYM(CommDestroy, "ncclCommDestroy")
    CVB_SYM(AllGather, "ncclAllGather")
    CVB_SYM(Send, "ncclSend")
    CVB_SYM(Recv, "ncclRecv")
    CVB_SYM(GroupStart, "ncclGroupStart")
    CVB_SYM(GroupEnd, "ncclGroupEnd")
    CVB_SYM(GetErrorString, "ncclGetErrorString")
#undef CVB_SYM
    g_nccl.handle = h;
    return true;
}

// ---- objects ----------------------------------------------------------------------------------------------
struct cvb_context {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    std::string err;
    cvb_math_mode math = CVB_MATH_FAST;
    int tile_rows = 0;
    cvb_stats stats{};
    double *d_atan_tab = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaStream_t copy_stream = nullptr;  // host-to-device copies overlapped with compute (upload_image_smooth)
    cudaEvent_t copy_ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    // multi-GPU
    ncclComm_t comm = nullptr;
    int nranks = 1, rank = 0;
    // the session behind the one-shot calls: kept between calls of the same shape, so that a caller looping over
    // cvb_segment / cvb_csv_run / ... does not pay a cudaMalloc + cudaFree of every plane per call (cvb_context_trim frees it)
    struct cvb_session *oneshot = nullptr;
};

struct Job {
    cvb_context *ctx = nullptr;
    Geom g{};
    cvb_precision prec = CVB_PRECISION_F64;
    uint8_t *d_img = nullptr;
    uint8_t *d_img_saved = nullptr;  // cvb_*_save_image
    double *d_u[2] = {nullptr, nullptr};
    double *d_pm[2] = {nullptr, nullptr};
    double *d_aux = nullptr;  // one fp64 plane set (curvature output), lazy
    CsvState *d_state = nullptr;
    CsvState *h_state = nullptr;  // pinned, 2 * count
    double *d_partials = nullptr;
    double *d_seg_sums = nullptr;          // one sum per row segment (reduce.cuh)
    unsigned int *d_seg_ticket = nullptr;
    int seg_level = 0;
    double *d_group = nullptr;
    signed char *d_sign = nullptr;  // checkerboard sign vectors
    uint8_t *d_bits = nullptr;      // packed masks of a whole batch (cvb_batch_masks_packed), lazy
    int ngroups_local = 0;
    int group_lo = 0, group_hi = NGROUPS;  // groups owned by this rank
    bool slab = false;
    // P2P row-slab run: peers' buffers mapped with CUDA IPC (see comm.cuh / reduce.cuh)
    bool p2p = false;
    CommBox *d_box = nullptr;
    CommView cv{};
    std::vector<void *> ipc_opened;
#ifdef CSV_TMA
    CUtensorMap tm_u[2], tm_img;
#endif
    unsigned int pm_seq = 0;
    cudaEvent_t ev_prefetched = nullptr, ev_restored = nullptr;  // prefetch_image / restore_image hand-over
    bool prefetch_pending = false, prefetch_needs_halo = false;
    unsigned long long seen_wait_ns = 0;  // CommBox::wait_ns / wait_count already added to the context's stats
    unsigned int seen_wait_count = 0;
    int pm_cur = -1;  // PM state buffer that holds the input of the last (quantising) step of the newest run; -1: none
};
static inline size_t esz(const Job *j) { return j->prec == CVB_PRECISION_F32 ? sizeof(float) : sizeof(double); }
static inline bool is_f32(const Job *j) { return j->prec == CVB_PRECISION_F32; }
// plane `m` of level-set buffer `b` (element size depends on the precision)
static inline char *u_plane(const Job *j, int b, size_t m) {
    return reinterpret_cast<char *>(j->d_u[b]) + m * (size_t)j->g.plane_elems * esz(j);
}
struct cvb_session : Job {};
struct cvb_batch : Job {};

static thread_local std::string g_create_err;

static cvb_status fail(cvb_context *ctx, cvb_status st, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx)
        ctx->err = buf;
    else
        g_create_err = buf;
    return st;
}
#define CU(ctx, expr)                                                                                    \
    do {                                                                                                 \
        cudaError_t e__ = (expr);                                                                        \
        if (e__ != cudaSuccess)                                                                          \
            return fail(ctx, e__ == cudaErrorMemoryAllocation ? CVB_ERR_OUT_OF_MEMORY : CVB_ERR_CUDA,    \
                        "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__);    \
    } while (0)
#define NC(ctx, expr)                                                                                    \
    do {                                                                                                 \
        int r__ = (expr);                                                                                \
        if (r__ != 0)                                                                                    \
            return fail(ctx, CVB_ERR_COMM, "%s failed: %s", #expr, g_nccl.GetErrorString(r__));         \
    } while (0)
#define TRY(expr)                          \
    do {                                   \
        cvb_status s__ = (expr);           \
        if (s__ != CVB_OK) return s__;     \
    } while (0)

// ---- host-side helpers (no GPU) ---------------------------------------------------------------------------
extern "C" const char *cvb_version(void) { return "chan_vese_b200 0.1 (sm_100a)"; }

extern "C" int cvb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

// src/main.cpp:498: for (double t = 0; t < T; t += L)
extern "C" int cvb_pm_num_steps(double L, double T) {
    if (!(L > 0.0)) return -1;
    int n = 0;
    for (double t = 0; t < T; t += L) {
        if (n == INT_MAX) return -1;
        ++n;
    }
    return n;
}

static const double kPi = 3.14159265358979323846;  // boost::math::constants::pi<double>()
static signed char sign_of(double p) { return (signed char)((p > 0) - (p < 0)); }

// src/main.cpp:221-233: sign(sin(pi*i/5) * sin(pi*j/5)), glibc sin on the host
extern "C" cvb_status cvb_levelset_checkerboard(int h, int w, double *u) {
    if (h <= 0 || w <= 0 || !u) return CVB_ERR_INVALID_ARGUMENT;
    std::vector<double> sj(w);
    for (int j = 0; j < w; ++j) sj[j] = sin(kPi * j / 5);
    for (int i = 0; i < h; ++i) {
        const double si = sin(kPi * i / 5);
        for (int j = 0; j < w; ++j) u[(size_t)i * w + j] = (double)sign_of(si * sj[j]);
    }
    return CVB_OK;
}
// InteractiveDataRect::get_levelset, src/InteractiveDataRect.cpp:20-27
extern "C" cvb_status cvb_levelset_rect(int h, int w, int x, int y, int rw, int rh, double *u) {
    if (h <= 0 || w <= 0 || !u) return CVB_ERR_INVALID_ARGUMENT;
    memset(u, 0, sizeof(double) * (size_t)h * w);
    for (int i = (y < 0 ? 0 : y); i < y + rh && i < h; ++i)
        for (int j = (x < 0 ? 0 : x); j < x + rw && j < w; ++j) u[(size_t)i * w + j] = 1.0;
    return CVB_OK;
}
// InteractiveDataCirc::get_levelset, src/InteractiveDataCirc.cpp:18-25: cv::circle(u, c, r, 1) with the default
// thickness 1 = the 8-way symmetric midpoint circle, clipped to the image.
extern "C" cvb_status cvb_levelset_circ(int h, int w, int cx, int cy, int radius, double *u) {
    if (h <= 0 || w <= 0 || !u || radius < 0) return CVB_ERR_INVALID_ARGUMENT;
    memset(u, 0, sizeof(double) * (size_t)h * w);
    auto put = [&](int px, int py) {
        if (px >= 0 && px < w && py >= 0 && py < h) u[(size_t)py * w + px] = 1.0;
    };
    int dx = radius, dy = 0, err = 0, inc = 1, dec = 2 * radius - 1;
    while (dx >= dy) {
        put(cx - dx, cy - dy); put(cx - dx, cy + dy); put(cx + dx, cy - dy); put(cx + dx, cy + dy);
        put(cx - dy, cy - dx); put(cx - dy, cy + dx); put(cx + dy, cy - dx); put(cx + dy, cy + dx);
        ++dy;
        err += inc;
        inc += 2;
        if (err > 0) {
            err -= dec;
            --dx;
            dec -= 2;
        }
    }
    return CVB_OK;
}

// Rows per segment.  A CTA is one warp working down one segment of one 64-column strip; the grid is
// planes x segments x strips CTAs over SLOTS = SMs x resident CTAs per SM slots.  Long segments amortise the row
// priming (3-4 extra rows and one pipeline fill per CTA), but the LAST wave must be full too: among the candidates
// pick the one whose wave count wastes the least ("wave quantisation"; at 16384^2 on 8 GPUs 128-row segments give
// 1.79 waves = 10 % idle, 79-row segments 2.91 waves = 3 %).  Small jobs keep >= 2 waves with at least 4 rows.
static const int kSlots = 148 * 16;
// min_tail > 0: lengths that leave a last segment of fewer than min_tail rows are not considered.
// slots: resident CTAs of the whole GPU; prime: rows' worth of work a CTA spends before its first output row.
static int wave_aware_rows(int rows_per_rank, long long ctas_per_seg, int min_tail = 0, int slots = kSlots, double prime = 3.5) {
    int best = 0;
    double best_eff = -1.0;
    for (int r = 192; r >= 48; --r) {
        const int tail = rows_per_rank % r;
        if (tail != 0 && tail < min_tail) continue;
        const long long ctas = (long long)ceil_div(rows_per_rank, r) * ctas_per_seg;
        const double waves = (double)ctas / slots;
        if (waves < 2.0) continue;
        const double eff = waves / ceil(waves) * (1.0 - prime / (r + prime));  // tail loss x priming overhead
        if (eff > best_eff + 1e-9) {
            best_eff = eff;
            best = r;
        }
    }
    return best;
}
// Modelled efficiency of tile length r for an image cut into nranks row slabs (cvb_slab_partition): slabs are unions of
// whole segments, the rank with the most rows sets the pace (every step ends in an all-rank reduction), and every CTA pays
// a few rows' worth of priming.  0 when a slab would have fewer than 1.9 waves of CTAs.
// Measured on B200 (profiles/README.md, r2g / r2k): at 16384^2 the csv_step time follows the rows per rank, NOT
// ceil(waves) x tile length -- tiles of 40, 64 and 123 rows on 8 GPUs and 123 / 124 / 126 / 128 rows on one GPU are within
// 1 % of each other -- so the model has no wave-quantisation term.
static double slab_tile_eff(int h, int ncb, int r, int nranks) {
    const int nseg = ceil_div(h, r);
    int most_rows = 0, most_segs = 0;
    for (int k = 0; k < nranks; ++k) {
        const int s0 = group_seg_begin(k * (NGROUPS / nranks), nseg), s1 = group_seg_begin((k + 1) * (NGROUPS / nranks), nseg);
        most_rows = std::max(most_rows, std::min(s1 * r, h) - std::min(s0 * r, h));
        most_segs = std::max(most_segs, s1 - s0);
    }
    if ((double)most_segs * ncb / kSlots < 1.9 || most_rows <= 0) return 0.0;  // (nothing below 1.9 waves has been measured)
    return ((double)h / nranks) / most_rows * (r / (r + 3.5));
}
// Rows per tile of a single image.  The tiling fixes the order of the fused sums (reduce.cuh), so it must NOT depend on
// the number of GPUs the image is spread over: a run on 1, 2, 4 or 8 GPUs then gives bit-identical results with the
// automatic choice too.  The choice maximises the single-GPU efficiency plus the worst efficiency among the slab
// decompositions (2, 4, 8 ranks) that are large enough to be worth running.
static int auto_seg_rows(int h, int w, int count, int /*nranks: deliberately unused*/) {
    const int ncb = ceil_div(w, CSV_CB);
    if (count == 1) {
        int best = 0;
        double best_score = 0.0;
        for (int r = 192; r >= 24; --r) {
            const double e1 = slab_tile_eff(h, ncb, r, 1);
            if (e1 <= 0.0) continue;
            double worst = e1;
            bool ok = true;
            for (int n : {2, 4, 8}) {
                if ((double)h / n * ncb / 48.0 < 2.0 * kSlots) continue;  // not worth a slab run at any tile length
                const double e = slab_tile_eff(h, ncb, r, n);
                if (e <= 0.0) ok = false;
                worst = std::min(worst, e);
            }
            if (!ok) continue;
            const double score = e1 + worst;
            if (score > best_score + 1e-9) {
                best_score = score;
                best = r;
            }
        }
        if (best > 0) return best;
    }
    const int r = wave_aware_rows(h, (long long)count * ncb);
    if (r > 0) return r;
    const int cands[] = {32, 16, 8, 4};
    for (int s : cands)
        if ((long long)count * ceil_div(h, s) * ncb >= 2LL * kSlots) return s;
    return 4;
}
// Segments per CTA of the production CSV kernel (Geom::seg_mult).  The segment length is the same for every GPU count (it
// fixes the order of the sums); how many of them a CTA should march through is not: a job of many waves wants long row
// loops (priming amortised), a slab of two or three waves wants short CTAs -- the last, partly filled wave and the drain
// of a launch cost about one CTA's duration.  Model: CTAs of L = m * seg_rows rows (+ priming) run in ceil(waves)
// generations, the last one weighted by how full it is.  CVB_SEG_MULT overrides.
static int auto_seg_mult(int seg_rows, int nseg, int ncb, int count) {
    if (const char *e = getenv("CVB_SEG_MULT"))
        if (atoi(e) >= 1) return (seg_rows % 4 == 0 || atoi(e) == 1) ? atoi(e) : 1;
    if (seg_rows % 4 != 0) return 1;  // the 4x unrolled row loop delivers a segment's sums between two groups of four rows
    // Measured (profiles/README.md, r2g): at 16384^2 on 8 GPUs tiles of 40, 64 and 123 rows give the same csv_step time
    // within 1 % -- the shorter tail of short CTAs and their extra priming cancel -- so the automatic tile length already
    // serves every GPU count and one segment per CTA stays the default; the model below is used for explicit small tiles.
    if (seg_rows >= 48) return 1;
    int best = 1;
    double best_t = 1e300;
    for (int m = 1; m <= 8 && m * seg_rows <= 256; ++m) {
        const double ctas = (double)count * ceil_div(nseg, m) * ncb;
        const double waves = ctas / kSlots;
        const double len = (double)m * seg_rows + 3.5;
        const double full = floor(waves), part = waves - full;
        // a partly filled last generation runs faster than a full one (fewer warps share an SM), but not in proportion
        const double t = (full + (part > 0 ? 0.35 + 0.65 * part : 0.0)) * len + 0.25 * len;
        if (t < best_t * 0.995) {
            best_t = t;
            best = m;
        }
    }
    return best;
}

// min_tail = HALO for the row slabs of a multi-rank run: the slab's last HALO rows should lie in ONE segment (they are
// pushed to the neighbour by the CTAs of that segment; pm_push_boundary copes with a one-row last segment, but there is
// no need to make one).  0 otherwise: whole images keep the plain choice.
static int auto_pm_seg_rows(int rows, int w, int planes, int min_tail, int strip = PM_CB, int slots = kSlots, double prime = 3.5) {
    const int ncb = ceil_div(w, strip);
    auto tail_ok = [&](int s) { return rows % s == 0 || rows % s >= min_tail; };
    const int r = wave_aware_rows(rows, (long long)planes * ncb, min_tail, slots, prime);
    if (r > 0) return r;
    const int cands[] = {32, 16, 8, 4};
    for (int s : cands)
        if ((long long)planes * ceil_div(rows, s) * ncb >= 2LL * slots && tail_ok(s)) return s;
    for (int s : {4, 5, 6, 7})
        if (tail_ok(s)) return s;
    return 4;
}
extern "C" int cvb_auto_tile_rows(int h, int w, int count, int nranks) {
    if (h <= 0 || w <= 0 || count <= 0 || nranks <= 0) return 0;
    return auto_seg_rows(h, w, count, nranks);
}
extern "C" cvb_status cvb_slab_partition(int h, int tile_rows, int nranks, int rank, int *row_lo, int *row_hi) {
    if (h <= 0 || tile_rows <= 0 || nranks <= 0 || rank < 0 || rank >= nranks || NGROUPS % nranks != 0 || !row_lo ||
        !row_hi)
        return CVB_ERR_INVALID_ARGUMENT;
    const int nseg = ceil_div(h, tile_rows);
    const int g0 = rank * (NGROUPS / nranks), g1 = (rank + 1) * (NGROUPS / nranks);
    const int s0 = group_seg_begin(g0, nseg), s1 = group_seg_begin(g1, nseg);
    *row_lo = std::min(s0 * tile_rows, h);
    *row_hi = std::min(s1 * tile_rows, h);
    return (*row_hi > *row_lo) ? CVB_OK : CVB_ERR_INVALID_ARGUMENT;
}

// ---- context ------------------------------------------------------------------------------------------------
extern "C" cvb_status cvb_context_create(int device, void *stream, cvb_context **out) {
    if (!out) return fail(nullptr, CVB_ERR_INVALID_ARGUMENT, "out is NULL");
    *out = nullptr;
    int n = cvb_device_count();
    if (n <= 0) return fail(nullptr, CVB_ERR_NO_DEVICE, "no CUDA device available (this library has no CPU path)");
    if (device < 0 || device >= n) return fail(nullptr, CVB_ERR_INVALID_ARGUMENT, "device %d out of range [0,%d)", device, n);
    CU(nullptr, cudaSetDevice(device));
    cvb_context *c = new cvb_context;
    c->device = device;
    if (stream) {
        c->stream = (cudaStream_t)stream;
    } else {
        cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) {
            delete c;
            return fail(nullptr, CVB_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e));
        }
        c->own_stream = true;
    }
    for (auto &e : c->ev) cudaEventCreate(&e);
    // atan(c_q)/pi table of math.cuh, from the host libm: quarter-octave centres from 2^-4 up to 2^44
    double tab[192];
    for (int q = 0; q < 192; ++q) {
        const uint64_t bits = ((uint64_t)(((1019u * 4u + (unsigned)q) << 18) | 0x00020000u)) << 32;
        double cq;
        memcpy(&cq, &bits, 8);
        tab[q] = atan(cq) / kPi;
    }
    if (cudaMalloc(&c->d_atan_tab, sizeof tab) != cudaSuccess ||
        cudaMemcpy(c->d_atan_tab, tab, sizeof tab, cudaMemcpyHostToDevice) != cudaSuccess) {
        cvb_status st = fail(nullptr, CVB_ERR_CUDA, "context set-up: %s", cudaGetErrorString(cudaGetLastError()));
        delete c;
        return st;
    }
    *out = c;
    return CVB_OK;
}
extern "C" void cvb_session_destroy(cvb_session *s);
extern "C" cvb_status cvb_context_trim(cvb_context *c) {
    if (!c) return CVB_ERR_INVALID_ARGUMENT;
    cvb_session_destroy(c->oneshot);
    c->oneshot = nullptr;
    return CVB_OK;
}
extern "C" void cvb_context_destroy(cvb_context *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cvb_context_trim(c);
    if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
    cudaFree(c->d_atan_tab);
    for (auto &e : c->ev)
        if (e) cudaEventDestroy(e);
    for (auto &e : c->copy_ev)
        if (e) cudaEventDestroy(e);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->own_stream) cudaStreamDestroy(c->stream);
    delete c;
}
extern "C" const char *cvb_last_error(const cvb_context *c) { return c ? c->err.c_str() : g_create_err.c_str(); }
extern "C" cvb_status cvb_context_trim(cvb_context *c);
extern "C" cvb_status cvb_context_set_math_mode(cvb_context *c, cvb_math_mode m) {
    if (!c || (m != CVB_MATH_FAST && m != CVB_MATH_STRICT)) return CVB_ERR_INVALID_ARGUMENT;
    c->math = m;
    return CVB_OK;
}
extern "C" cvb_status cvb_context_set_tile_rows(cvb_context *c, int rows) {
    if (!c || rows < 0 || rows > 4096) return CVB_ERR_INVALID_ARGUMENT;
    if (rows != c->tile_rows && c->oneshot) cvb_context_trim(c);  // the cached one-shot session was tiled differently
    c->tile_rows = rows;
    return CVB_OK;
}
extern "C" cvb_status cvb_context_get_stats(cvb_context *c, cvb_stats *out) {
    if (!c || !out) return CVB_ERR_INVALID_ARGUMENT;
    *out = c->stats;
    return CVB_OK;
}
extern "C" cvb_status cvb_context_reset_stats(cvb_context *c) {
    if (!c) return CVB_ERR_INVALID_ARGUMENT;
    c->stats = cvb_stats{};
    return CVB_OK;
}
extern "C" cvb_status cvb_context_synchronize(cvb_context *c) {
    if (!c) return CVB_ERR_INVALID_ARGUMENT;
    CU(c, cudaStreamSynchronize(c->stream));
    return CVB_OK;
}
extern "C" cvb_status cvb_host_alloc(size_t bytes, void **out) {
    if (!out) return CVB_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    if (cudaMallocHost(out, bytes ? bytes : 1) != cudaSuccess) {
        cudaGetLastError();
        return CVB_ERR_OUT_OF_MEMORY;
    }
    return CVB_OK;
}
extern "C" void cvb_host_free(void *p) {
    if (p) cudaFreeHost(p);
}

// ---- multi-GPU plumbing ---------------------------------------------------------------------------------------
extern "C" cvb_status cvb_comm_create_id(cvb_context *c, void *id_out) {
    if (!c || !id_out) return CVB_ERR_INVALID_ARGUMENT;
    if (!load_nccl(c->err)) return CVB_ERR_COMM;
    ncclUniqueId id;
    NC(c, g_nccl.GetUniqueId(&id));
    memcpy(id_out, &id, sizeof id);
    return CVB_OK;
}
extern "C" cvb_status cvb_comm_init(cvb_context *c, const void *id, int nranks, int rank) {
    if (!c || !id || nranks < 1 || rank < 0 || rank >= nranks || NGROUPS % nranks != 0)
        return fail(c, CVB_ERR_INVALID_ARGUMENT, "nranks must divide %d", NGROUPS);
    if (c->comm) return fail(c, CVB_ERR_STATE, "communicator already initialised");
    if (!load_nccl(c->err)) return CVB_ERR_COMM;
    CU(c, cudaSetDevice(c->device));
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof uid);
    NC(c, g_nccl.CommInitRank(&c->comm, nranks, uid, rank));
    c->nranks = nranks;
    c->rank = rank;
    return CVB_OK;
}
extern "C" cvb_status cvb_comm_destroy(cvb_context *c) {
    if (!c) return CVB_ERR_INVALID_ARGUMENT;
    if (c->comm) {
        cudaStreamSynchronize(c->stream);
        g_nccl.CommDestroy(c->comm);
        c->comm = nullptr;
    }
    c->nranks = 1;
    c->rank = 0;
    return CVB_OK;
}

// Exchange HALO rows with the slab above and below: my top halo <- last rows of rank-1, my bottom halo <- first
// rows of rank+1.  base points at plane 0; elem = bytes per element.
static cvb_status exchange_halo(Job *j, void *base, size_t elem, int nplanes) {
    cvb_context *c = j->ctx;
    if (!j->slab || c->nranks == 1) return CVB_OK;
    const Geom &g = j->g;
    const size_t rowb = (size_t)g.pitch * elem, hb = rowb * HALO;
    const int rows = g.row_hi - g.row_lo;
    NC(c, g_nccl.GroupStart());
    for (int p = 0; p < nplanes; ++p) {
        char *pl = (char *)base + (size_t)p * g.plane_elems * elem;
        if (c->rank > 0) {
            NC(c, g_nccl.Send(pl + hb, hb, ncclUint8, c->rank - 1, c->comm, c->stream));              // my first rows up
            NC(c, g_nccl.Recv(pl, hb, ncclUint8, c->rank - 1, c->comm, c->stream));                   // top halo
        }
        if (c->rank < c->nranks - 1) {
            NC(c, g_nccl.Send(pl + rowb * rows, hb, ncclUint8, c->rank + 1, c->comm, c->stream));     // my last rows down
            NC(c, g_nccl.Recv(pl + rowb * (rows + HALO), hb, ncclUint8, c->rank + 1, c->comm, c->stream));
        }
    }
    NC(c, g_nccl.GroupEnd());
    return CVB_OK;
}

// ---- jobs -------------------------------------------------------------------------------------------------------
static void job_free(Job *j) {
    if (!j) return;
    cudaSetDevice(j->ctx->device);
    cudaStreamSynchronize(j->ctx->stream);
    if (j->ctx->copy_stream) cudaStreamSynchronize(j->ctx->copy_stream);
    if (j->ev_prefetched) cudaEventDestroy(j->ev_prefetched);
    if (j->ev_restored) cudaEventDestroy(j->ev_restored);
    cudaFree(j->d_img);
    cudaFree(j->d_img_saved);
    cudaFree(j->d_u[0]);
    cudaFree(j->d_u[1]);
    cudaFree(j->d_pm[0]);
    cudaFree(j->d_pm[1]);
    cudaFree(j->d_aux);
    cudaFree(j->d_state);
    cudaFree(j->d_partials);
    cudaFree(j->d_seg_sums);
    cudaFree(j->d_seg_ticket);
    cudaFree(j->d_group);
    cudaFree(j->d_sign);
    cudaFree(j->d_bits);
    if (!j->ipc_opened.empty()) {
        for (void *p : j->ipc_opened) cudaIpcCloseMemHandle(p);
        j->ipc_opened.clear();
        // nobody frees a buffer a peer may still have mapped: a tiny all-gather is the barrier
        if (j->ctx->comm && j->d_box && g_nccl.AllGather) {
            g_nccl.AllGather(reinterpret_cast<char *>(j->d_box) + j->ctx->rank, j->d_box, 1, ncclUint8, j->ctx->comm, j->ctx->stream);
            cudaStreamSynchronize(j->ctx->stream);
        }
    }
    cudaFree(j->d_box);
    if (j->h_state) cudaFreeHost(j->h_state);
}

#ifdef CSV_TMA
static cvb_status job_make_tensor_maps(Job *j);
#endif
static cvb_status job_init(Job *j, cvb_context *c, int count, int n, int h, int w, int row_lo, int row_hi, bool slab,
                           cvb_precision prec) {
    if (!c) return CVB_ERR_INVALID_ARGUMENT;
    if (count <= 0 || (n != 1 && n != 3) || h <= 0 || w <= 0)
        return fail(c, CVB_ERR_INVALID_ARGUMENT, "bad job shape: count=%d n=%d h=%d w=%d (n must be 1 or 3)", count, n, h, w);
    if (prec != CVB_PRECISION_F64 && prec != CVB_PRECISION_F32) return fail(c, CVB_ERR_INVALID_ARGUMENT, "unknown precision");
    if (prec == CVB_PRECISION_F32 && slab)
        return fail(c, CVB_ERR_INVALID_ARGUMENT, "the fp32 variant runs whole images and batches, not row slabs");
    if (row_lo < 0 || row_hi > h || row_lo >= row_hi) return fail(c, CVB_ERR_INVALID_ARGUMENT, "bad slab rows [%d,%d)", row_lo, row_hi);
    CU(c, cudaSetDevice(c->device));
    j->ctx = c;
    j->prec = prec;
    j->slab = slab;
    Geom &g = j->g;
    g.h = h;
    g.w = w;
    g.row_lo = row_lo;
    g.row_hi = row_hi;
    g.pitch = (w + 15) / 16 * 16;
    g.rows_alloc = row_hi - row_lo + 2 * HALO;
    g.nch = n;
    g.count = count;
    g.seg_rows = c->tile_rows > 0 ? c->tile_rows : auto_seg_rows(h, w, count, slab ? c->nranks : 1);
    g.nseg_global = ceil_div(h, g.seg_rows);
    if (row_lo % g.seg_rows != 0 || (row_hi != h && row_hi % g.seg_rows != 0))
        return fail(c, CVB_ERR_INVALID_ARGUMENT, "slab rows [%d,%d) are not aligned to tile_rows=%d (use cvb_slab_partition)",
                    row_lo, row_hi, g.seg_rows);
    g.seg0 = row_lo / g.seg_rows;
    g.nseg = ceil_div(row_hi, g.seg_rows) - g.seg0;
    g.seg_mult = auto_seg_mult(g.seg_rows, g.nseg, ceil_div(w, CSV_CB), count);
    g.ncb_csv = ceil_div(w, CSV_CB);
    g.ncb_pm = ceil_div(w, PM_CB);
    // PM has no reductions, so its segments need not follow the reduction groups: own wave-aware segment length
    g.pm_seg_rows = auto_pm_seg_rows(row_hi - row_lo, w, count * n, (slab && c->nranks > 1) ? HALO : 0);
    if (const char *e = getenv("CVB_PM_SEG_ROWS"))  // tuning knob (PM results do not depend on the tiling)
        if (atoi(e) >= 4) g.pm_seg_rows = atoi(e);
    g.pm_nseg = ceil_div(row_hi - row_lo, g.pm_seg_rows);
    // the fused two-step PM kernel: 56-column strips, 12 resident CTAs per SM, 4 priming rows more per segment
    g.ncb_pm2 = ceil_div(w, PM2_CB);
    g.pm2_seg_rows = auto_pm_seg_rows(row_hi - row_lo, w, count * n, (slab && c->nranks > 1) ? HALO : 0, PM2_CB, kSlots, 7.5);
    if (const char *e = getenv("CVB_PM2_SEG_ROWS"))
        if (atoi(e) >= 4) g.pm2_seg_rows = atoi(e);
    g.pm2_nseg = ceil_div(row_hi - row_lo, g.pm2_seg_rows);
    g.plane_elems = (long long)g.rows_alloc * g.pitch;
    if ((long long)count * n * std::max(std::max(g.nseg, g.pm_nseg), g.pm2_nseg) * std::max(g.ncb_csv, g.ncb_pm2) > 0x7fffffffLL)
        return fail(c, CVB_ERR_INVALID_ARGUMENT, "job too large for one launch");
    // groups owned by this job
    j->ngroups_local = 0;
    j->group_lo = NGROUPS;
    j->group_hi = 0;
    for (int grp = 0; grp < NGROUPS; ++grp) {
        const int sb = std::max(group_seg_begin(grp, g.nseg_global), g.seg0);
        const int se = std::min(group_seg_begin(grp + 1, g.nseg_global), g.seg0 + g.nseg);
        if (se > sb) {
            ++j->ngroups_local;
            j->group_lo = std::min(j->group_lo, grp);
            j->group_hi = std::max(j->group_hi, grp + 1);
            // a group must not straddle two ranks
            if (group_seg_begin(grp, g.nseg_global) < g.seg0 || group_seg_begin(grp + 1, g.nseg_global) > g.seg0 + g.nseg)
                return fail(c, CVB_ERR_INVALID_ARGUMENT, "slab rows [%d,%d) split reduction group %d (use cvb_slab_partition)",
                            row_lo, row_hi, grp);
        }
    }
    const size_t pe = (size_t)g.plane_elems;
    // TAIL_ROWS rows of slack after the last plane: the CSV row loop fetches and prefetches past the end of its segment
    const size_t tail = (size_t)TAIL_ROWS * g.pitch;
    CU(c, cudaMalloc(&j->d_img, (size_t)count * n * pe + tail));
    CU(c, cudaMalloc(&j->d_u[0], ((size_t)count * pe + tail) * esz(j)));
    CU(c, cudaMalloc(&j->d_u[1], ((size_t)count * pe + tail) * esz(j)));
    CU(c, cudaMalloc(&j->d_state, (size_t)count * sizeof(CsvState)));
    const size_t npart = (size_t)count * g.nseg * g.ncb_csv * WARPS_PER_CTA * NACC;
    CU(c, cudaMalloc(&j->d_partials, npart * sizeof(double)));
    CU(c, cudaMalloc(&j->d_group, 2 * (size_t)NGROUPS * count * NACC * sizeof(double)));
    const size_t nsegs = (size_t)count * g.nseg;
    CU(c, cudaMalloc(&j->d_seg_sums, nsegs * NACC * sizeof(double)));
    CU(c, cudaMalloc(&j->d_seg_ticket, nsegs * sizeof(unsigned int)));
    CU(c, cudaMemsetAsync(j->d_seg_sums, 0, nsegs * NACC * sizeof(double), c->stream));
    CU(c, cudaMemsetAsync(j->d_seg_ticket, 0, nsegs * sizeof(unsigned int), c->stream));
    // a reduction group of more than 128 partial vectors is summed per segment first; keyed to the GLOBAL geometry, so
    // every slab of an image takes the same decision
    j->seg_level = ((long long)ceil_div(g.nseg_global, NGROUPS) * g.ncb_csv > 128) ? 1 : 0;
#ifdef CSV_TMA
    TRY(job_make_tensor_maps(j));
#endif
    CU(c, cudaMallocHost(&j->h_state, 2 * (size_t)count * sizeof(CsvState)));
    CU(c, cudaMemsetAsync(j->d_img, 0, (size_t)count * n * pe + tail, c->stream));
    CU(c, cudaMemsetAsync(j->d_u[0], 0, ((size_t)count * pe + tail) * esz(j), c->stream));
    CU(c, cudaMemsetAsync(j->d_u[1], 0, ((size_t)count * pe + tail) * esz(j), c->stream));
    CU(c, cudaMemsetAsync(j->d_state, 0, (size_t)count * sizeof(CsvState), c->stream));
    CU(c, cudaMemsetAsync(j->d_partials, 0, npart * sizeof(double), c->stream));
    CU(c, cudaMemsetAsync(j->d_group, 0, 2 * (size_t)NGROUPS * count * NACC * sizeof(double), c->stream));
    return CVB_OK;
}

#ifdef CSV_TMA
// Tensor maps of the job's level-set buffers and image planes (build variant CSV_TMA, csv_kernels.cu).
static cvb_status job_make_tensor_maps(Job *j) {
    cvb_context *c = j->ctx;
    typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CU(c, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) return fail(c, CVB_ERR_CUDA, "cuTensorMapEncodeTiled is not available");
    const Geom &g = j->g;
    const cuuint32_t one[3] = {1, 1, 1};
    if (!is_f32(j))
        for (int b = 0; b < 2; ++b) {
            const cuuint64_t dim[2] = {(cuuint64_t)g.pitch, (cuuint64_t)g.count * g.rows_alloc + TAIL_ROWS};
            const cuuint64_t str[1] = {(cuuint64_t)g.pitch * sizeof(double)};
            const cuuint32_t box[2] = {66, 4};
            const CUresult r = ((encode_fn)fn)(&j->tm_u[b], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, j->d_u[b], dim, str, box, one,
                                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                               CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) return fail(c, CVB_ERR_CUDA, "cuTensorMapEncodeTiled(u) failed: %d", (int)r);
        }
    const cuuint64_t dim[3] = {(cuuint64_t)g.pitch, (cuuint64_t)g.rows_alloc, (cuuint64_t)g.count * g.nch};
    const cuuint64_t str[2] = {(cuuint64_t)g.pitch, (cuuint64_t)g.plane_elems};
    const cuuint32_t box[3] = {80, 4, (cuuint32_t)g.nch};
    const CUresult r = ((encode_fn)fn)(&j->tm_img, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, j->d_img, dim, str, box, one,
                                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(c, CVB_ERR_CUDA, "cuTensorMapEncodeTiled(image) failed: %d", (int)r);
    return CVB_OK;
}
#endif

static void fill_args(const Job *j, const cvb_csv_params *p, double tol, CsvArgs &A) {
    memset(&A, 0, sizeof A);
#ifdef CSV_TMA
    A.tm_u[0] = j->tm_u[0];
    A.tm_u[1] = j->tm_u[1];
    A.tm_img = j->tm_img;
#endif
    A.u[0] = j->d_u[0];
    A.u[1] = j->d_u[1];
    A.img = j->d_img;
    A.state = j->d_state;
    A.partials = j->d_partials;
    A.group_sums = j->d_group;
    A.seg_sums = j->d_seg_sums;
    A.seg_ticket = j->d_seg_ticket;
    A.seg_level = j->seg_level;
    A.kappa_out = j->d_aux;
    A.atan_tab = j->ctx->d_atan_tab;
    A.eps = 1.0;
    if (p) {
        // src/main.cpp:985 as OpenCV's MatExpr executes it: addWeighted(kappa, mu*dt, u_diff, (1/N)*dt, -nu*dt)
        A.alpha = p->mu * p->dt;
        A.beta = (1.0 / j->g.nch) * p->dt;
        A.gamma = (-p->nu) * p->dt;
        A.eps = p->eps;
        for (int k = 0; k < MAX_CH; ++k) {
            A.lambda1[k] = p->lambda1[k];
            A.lambda2[k] = p->lambda2[k];
        }
    }
    A.inv_eps = 1.0 / A.eps;
    A.tol = tol;
    A.multi_rank = (j->slab && j->ctx->nranks > 1) ? 1 : 0;
    A.ngroups_local = j->ngroups_local;
    A.group_lo = j->group_lo;
    A.group_hi = j->group_hi;
    A.g = j->g;
    A.cv = j->cv;  // p2p == 0 unless the slab session mapped its peers
}

// all-gather of the group sums + finalize (multi-rank only)
static cvb_status reduce_across_ranks(Job *j, const CsvArgs &A, int mode) {
    cvb_context *c = j->ctx;
    if (!A.multi_rank) return CVB_OK;
    if (j->p2p) return CVB_OK;  // the kernel pushed its group sums to the peers, waited for theirs and folded
    const size_t per_rank = (size_t)(NGROUPS / c->nranks) * j->g.count * NACC;
    NC(c, g_nccl.AllGather(j->d_group + per_rank * c->rank, j->d_group, per_rank, ncclFloat64, c->comm, c->stream));
    CU(c, launch_csv_finalize(A, mode, c->stream));
    c->stats.kernel_launches += 1;
    return CVB_OK;
}

// P2P row slabs: a device-side wait for a peer's flag gives up after SPIN_TIMEOUT_NS and marks the CommBox; report it
// (the stream has been synchronised by the caller)
static cvb_status check_peer_timeout(Job *j) {
    cvb_context *c = j->ctx;
    if (!j->p2p || !j->d_box) return CVB_OK;
    unsigned int flag = 0;
    CommBox box;
    CU(c, cudaMemcpyAsync(&box, j->d_box, sizeof box, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    flag = box.timed_out;
    c->stats.peer_wait_ms += (double)(box.wait_ns - j->seen_wait_ns) * 1e-6;
    c->stats.peer_waits += box.wait_count - j->seen_wait_count;
    j->seen_wait_ns = box.wait_ns;
    j->seen_wait_count = box.wait_count;
    if (flag)
        return fail(c, CVB_ERR_COMM, "rank %d timed out waiting for a neighbouring rank's boundary rows / region sums",
                    c->rank);
    return CVB_OK;
}

static cvb_status job_upload_image(Job *j, const uint8_t *const *planes) {
    cvb_context *c = j->ctx;
    if (!planes) return fail(c, CVB_ERR_INVALID_ARGUMENT, "planes is NULL");
    CU(c, cudaSetDevice(c->device));
    const Geom &g = j->g;
    const int rows = g.row_hi - g.row_lo;
    for (int p = 0; p < g.count * g.nch; ++p) {
        if (!planes[p]) return fail(c, CVB_ERR_INVALID_ARGUMENT, "planes[%d] is NULL", p);
        CU(c, cudaMemcpy2DAsync(j->d_img + (size_t)p * g.plane_elems + (size_t)HALO * g.pitch, g.pitch, planes[p], g.w, g.w,
                                rows, cudaMemcpyHostToDevice, c->stream));
        c->stats.h2d_bytes += (uint64_t)rows * g.w;
    }
    CU(c, launch_replicate_halo(j->d_img, (size_t)g.plane_elems, (size_t)g.pitch, g.count * g.nch, rows, g.row_lo == 0,
                                g.row_hi == g.h, c->stream));
    TRY(exchange_halo(j, j->d_img, 1, g.count * g.nch));
    CU(c, cudaStreamSynchronize(c->stream));  // the caller may reuse its buffers on return
    return CVB_OK;
}
// Perona-Malik smooths the resident planes in place; a saved copy lets a caller re-run from the original image
// without another host-to-device transfer.
static cvb_status job_save_image(Job *j) {
    cvb_context *c = j->ctx;
    CU(c, cudaSetDevice(c->device));
    if (j->prefetch_pending) {  // an unconsumed prefetch is overwritten: let it finish first
        CU(c, cudaStreamWaitEvent(c->stream, j->ev_prefetched, 0));
        j->prefetch_pending = false;
    }
    j->prefetch_needs_halo = false;
    const size_t bytes = (size_t)j->g.count * j->g.nch * j->g.plane_elems;
    if (!j->d_img_saved) CU(c, cudaMalloc(&j->d_img_saved, bytes));
    CU(c, cudaMemcpyAsync(j->d_img_saved, j->d_img, bytes, cudaMemcpyDeviceToDevice, c->stream));
    return CVB_OK;
}
static cvb_status job_restore_image(Job *j) {
    cvb_context *c = j->ctx;
    if (!j->d_img_saved) return fail(c, CVB_ERR_STATE, "no saved image (call save_image first)");
    CU(c, cudaSetDevice(c->device));
    const size_t bytes = (size_t)j->g.count * j->g.nch * j->g.plane_elems;
    if (j->prefetch_pending) {  // the saved copy is being refilled by prefetch_image on the copy stream
        CU(c, cudaStreamWaitEvent(c->stream, j->ev_prefetched, 0));
        j->prefetch_pending = false;
    }
    CU(c, cudaMemcpyAsync(j->d_img, j->d_img_saved, bytes, cudaMemcpyDeviceToDevice, c->stream));
    if (j->ev_restored) CU(c, cudaEventRecord(j->ev_restored, c->stream));
    // a prefetched slab has no neighbour rows yet (a saved copy made by save_image has): exchange them now, collectively
    if (j->prefetch_needs_halo) TRY(exchange_halo(j, j->d_img, 1, j->g.count * j->g.nch));
    return CVB_OK;
}
// The NEXT image of a stream of images: copied from (pinned) host memory into the saved copy on the copy stream while the
// solver works on the current one; restore_image then waits for the copy and makes it the current image.  Whole images
// and batches; for row slabs the halo rows of the slab's first / last rows are filled by the caller's own neighbours'
// uploads -- see the header.  Returns at once: the host buffers must stay untouched until the next restore_image.
static cvb_status job_prefetch_image(Job *j, const uint8_t *const *planes) {
    cvb_context *c = j->ctx;
    if (!planes) return fail(c, CVB_ERR_INVALID_ARGUMENT, "planes is NULL");
    CU(c, cudaSetDevice(c->device));
    const Geom &g = j->g;
    const size_t bytes = (size_t)g.count * g.nch * g.plane_elems;
    if (!j->d_img_saved) {
        CU(c, cudaMalloc(&j->d_img_saved, bytes));
        CU(c, cudaMemsetAsync(j->d_img_saved, 0, bytes, c->stream));
    }
    if (!c->copy_stream) {
        CU(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        for (auto &e : c->copy_ev) CU(c, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    if (!j->ev_prefetched) {
        CU(c, cudaEventCreateWithFlags(&j->ev_prefetched, cudaEventDisableTiming));
        CU(c, cudaEventCreateWithFlags(&j->ev_restored, cudaEventDisableTiming));
        CU(c, cudaEventRecord(j->ev_restored, c->stream));
    }
    // the copy must not overtake a restore that still reads the saved copy (or the memset above)
    CU(c, cudaEventRecord(j->ev_restored, c->stream));
    CU(c, cudaStreamWaitEvent(c->copy_stream, j->ev_restored, 0));
    const int rows = g.row_hi - g.row_lo;
    for (int p = 0; p < g.count * g.nch; ++p) {
        if (!planes[p]) return fail(c, CVB_ERR_INVALID_ARGUMENT, "planes[%d] is NULL", p);
        CU(c, cudaMemcpy2DAsync(j->d_img_saved + (size_t)p * g.plane_elems + (size_t)HALO * g.pitch, g.pitch, planes[p], g.w, g.w,
                                rows, cudaMemcpyHostToDevice, c->copy_stream));
        c->stats.h2d_bytes += (uint64_t)rows * g.w;
    }
    CU(c, launch_replicate_halo(j->d_img_saved, (size_t)g.plane_elems, (size_t)g.pitch, g.count * g.nch, rows, g.row_lo == 0,
                                g.row_hi == g.h, c->copy_stream));
    CU(c, cudaEventRecord(j->ev_prefetched, c->copy_stream));
    j->prefetch_pending = true;
    j->prefetch_needs_halo = j->slab && c->nranks > 1;
    return CVB_OK;
}
static cvb_status job_download_image(Job *j, int first_plane, int nplanes, uint8_t *const *planes) {
    cvb_context *c = j->ctx;
    if (!planes) return fail(c, CVB_ERR_INVALID_ARGUMENT, "planes is NULL");
    CU(c, cudaSetDevice(c->device));
    const Geom &g = j->g;
    const int rows = g.row_hi - g.row_lo;
    for (int p = 0; p < nplanes; ++p) {
        if (!planes[p]) return fail(c, CVB_ERR_INVALID_ARGUMENT, "planes[%d] is NULL", p);
        CU(c, cudaMemcpy2DAsync(planes[p], g.w, j->d_img + (size_t)(first_plane + p) * g.plane_elems + (size_t)HALO * g.pitch,
                                g.pitch, g.w, rows, cudaMemcpyDeviceToHost, c->stream));
        c->stats.d2h_bytes += (uint64_t)rows * g.w;
    }
    CU(c, cudaStreamSynchronize(c->stream));
    return CVB_OK;
}
// u0 into buffer 0 of image `index` (or of all images when index < 0); resets the step counters
static cvb_status job_upload_levelset(Job *j, int index, const double *u) {
    cvb_context *c = j->ctx;
    if (!u) return fail(c, CVB_ERR_INVALID_ARGUMENT, "u is NULL");
    CU(c, cudaSetDevice(c->device));
    const Geom &g = j->g;
    const int rows = g.row_hi - g.row_lo;
    const int lo = index < 0 ? 0 : index, hi = index < 0 ? g.count : index + 1;
    const size_t pbytes = (size_t)g.plane_elems * esz(j);
    if (is_f32(j) && !j->d_aux) CU(c, cudaMalloc(&j->d_aux, (size_t)g.plane_elems * sizeof(double)));
    for (int m = lo; m < hi; ++m) {
        if (m == lo) {
            // fp64: straight into the plane; fp32: through the fp64 staging plane and a conversion kernel
            double *dst = (is_f32(j) ? j->d_aux : reinterpret_cast<double *>(u_plane(j, 0, m))) + (size_t)HALO * g.pitch;
            CU(c, cudaMemcpy2DAsync(dst, g.pitch * sizeof(double), u, g.w * sizeof(double), g.w * sizeof(double), rows,
                                    cudaMemcpyHostToDevice, c->stream));
            c->stats.h2d_bytes += (uint64_t)rows * g.w * sizeof(double);
            if (is_f32(j)) {
                CU(c, launch_convert_d2f(j->d_aux, reinterpret_cast<float *>(u_plane(j, 0, m)), (size_t)g.plane_elems, c->stream));
                c->stats.kernel_launches += 1;
            }
        } else {  // one u0 shared by all images: replicate on the device
            CU(c, cudaMemcpyAsync(u_plane(j, 0, m), u_plane(j, 0, lo), pbytes, cudaMemcpyDeviceToDevice, c->stream));
        }
    }
    CU(c, launch_replicate_halo(u_plane(j, 0, lo), pbytes, (size_t)g.pitch * esz(j), hi - lo, rows, g.row_lo == 0,
                                g.row_hi == g.h, c->stream));
    CU(c, cudaMemsetAsync(j->d_state + lo, 0, (size_t)(hi - lo) * sizeof(CsvState), c->stream));
    CU(c, cudaStreamSynchronize(c->stream));  // the caller may reuse its buffer on return
    return CVB_OK;
}
static cvb_status job_fetch_state(Job *j) {
    cvb_context *c = j->ctx;
    CU(c, cudaMemcpyAsync(j->h_state, j->d_state, (size_t)j->g.count * sizeof(CsvState), cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    return CVB_OK;
}
static cvb_status job_download_levelset(Job *j, int index, double *u) {
    cvb_context *c = j->ctx;
    if (!u || index < 0 || index >= j->g.count) return fail(c, CVB_ERR_INVALID_ARGUMENT, "bad download arguments");
    CU(c, cudaSetDevice(c->device));
    TRY(job_fetch_state(j));
    const Geom &g = j->g;
    const int rows = g.row_hi - g.row_lo;
    const double *src = reinterpret_cast<const double *>(u_plane(j, j->h_state[index].steps_done & 1, index));
    if (is_f32(j)) {
        if (!j->d_aux) CU(c, cudaMalloc(&j->d_aux, (size_t)g.plane_elems * sizeof(double)));
        CU(c, launch_convert_f2d(reinterpret_cast<const float *>(src), j->d_aux, (size_t)g.plane_elems, c->stream));
        c->stats.kernel_launches += 1;
        src = j->d_aux;
    }
    src += (size_t)HALO * g.pitch;
    CU(c, cudaMemcpy2DAsync(u, g.w * sizeof(double), src, g.pitch * sizeof(double), g.w * sizeof(double), rows,
                            cudaMemcpyDeviceToHost, c->stream));
    c->stats.d2h_bytes += (uint64_t)rows * g.w * sizeof(double);
    CU(c, cudaStreamSynchronize(c->stream));
    return CVB_OK;
}
static cvb_status job_mask(Job *j, int index, int invert, uint8_t *mask) {
    cvb_context *c = j->ctx;
    if (!mask || index < 0 || index >= j->g.count) return fail(c, CVB_ERR_INVALID_ARGUMENT, "bad mask arguments");
    CU(c, cudaSetDevice(c->device));
    TRY(job_fetch_state(j));
    const Geom &g = j->g;
    const int rows = g.row_hi - g.row_lo;
    // the mask is written into the idle level-set buffer's storage (as bytes), then copied out
    const int cur = j->h_state[index].steps_done & 1;
    uint8_t *tmp = reinterpret_cast<uint8_t *>(u_plane(j, cur ^ 1, index));
    if (is_f32(j))
        CU(c, launch_mask_f32(reinterpret_cast<const float *>(u_plane(j, cur, index)) + (size_t)HALO * g.pitch, tmp, rows, g.w,
                              g.pitch, invert, c->stream));
    else
        CU(c, launch_mask(reinterpret_cast<const double *>(u_plane(j, cur, index)) + (size_t)HALO * g.pitch, tmp, rows, g.w,
                          g.pitch, invert, c->stream));
    c->stats.kernel_launches += 1;
    CU(c, cudaMemcpy2DAsync(mask, g.w, tmp, g.pitch, g.w, rows, cudaMemcpyDeviceToHost, c->stream));
    c->stats.d2h_bytes += (uint64_t)rows * g.w;
    CU(c, cudaStreamSynchronize(c->stream));
    return CVB_OK;
}

// bit-packed mask (MSB first within a byte, rows padded to whole bytes): 1/8 of the device-to-host traffic
static cvb_status job_mask_packed(Job *j, int index, int invert, uint8_t *bits) {
    cvb_context *c = j->ctx;
    if (!bits || index < 0 || index >= j->g.count) return fail(c, CVB_ERR_INVALID_ARGUMENT, "bad mask arguments");
    CU(c, cudaSetDevice(c->device));
    TRY(job_fetch_state(j));
    const Geom &g = j->g;
    const int rows = g.row_hi - g.row_lo, wb = (g.w + 7) / 8;
    const int cur = j->h_state[index].steps_done & 1;
    uint8_t *tmp = reinterpret_cast<uint8_t *>(u_plane(j, cur ^ 1, index));
    CU(c, launch_mask_packed(u_plane(j, cur, index) + (size_t)HALO * g.pitch * esz(j), is_f32(j) ? 1 : 0, tmp, rows, g.w, g.pitch,
                             invert, c->stream));
    c->stats.kernel_launches += 1;
    CU(c, cudaMemcpyAsync(bits, tmp, (size_t)rows * wb, cudaMemcpyDeviceToHost, c->stream));
    c->stats.d2h_bytes += (uint64_t)rows * wb;
    CU(c, cudaStreamSynchronize(c->stream));
    return CVB_OK;
}

// the packed masks of every image of the job with one launch and one device-to-host copy
static cvb_status job_masks_packed_all(Job *j, int invert, uint8_t *bits) {
    cvb_context *c = j->ctx;
    if (!bits) return fail(c, CVB_ERR_INVALID_ARGUMENT, "bits is NULL");
    CU(c, cudaSetDevice(c->device));
    const Geom &g = j->g;
    if (g.count > 65535) return fail(c, CVB_ERR_INVALID_ARGUMENT, "masks_packed: more than 65535 images");
    const int rows = g.row_hi - g.row_lo, wb = (g.w + 7) / 8;
    const size_t bytes = (size_t)g.count * rows * wb;
    if (!j->d_bits) CU(c, cudaMalloc(&j->d_bits, bytes));
    CU(c, launch_mask_packed_batch(j->d_u[0], j->d_u[1], j->d_state, is_f32(j) ? 1 : 0, j->d_bits, g.count, rows, g.w, g.pitch,
                                   (size_t)g.plane_elems * esz(j), invert, c->stream));
    c->stats.kernel_launches += 1;
    CU(c, cudaMemcpyAsync(bits, j->d_bits, bytes, cudaMemcpyDeviceToHost, c->stream));
    c->stats.d2h_bytes += bytes;
    CU(c, cudaStreamSynchronize(c->stream));
    return CVB_OK;
}

static cvb_status job_init_checkerboard(Job *j) {
    cvb_context *c = j->ctx;
    CU(c, cudaSetDevice(c->device));
    const Geom &g = j->g;
    if (!j->d_sign) {  // the sign vectors depend on h and w only: computed and uploaded once per job
        std::vector<signed char> s((size_t)g.h + g.w);
        // sign(si*sj) = sign(si)*sign(sj) unless the product underflows to zero: |sin| >= ~1e-16 or exactly 0 here
        for (int i = 0; i < g.h; ++i) s[i] = sign_of(sin(kPi * i / 5));
        for (int jx = 0; jx < g.w; ++jx) s[(size_t)g.h + jx] = sign_of(sin(kPi * jx / 5));
        CU(c, cudaMalloc(&j->d_sign, s.size()));
        CU(c, cudaMemcpy(j->d_sign, s.data(), s.size(), cudaMemcpyHostToDevice));
        c->stats.h2d_bytes += s.size();
    }
    for (int m = 0; m < g.count; ++m) {
        if (is_f32(j))
            CU(c, launch_checkerboard_f32(reinterpret_cast<float *>(u_plane(j, 0, m)), j->d_sign, j->d_sign + g.h, g.row_lo,
                                          g.row_hi - g.row_lo, g.w, g.pitch, c->stream));
        else
            CU(c, launch_checkerboard(reinterpret_cast<double *>(u_plane(j, 0, m)), j->d_sign, j->d_sign + g.h, g.row_lo,
                                      g.row_hi - g.row_lo, g.w, g.pitch, c->stream));
        c->stats.kernel_launches += 1;
    }
    CU(c, launch_replicate_halo(j->d_u[0], (size_t)g.plane_elems * esz(j), (size_t)g.pitch * esz(j), g.count,
                                g.row_hi - g.row_lo, g.row_lo == 0, g.row_hi == g.h, c->stream));
    CU(c, cudaMemsetAsync(j->d_state, 0, (size_t)g.count * sizeof(CsvState), c->stream));
    return CVB_OK;  // asynchronous: everything that follows is ordered on the stream
}

// perona_malik, src/main.cpp:478-560, on the resident image planes (in place)
// nsteps PM launches on planes [plane0, plane0 + np) of the resident image (no timing, no synchronisation)
static cvb_status pm_run_planes(Job *j, int plane0, int np, double K, double L, int nsteps) {
    cvb_context *c = j->ctx;
    const Geom &g = j->g;
    const bool strict = c->math == CVB_MATH_STRICT;
    const size_t poff = (size_t)plane0 * g.plane_elems;
    PmArgs A;
    memset(&A, 0, sizeof A);
    A.K = K;
    A.L = L;
    A.inv_k2 = 1.0 / (K * K);
    A.g = g;
    A.g.count = 1;  // the PM kernels only use count * nch = number of planes
    A.g.nch = np;
    A.cv = j->cv;
    if (!j->p2p && j->slab) {  // NCCL halo exchange between the launches: the launcher only needs to know that peers exist
        A.cv.nranks = c->nranks;
        A.cv.rank = c->rank;
    }
    for (int b = 0; b < 2; ++b) {  // the kernels index the neighbours' buffers by plane - plane0, like their own
        if (A.cv.up_pm[b]) A.cv.up_pm[b] += (size_t)plane0 * (size_t)(A.cv.up_rows + 2 * HALO) * g.pitch;
        if (A.cv.dn_pm[b]) A.cv.dn_pm[b] += (size_t)plane0 * (size_t)(A.cv.dn_rows + 2 * HALO) * g.pitch;
    }
    uint8_t *img = j->d_img + poff;
    char *pm[2] = {reinterpret_cast<char *>(j->d_pm[0]) + poff * esz(j), reinterpret_cast<char *>(j->d_pm[1]) + poff * esz(j)};
    // Step 1 reads the uint8 image, step nsteps writes it; the fp64 -> fp64 steps in between run two per launch
    // (pm2_step_kernel, temporal blocking), a left-over one alone.  CVB_PM_FUSE=0: one step per launch throughout.
    const char *fuse_env = getenv("CVB_PM_FUSE");  // read per call: the bit-identity test switches it between two runs
    const bool can_fuse = !(fuse_env && fuse_env[0] == '0') && !is_f32(j) && !strict;
    int cur = -1;  // PM state buffer holding the newest fp64 planes (-1: the uint8 image)
    for (int s = 1; s <= nsteps;) {
        const bool first = s == 1;
        const bool fused = can_fuse && !first && s + 2 <= nsteps;  // steps s, s+1 are both fp64 -> fp64
        const bool last = !fused && s == nsteps && nsteps >= 2;
        const int nxt = cur < 0 ? 0 : cur ^ 1;
        A.in = first ? (const void *)img : (const void *)pm[cur];
        A.out = last ? (void *)img : (void *)pm[nxt];
        A.out_buf = last ? -1 : nxt;
        if (j->p2p) {
            // boundary rows travel inside the kernel (stores into the neighbours' halos + a flag); before a launch
            // that READS pushed rows, one warp waits for both neighbours' flags of the previous launch
            if (!first) {
                CU(c, launch_pm_wait(j->d_box, j->pm_seq, c->rank > 0, c->rank < c->nranks - 1, c->stream));
                c->stats.kernel_launches += 1;
            }
            A.cv.pm_seq = ++j->pm_seq;
        }
        if (is_f32(j))
            CU(c, launch_pm_step_f32(A, first, last, c->stream));
        else if (fused)
            CU(c, launch_pm2_step(A, c->stream));
        else
            CU(c, launch_pm_step(A, first, last, strict, c->stream));
        c->stats.kernel_launches += 1;
        c->stats.pm_step_launches += 1;
        if (last)
            TRY(exchange_halo(j, img, 1, np));
        else if (s + (fused ? 1 : 0) < nsteps && !j->p2p)
            TRY(exchange_halo(j, pm[nxt], sizeof(double), np));
        if (!last) cur = nxt;
        s += fused ? 2 : 1;
    }
    j->pm_cur = cur;
    if (nsteps == 1) {  // u8 -> fp64 -> u8: the single step cannot write the plane it reads
        if (is_f32(j))
            CU(c, launch_quantise_f32(reinterpret_cast<const float *>(pm[0]), img, (size_t)np * g.plane_elems, c->stream));
        else
            CU(c, launch_pm_quantise(reinterpret_cast<const double *>(pm[0]), img, (size_t)np * g.plane_elems, c->stream));
        c->stats.kernel_launches += 1;
        TRY(exchange_halo(j, img, 1, np));
    }
    return CVB_OK;
}
static cvb_status pm_prepare(Job *j, double K, double L, double T, int *steps, int *nsteps_out) {
    cvb_context *c = j->ctx;
    if (!(L > 0.0) || !(K != 0.0)) return fail(c, CVB_ERR_INVALID_ARGUMENT, "perona_malik needs L > 0 and K != 0");
    const int nsteps = cvb_pm_num_steps(L, T);
    if (nsteps < 0) return fail(c, CVB_ERR_INVALID_ARGUMENT, "perona_malik step count overflows");
    if (steps) *steps = nsteps;
    *nsteps_out = nsteps;
    if (nsteps == 0) return CVB_OK;
    CU(c, cudaSetDevice(c->device));
    const Geom &g = j->g;
    // + TAIL_ROWS rows of slack: the PM row ring requests rows past the end of its segment
    const size_t bytes = ((size_t)g.count * g.nch * g.plane_elems + (size_t)TAIL_ROWS * g.pitch) * esz(j);
    for (int b = 0; b < (nsteps > 2 ? 2 : 1); ++b)
        if (!j->d_pm[b]) {
            CU(c, cudaMalloc(&j->d_pm[b], bytes));
            CU(c, cudaMemsetAsync(j->d_pm[b], 0, bytes, c->stream));
        }
    return CVB_OK;
}

// perona_malik, src/main.cpp:478-560, on the resident image planes (in place)
static cvb_status job_perona_malik(Job *j, double K, double L, double T, int *steps) {
    NvtxRange nvtx("cvb.pm");
    cvb_context *c = j->ctx;
    int nsteps = 0;
    TRY(pm_prepare(j, K, L, T, steps, &nsteps));
    if (nsteps == 0) return CVB_OK;  // the reference returns an unset image here; the planes are left unchanged
    CU(c, cudaEventRecord(c->ev[0], c->stream));
    TRY(pm_run_planes(j, 0, j->g.count * j->g.nch, K, L, nsteps));
    CU(c, cudaEventRecord(c->ev[1], c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    float ms = 0;
    cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]);
    c->stats.pm_ms += ms;
    return check_peer_timeout(j);
}

// upload_image + perona_malik with the host-to-device copies hidden behind the diffusion of the planes that have
// already arrived (channels diffuse independently, src/main.cpp:489): planes go up in chunks on a copy stream, each
// chunk's PM launches wait only for that chunk's copy.
static cvb_status job_upload_image_smooth(Job *j, const uint8_t *const *planes, double K, double L, double T, int *steps) {
    NvtxRange nvtx("cvb.upload+pm");
    cvb_context *c = j->ctx;
    if (!planes) return fail(c, CVB_ERR_INVALID_ARGUMENT, "planes is NULL");
    const Geom &g = j->g;
    const int nplanes = g.count * g.nch;
    int nsteps = 0;
    TRY(pm_prepare(j, K, L, T, steps, &nsteps));
    // The overlapped sequence is the default wherever there is something to overlap with (whole images, batches, P2P row
    // slabs of any rank count); CVB_OVERLAP_UPLOAD=0 forces the plain sequence (upload everything, then diffuse).
    const char *overlap_env = getenv("CVB_OVERLAP_UPLOAD");
    const bool multi = j->slab && j->ctx->nranks > 1;
    const bool can = !multi || j->p2p;  // NCCL halo exchanges between the launches: nothing to overlap with
    const bool overlap = can && !(overlap_env && overlap_env[0] == '0');
    if (!overlap || nsteps == 0 || nplanes < 2) {  // nothing to overlap: the plain sequence
        TRY(job_upload_image(j, planes));
        return nsteps ? job_perona_malik(j, K, L, T, nullptr) : CVB_OK;
    }
    if (!c->copy_stream) {
        CU(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        for (auto &e : c->copy_ev) CU(c, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    const int rows = g.row_hi - g.row_lo;
    const int nchunk = std::min(nplanes, 4);
    // the copy stream must not overwrite planes the compute stream may still be using
    CU(c, cudaEventRecord(c->copy_ev[4], c->stream));
    CU(c, cudaStreamWaitEvent(c->copy_stream, c->copy_ev[4], 0));
    CU(c, cudaEventRecord(c->ev[0], c->stream));
    float pm_ms = 0;
    for (int k = 0; k < nchunk; ++k) {
        const int p0 = (int)((long long)nplanes * k / nchunk), p1 = (int)((long long)nplanes * (k + 1) / nchunk);
        for (int p = p0; p < p1; ++p) {
            if (!planes[p]) return fail(c, CVB_ERR_INVALID_ARGUMENT, "planes[%d] is NULL", p);
            CU(c, cudaMemcpy2DAsync(j->d_img + (size_t)p * g.plane_elems + (size_t)HALO * g.pitch, g.pitch, planes[p], g.w, g.w,
                                    rows, cudaMemcpyHostToDevice, c->copy_stream));
            c->stats.h2d_bytes += (uint64_t)rows * g.w;
        }
        CU(c, cudaEventRecord(c->copy_ev[k], c->copy_stream));
        CU(c, cudaStreamWaitEvent(c->stream, c->copy_ev[k], 0));
        CU(c, launch_replicate_halo(j->d_img + (size_t)p0 * g.plane_elems, (size_t)g.plane_elems, (size_t)g.pitch, p1 - p0, rows,
                                    g.row_lo == 0, g.row_hi == g.h, c->stream));
        TRY(exchange_halo(j, j->d_img + (size_t)p0 * g.plane_elems, 1, p1 - p0));  // row slabs: the first step reads halo rows
        TRY(pm_run_planes(j, p0, p1 - p0, K, L, nsteps));
    }
    CU(c, cudaEventRecord(c->ev[1], c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    cudaEventElapsedTime(&pm_ms, c->ev[0], c->ev[1]);
    c->stats.pm_ms += pm_ms;  // includes the exposed part of the copies
    return check_peer_timeout(j);
}
// A P2P slab session has exported its PM state buffers over CUDA IPC and the neighbours push boundary rows into them:
// they live as long as the session (release is refused; freeing them would leave the peers with stale mappings).
static cvb_status job_release_pm(Job *j) {
    if (j->p2p)
        return fail(j->ctx, CVB_ERR_STATE, "release_scratch: the Perona-Malik planes of a multi-GPU slab session are mapped by "
                                           "the neighbouring ranks and stay allocated until the session is destroyed");
    cudaStreamSynchronize(j->ctx->stream);
    cudaFree(j->d_pm[0]);
    cudaFree(j->d_pm[1]);
    j->d_pm[0] = j->d_pm[1] = nullptr;
    return CVB_OK;
}

static cvb_status check_params(cvb_context *c, const cvb_csv_params *p) {
    if (!p) return fail(c, CVB_ERR_INVALID_ARGUMENT, "params is NULL");
    if (!(p->eps > 0.0)) return fail(c, CVB_ERR_INVALID_ARGUMENT, "epsilon must be > 0");
    return CVB_OK;
}

// sums of the current level set -> c1/c2 (mode 2) or the full set-up of a run (mode 1)
static cvb_status job_csv_init(Job *j, const CsvArgs &A, int mode) {
    cvb_context *c = j->ctx;
    CU(c, is_f32(j) ? launch_csv_init_f32(A, mode, c->stream) : launch_csv_init(A, mode, c->stream));
    c->stats.kernel_launches += 1;
    TRY(reduce_across_ranks(j, A, mode));
    return CVB_OK;
}
static cvb_status job_csv_launch_step(Job *j, CsvArgs &A, int step_index /* 0-based: parity of the input buffer */) {
    cvb_context *c = j->ctx;
    A.par = step_index & 1;
    const bool strict = c->math == CVB_MATH_STRICT;
    CU(c, is_f32(j) ? launch_csv_step_f32(A, c->stream) : launch_csv_step(A, strict, c->stream));
    c->stats.kernel_launches += 1;
    c->stats.csv_step_launches += 1;
    if (A.multi_rank) {
        TRY(reduce_across_ranks(j, A, 0));
        if (!j->p2p) TRY(exchange_halo(j, j->d_u[(step_index + 1) & 1], sizeof(double), j->g.count));
    }
    return CVB_OK;
}

// The time-step loop, src/main.cpp:949-1001
static cvb_status job_csv_run(Job *j, const cvb_csv_params *p, double tol, int max_steps, int *steps_done,
                              double *last_norm, cvb_frame_fn frame, void *user) {
    NvtxRange nvtx("cvb.csv");
    cvb_context *c = j->ctx;
    TRY(check_params(c, p));
    CU(c, cudaSetDevice(c->device));
    const Geom &g = j->g;
    if (frame && (g.count != 1 || j->slab || is_f32(j)))
        return fail(c, CVB_ERR_INVALID_ARGUMENT, "frame observer needs a whole single fp64 image");
    CsvArgs A;
    fill_args(j, p, tol, A);
    // a run starts from the level set in buffer (steps_done & 1); move it to buffer 0 if needed
    TRY(job_fetch_state(j));
    for (int m = 0; m < g.count; ++m)
        if (j->h_state[m].steps_done & 1)
            CU(c, cudaMemcpyAsync(u_plane(j, 0, m), u_plane(j, 1, m), (size_t)g.plane_elems * esz(j), cudaMemcpyDeviceToDevice,
                                  c->stream));
    if (A.multi_rank) TRY(exchange_halo(j, j->d_u[0], sizeof(double), g.count));
    TRY(job_csv_init(j, A, 1));
    const long long limit = max_steps < 0 ? (long long)INT_MAX : (long long)max_steps;  // :890
    std::vector<double> frame_buf;
    if (frame) frame_buf.resize((size_t)g.h * g.w);
    long long launched = 0;
    int chunk = frame ? 1 : 8;
    bool all_done = false;
    int pending = -1;  // state snapshot in flight (index into h_state halves)
    CU(c, cudaEventRecord(c->ev[0], c->stream));
    while (launched < limit && !all_done) {
        const int n = (int)std::min<long long>(chunk, limit - launched);
        for (int s = 0; s < n; ++s) TRY(job_csv_launch_step(j, A, (int)((launched + s) & 1)));
        launched += n;
        // snapshot of the states after this chunk; wait for the PREVIOUS chunk's snapshot while this one runs
        const int half = (pending + 1) & 1;
        CU(c, cudaMemcpyAsync(j->h_state + (size_t)half * g.count, j->d_state, (size_t)g.count * sizeof(CsvState),
                              cudaMemcpyDeviceToHost, c->stream));
        CU(c, cudaEventRecord(c->ev[2 + half], c->stream));
        const int wait_half = frame ? half : (pending >= 0 ? pending : -1);
        if (wait_half >= 0) {
            CU(c, cudaEventSynchronize(c->ev[2 + wait_half]));
            all_done = true;
            for (int m = 0; m < g.count; ++m) all_done = all_done && j->h_state[(size_t)wait_half * g.count + m].done;
        }
        pending = half;
        if (frame) {  // vwm.write_frame(u, "t = n"), src/main.cpp:997
            const CsvState &st = j->h_state[(size_t)half * g.count];
            const double *src = j->d_u[st.steps_done & 1] + (size_t)HALO * g.pitch;
            CU(c, cudaMemcpy2DAsync(frame_buf.data(), g.w * sizeof(double), src, g.pitch * sizeof(double), g.w * sizeof(double),
                                    g.h, cudaMemcpyDeviceToHost, c->stream));
            CU(c, cudaStreamSynchronize(c->stream));
            c->stats.d2h_bytes += (uint64_t)g.h * g.w * sizeof(double);
            if (frame(frame_buf.data(), g.h, g.w, st.steps_done, user) != 0)
                return fail(c, CVB_ERR_CALLBACK, "frame observer aborted the run at step %d", st.steps_done);
        }
        chunk = frame ? 1 : std::min(chunk * 2, 64);
    }
    CU(c, cudaEventRecord(c->ev[1], c->stream));
    TRY(job_fetch_state(j));
    float ms = 0;
    cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]);
    c->stats.csv_ms += ms;
    c->stats.d2h_bytes += (uint64_t)g.count * sizeof(CsvState);
    for (int m = 0; m < g.count; ++m) {
        if (steps_done) steps_done[m] = j->h_state[m].steps_done;
        if (last_norm) last_norm[m] = j->h_state[m].norm;
    }
    return check_peer_timeout(j);
}

// The time-step loop with an observer of the SEGMENTATION (the seam of vwm.write_frame, src/main.cpp:997, for consumers
// that draw the contour): after every step the bit-packed mask of the new level set is produced on the device (1/64 of
// the bytes of u), copied on the copy stream into a ring of pinned host slots together with a snapshot of the solver
// state, and handed to the callback in step order while later steps are already running.  The loop never waits for the
// host unless all kRing slots are in flight.
static cvb_status job_csv_run_masks(Job *j, const cvb_csv_params *p, double tol, int max_steps, int *steps_done, double *last_norm,
                                    int rule, cvb_mask_fn fn, void *user) {
    NvtxRange nvtx("cvb.csv+masks");
    cvb_context *c = j->ctx;
    TRY(check_params(c, p));
    const Geom &g = j->g;
    if (!fn || (rule != 0 && rule != 1)) return fail(c, CVB_ERR_INVALID_ARGUMENT, "mask observer: fn is NULL or unknown rule");
    if (g.count != 1 || j->slab) return fail(c, CVB_ERR_INVALID_ARGUMENT, "mask observer needs a whole single image");
    CU(c, cudaSetDevice(c->device));
    constexpr int kRing = 4;
    const int rows = g.h, wb = (g.w + 7) / 8;
    const size_t slot = ((size_t)rows * wb + 255) / 256 * 256;              // bits ...
    const size_t hslot = slot + (sizeof(CsvState) + 255) / 256 * 256;       // ... then the state snapshot
    CsvArgs A;
    fill_args(j, p, tol, A);
    TRY(job_fetch_state(j));
    if (j->h_state[0].steps_done & 1)
        CU(c, cudaMemcpyAsync(u_plane(j, 0, 0), u_plane(j, 1, 0), (size_t)g.plane_elems * esz(j), cudaMemcpyDeviceToDevice, c->stream));
    TRY(job_csv_init(j, A, 1));
    if (!c->copy_stream) {
        CU(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        for (auto &e : c->copy_ev) CU(c, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    struct Ring {  // freed on every exit path
        uint8_t *d = nullptr, *h = nullptr;
        cudaEvent_t done[kRing] = {}, copied[kRing] = {};
        cudaStream_t s0, s1;
        ~Ring() {
            cudaStreamSynchronize(s0);
            cudaStreamSynchronize(s1);
            cudaFree(d);
            if (h) cudaFreeHost(h);
            for (auto e : done)
                if (e) cudaEventDestroy(e);
            for (auto e : copied)
                if (e) cudaEventDestroy(e);
        }
    } R;
    R.s0 = c->stream;
    R.s1 = c->copy_stream;
    CU(c, cudaMalloc(&R.d, kRing * hslot));
    CU(c, cudaMallocHost(&R.h, kRing * hslot));
    for (int k = 0; k < kRing; ++k) {
        CU(c, cudaEventCreateWithFlags(&R.done[k], cudaEventDisableTiming));
        CU(c, cudaEventCreateWithFlags(&R.copied[k], cudaEventDisableTiming));
    }
    const long long limit = max_steps < 0 ? (long long)INT_MAX : (long long)max_steps;  // :890
    long long launched = 0, delivered = 0;
    bool stop_launching = false, finished = limit == 0;
    CU(c, cudaEventRecord(c->ev[0], c->stream));
    auto deliver = [&]() -> cvb_status {  // the oldest slot in flight (its copy has completed)
        const int k = (int)(delivered % kRing);
        CsvState st;
        memcpy(&st, R.h + k * hslot + slot, sizeof st);
        ++delivered;
        if (st.steps_done != delivered) {  // the run had stopped before this launch: it was a no-op, nothing to show
            stop_launching = finished = true;
            return CVB_OK;
        }
        if (st.done) stop_launching = true;  // the breaking step: shown (its update is applied, :994,:1000), then no more
        if (fn(R.h + k * hslot, g.h, g.w, (int)delivered, user) != 0)
            return fail(c, CVB_ERR_CALLBACK, "mask observer aborted the run at step %d", (int)delivered);
        return CVB_OK;
    };
    while (!finished) {
        if (!stop_launching && launched < limit && launched - delivered < kRing) {
            const int k = (int)(launched % kRing);
            TRY(job_csv_launch_step(j, A, (int)(launched & 1)));
            CU(c, launch_mask_packed(u_plane(j, (int)((launched + 1) & 1), 0) + (size_t)HALO * g.pitch * esz(j), is_f32(j) ? 1 : 0,
                                     R.d + k * hslot, rows, g.w, g.pitch, 0, c->stream, rule));
            c->stats.kernel_launches += 1;
            CU(c, cudaMemcpyAsync(R.d + k * hslot + slot, j->d_state, sizeof(CsvState), cudaMemcpyDeviceToDevice, c->stream));
            CU(c, cudaEventRecord(R.done[k], c->stream));
            CU(c, cudaStreamWaitEvent(c->copy_stream, R.done[k], 0));
            CU(c, cudaMemcpyAsync(R.h + k * hslot, R.d + k * hslot, slot + sizeof(CsvState), cudaMemcpyDeviceToHost, c->copy_stream));
            CU(c, cudaEventRecord(R.copied[k], c->copy_stream));
            c->stats.d2h_bytes += (uint64_t)rows * wb + sizeof(CsvState);
            ++launched;
            while (delivered < launched && !finished && cudaEventQuery(R.copied[delivered % kRing]) == cudaSuccess) TRY(deliver());
        } else if (delivered < launched) {
            CU(c, cudaEventSynchronize(R.copied[delivered % kRing]));
            TRY(deliver());
        } else {
            finished = true;  // everything launched has been shown and nothing more may be launched
        }
    }
    CU(c, cudaEventRecord(c->ev[1], c->stream));
    TRY(job_fetch_state(j));
    float ms = 0;
    cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]);
    c->stats.csv_ms += ms;
    if (steps_done) *steps_done = j->h_state[0].steps_done;
    if (last_norm) *last_norm = j->h_state[0].norm;
    return CVB_OK;
}

static cvb_status job_region_means(Job *j, double eps, double *c1, double *c2, int index) {
    cvb_context *c = j->ctx;
    if (!(eps > 0.0) || !c1 || !c2) return fail(c, CVB_ERR_INVALID_ARGUMENT, "bad region_means arguments");
    CU(c, cudaSetDevice(c->device));
    cvb_csv_params p{};
    p.eps = eps;
    CsvArgs A;
    fill_args(j, &p, 0.0, A);
    TRY(job_csv_init(j, A, 2));
    TRY(job_fetch_state(j));
    for (int k = 0; k < j->g.nch; ++k) {
        c1[k] = j->h_state[index].c1[k];
        c2[k] = j->h_state[index].c2[k];
    }
    return CVB_OK;
}

static cvb_status job_csv_step(Job *j, const cvb_csv_params *p, const double *c1, const double *c2, double *norm) {
    cvb_context *c = j->ctx;
    TRY(check_params(c, p));
    if (j->g.count != 1) return fail(c, CVB_ERR_INVALID_ARGUMENT, "csv_step works on a single image");
    CU(c, cudaSetDevice(c->device));
    CsvArgs A;
    fill_args(j, p, 0.0, A);
    TRY(job_fetch_state(j));
    const int before = j->h_state[0].steps_done;
    if (A.multi_rank) TRY(exchange_halo(j, j->d_u[before & 1], sizeof(double), 1));
    TRY(job_csv_init(j, A, 2));  // means of the current u (also refreshes sumI)
    if (c1 && c2) {              // test hook: given means
        CU(c, cudaMemcpyAsync(&j->d_state->c1[0], c1, sizeof(double) * j->g.nch, cudaMemcpyHostToDevice, c->stream));
        CU(c, cudaMemcpyAsync(&j->d_state->c2[0], c2, sizeof(double) * j->g.nch, cudaMemcpyHostToDevice, c->stream));
    }
    // a single step must not be suppressed by a stale done flag / stop value
    const int zero = 0;
    const double neg = -1.0;
    CU(c, cudaMemcpyAsync(&j->d_state->done, &zero, sizeof zero, cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaMemcpyAsync(&j->d_state->stop, &neg, sizeof neg, cudaMemcpyHostToDevice, c->stream));
    TRY(job_csv_launch_step(j, A, before & 1));
    TRY(job_fetch_state(j));
    if (norm) *norm = j->h_state[0].norm;
    return CVB_OK;
}

// Map the peers' buffers of a slab session (CUDA IPC over NVLink).  Collective: every rank calls it for its slab.
struct IpcRecord {
    cudaIpcMemHandle_t u[2], pm[2], group, box;
    int rows;
    int pad[3];
};
static cvb_status job_setup_p2p(Job *j) {
    cvb_context *c = j->ctx;
    const Geom &g = j->g;
    const char *mode = getenv("CVB_COMM");
    if (mode && strcmp(mode, "nccl") == 0) return CVB_OK;  // keep NCCL in the step loop (comparison / fallback)
    if (c->nranks > MAX_RANKS) return CVB_OK;
    const int nplanes = g.count * g.nch;
    const size_t pm_bytes = ((size_t)nplanes * g.plane_elems + (size_t)TAIL_ROWS * g.pitch) * sizeof(double);
    for (int b = 0; b < 2; ++b)
        if (!j->d_pm[b]) {
            CU(c, cudaMalloc(&j->d_pm[b], pm_bytes));
            CU(c, cudaMemsetAsync(j->d_pm[b], 0, pm_bytes, c->stream));
        }
    CU(c, cudaMalloc(&j->d_box, sizeof(CommBox)));
    CU(c, cudaMemsetAsync(j->d_box, 0, sizeof(CommBox), c->stream));
    IpcRecord mine;
    memset(&mine, 0, sizeof mine);
    CU(c, cudaIpcGetMemHandle(&mine.u[0], j->d_u[0]));
    CU(c, cudaIpcGetMemHandle(&mine.u[1], j->d_u[1]));
    CU(c, cudaIpcGetMemHandle(&mine.pm[0], j->d_pm[0]));
    CU(c, cudaIpcGetMemHandle(&mine.pm[1], j->d_pm[1]));
    CU(c, cudaIpcGetMemHandle(&mine.group, j->d_group));
    CU(c, cudaIpcGetMemHandle(&mine.box, j->d_box));
    mine.rows = g.row_hi - g.row_lo;
    // all-gather the records through the communicator that already exists
    IpcRecord *d_rec = nullptr;
    CU(c, cudaMalloc(&d_rec, sizeof(IpcRecord) * c->nranks));
    CU(c, cudaMemcpyAsync(d_rec + c->rank, &mine, sizeof mine, cudaMemcpyHostToDevice, c->stream));
    NC(c, g_nccl.AllGather(d_rec + c->rank, d_rec, sizeof(IpcRecord), ncclUint8, c->comm, c->stream));
    std::vector<IpcRecord> all(c->nranks);
    CU(c, cudaMemcpyAsync(all.data(), d_rec, sizeof(IpcRecord) * c->nranks, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    cudaFree(d_rec);
    CommView &v = j->cv;
    memset(&v, 0, sizeof v);
    v.nranks = c->nranks;
    v.rank = c->rank;
    v.box = j->d_box;
    auto open = [&](const cudaIpcMemHandle_t &h, void **out) -> cudaError_t {
        cudaError_t e = cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess);
        if (e == cudaSuccess) j->ipc_opened.push_back(*out);
        return e;
    };
    for (int p = 0; p < c->nranks; ++p) {
        if (p == c->rank) {
            v.peer_box[p] = j->d_box;
            v.peer_group[p] = j->d_group;
            continue;
        }
        CU(c, open(all[p].box, (void **)&v.peer_box[p]));
        CU(c, open(all[p].group, (void **)&v.peer_group[p]));
    }
    if (c->rank > 0) {
        const IpcRecord &r = all[c->rank - 1];
        v.up_rows = r.rows;
        for (int b = 0; b < 2; ++b) {
            CU(c, open(r.u[b], (void **)&v.up_u[b]));
            CU(c, open(r.pm[b], (void **)&v.up_pm[b]));
        }
    }
    if (c->rank < c->nranks - 1) {
        const IpcRecord &r = all[c->rank + 1];
        v.dn_rows = r.rows;
        for (int b = 0; b < 2; ++b) {
            CU(c, open(r.u[b], (void **)&v.dn_u[b]));
            CU(c, open(r.pm[b], (void **)&v.dn_pm[b]));
        }
    }
    v.p2p = 1;
    j->p2p = true;
    return CVB_OK;
}

// ---- sessions --------------------------------------------------------------------------------------------------
extern "C" cvb_status cvb_session_create_slab(cvb_context *c, int n, int h, int w, int row_lo, int row_hi,
                                              cvb_precision prec, cvb_session **out) {
    if (!c || !out) return CVB_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    cvb_session *s = new cvb_session;
    cvb_status st = job_init(s, c, 1, n, h, w, row_lo, row_hi, true, prec);
    if (st != CVB_OK) {
        s->ctx = c;
        job_free(s);
        delete s;
        return st;
    }
    // the slab must own exactly this rank's groups when a communicator exists
    if (c->nranks > 1) {
        const int per = NGROUPS / c->nranks;
        if (s->group_lo < c->rank * per || s->group_hi > (c->rank + 1) * per) {
            job_free(s);
            delete s;
            return fail(c, CVB_ERR_INVALID_ARGUMENT, "slab rows do not match rank %d of %d (use cvb_slab_partition)", c->rank, c->nranks);
        }
        st = job_setup_p2p(s);
        if (st != CVB_OK) {
            job_free(s);
            delete s;
            return st;
        }
    }
    *out = s;
    return CVB_OK;
}
extern "C" cvb_status cvb_session_create(cvb_context *c, int n, int h, int w, cvb_precision prec, cvb_session **out) {
    if (!c || !out) return CVB_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    cvb_session *s = new cvb_session;
    cvb_status st = job_init(s, c, 1, n, h, w, 0, h, false, prec);
    if (st != CVB_OK) {
        s->ctx = c;
        job_free(s);
        delete s;
        return st;
    }
    *out = s;
    return CVB_OK;
}
extern "C" void cvb_session_destroy(cvb_session *s) {
    if (!s) return;
    job_free(s);
    delete s;
}
extern "C" cvb_status cvb_session_upload_image(cvb_session *s, const uint8_t *const *planes) {
    return s ? job_upload_image(s, planes) : CVB_ERR_INVALID_ARGUMENT;
}
extern "C" cvb_status cvb_session_upload_levelset(cvb_session *s, const double *u) {
    return s ? job_upload_levelset(s, 0, u) : CVB_ERR_INVALID_ARGUMENT;
}
extern "C" cvb_status cvb_session_init_checkerboard(cvb_session *s) {
    return s ? job_init_checkerboard(s) : CVB_ERR_INVALID_ARGUMENT;
}
extern "C" cvb_status cvb_session_perona_malik(cvb_session *s, double K, double L, double T, int *steps) {
    return s ? job_perona_malik(s, K, L, T, steps) : CVB_ERR_INVALID_ARGUMENT;
}
extern "C" cvb_status cvb_session_csv_run(cvb_session *s, const cvb_csv_params *p, double tol, int max_steps,
                                          int *steps_done, double *last_norm, cvb_frame_fn frame, void *user) {
    return s ? job_csv_run(s, p, tol, max_steps, steps_done, last_norm, frame, user) : CVB_ERR_INVALID_ARGUMENT;
}
extern "C" cvb_status cvb_session_csv_run_masks(cvb_session *s, const cvb_csv_params *p, double tol, int max_steps,
                                                int *steps_done, double *last_norm, int rule, cvb_mask_fn fn, void *user) {
    return s ? job_csv_run_masks(s, p, tol, max_steps, steps_done, last_norm, rule, fn, user) : CVB_ERR_INVALID_ARGUMENT;
}
extern "C" cvb_status cvb_session_csv_step(cvb_session *s, const cvb_csv_params *p, const double *c1, const double *c2,
                                           double *norm) {
    return s ? job_csv_step(s, p, c1, c2, norm) : CVB_ERR_INVALID_ARGUMENT;
}
extern "C" cvb_status cvb_session_region_means(cvb_session *s, double eps, double *c1, double *c2) {
    return s ? job_region_means(s, eps, c1, c2, 0) : CVB_ERR_INVALID_ARGUMENT;
}
extern "C" cvb_status cvb_session_download_levelset(cvb_session *s, double *u) {
    return s ? job_download_levelset(s, 0, u) : CVB_ERR_INVALID_ARGUMENT;
}
extern "C" cvb_status cvb_session_download_image(cvb_session *s, uint8_t *const *planes) {
    return s ? job_download_image(s, 0, s->g.nch, planes) : CVB_ERR_INVALID_ARGUMENT;
}
// test hook: the fp64 diffusion state the last (quantising) step of the newest perona_malik run read
extern "C" cvb_status cvb_session_download_pm_state(cvb_session *s, double *const *planes) {
    if (!s) return CVB_ERR_INVALID_ARGUMENT;
    Job *j = s;
    cvb_context *c = j->ctx;
    if (!planes || is_f32(j) || j->pm_cur < 0 || !j->d_pm[j->pm_cur])
        return fail(c, CVB_ERR_STATE, "no fp64 Perona-Malik state (run perona_malik with at least two steps first)");
    CU(c, cudaSetDevice(c->device));
    const Geom &g = j->g;
    const int rows = g.row_hi - g.row_lo;
    for (int p = 0; p < g.nch; ++p) {
        if (!planes[p]) return fail(c, CVB_ERR_INVALID_ARGUMENT, "planes[%d] is NULL", p);
        CU(c, cudaMemcpy2DAsync(planes[p], g.w * sizeof(double), j->d_pm[j->pm_cur] + (size_t)p * g.plane_elems + (size_t)HALO * g.pitch,
                                g.pitch * sizeof(double), g.w * sizeof(double), rows, cudaMemcpyDeviceToHost, c->stream));
    }
    CU(c, cudaStreamSynchronize(c->stream));
    return CVB_OK;
}
extern "C" cvb_status cvb_session_mask(cvb_session *s, int invert, uint8_t *mask) {
    return s ? job_mask(s, 0, invert, mask) : CVB_ERR_INVALID_ARGUMENT;
}
extern "C" cvb_status cvb_session_mask_packed(cvb_session *s, int invert, uint8_t *bits) {
    return s ? job_mask_packed(s, 0, invert, bits) : CVB_ERR_INVALID_ARGUMENT;
}
extern "C" cvb_status cvb_session_upload_image_smooth(cvb_session *s, const uint8_t *const *planes, double K, double L, double T,
                                                      int *steps) {
    return s ? job_upload_image_smooth(s, planes, K, L, T, steps) : CVB_ERR_INVALID_ARGUMENT;
}
extern "C" cvb_status cvb_session_save_image(cvb_session *s) { return s ? job_save_image(s) : CVB_ERR_INVALID_ARGUMENT; }
extern "C" cvb_status cvb_session_restore_image(cvb_session *s) { return s ? job_restore_image(s) : CVB_ERR_INVALID_ARGUMENT; }
extern "C" cvb_status cvb_session_prefetch_image(cvb_session *s, const uint8_t *const *planes) {
    return s ? job_prefetch_image(s, planes) : CVB_ERR_INVALID_ARGUMENT;
}
extern "C" cvb_status cvb_session_release_scratch(cvb_session *s) {
    return s ? job_release_pm(s) : CVB_ERR_INVALID_ARGUMENT;
}

// ---- batches ---------------------------------------------------------------------------------------------------
extern "C" cvb_status cvb_batch_create(cvb_context *c, int count, int n, int h, int w, cvb_precision prec, cvb_batch **out) {
    if (!c || !out) return CVB_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    cvb_batch *b = new cvb_batch;
    cvb_status st = job_init(b, c, count, n, h, w, 0, h, false, prec);
    if (st != CVB_OK) {
        b->ctx = c;
        job_free(b);
        delete b;
        return st;
    }
    *out = b;
    return CVB_OK;
}
extern "C" void cvb_batch_destroy(cvb_batch *b) {
    if (!b) return;
    job_free(b);
    delete b;
}
extern "C" cvb_status cvb_batch_upload_images(cvb_batch *b, const uint8_t *const *planes) {
    return b ? job_upload_image(b, planes) : CVB_ERR_INVALID_ARGUMENT;
}
extern "C" cvb_status cvb_batch_upload_levelset(cvb_batch *b, const double *u0) {
    return b ? job_upload_levelset(b, -1, u0) : CVB_ERR_INVALID_ARGUMENT;
}
extern "C" cvb_status cvb_batch_init_checkerboard(cvb_batch *b) {
    return b ? job_init_checkerboard(b) : CVB_ERR_INVALID_ARGUMENT;
}
extern "C" cvb_status cvb_batch_perona_malik(cvb_batch *b, double K, double L, double T, int *steps) {
    return b ? job_perona_malik(b, K, L, T, steps) : CVB_ERR_INVALID_ARGUMENT;
}
extern "C" cvb_status cvb_batch_csv_run(cvb_batch *b, const cvb_csv_params *p, double tol, int max_steps, int *steps_done,
                                        double *last_norm) {
    return b ? job_csv_run(b, p, tol, max_steps, steps_done, last_norm, nullptr, nullptr) : CVB_ERR_INVALID_ARGUMENT;
}
extern "C" cvb_status cvb_batch_download_levelset(cvb_batch *b, int index, double *u) {
    return b ? job_download_levelset(b, index, u) : CVB_ERR_INVALID_ARGUMENT;
}
extern "C" cvb_status cvb_batch_download_image(cvb_batch *b, int index, uint8_t *const *planes) {
    if (!b || index < 0 || index >= b->g.count) return CVB_ERR_INVALID_ARGUMENT;
    return job_download_image(b, index * b->g.nch, b->g.nch, planes);
}
extern "C" cvb_status cvb_batch_mask(cvb_batch *b, int index, int invert, uint8_t *mask) {
    return b ? job_mask(b, index, invert, mask) : CVB_ERR_INVALID_ARGUMENT;
}
extern "C" cvb_status cvb_batch_mask_packed(cvb_batch *b, int index, int invert, uint8_t *bits) {
    return b ? job_mask_packed(b, index, invert, bits) : CVB_ERR_INVALID_ARGUMENT;
}
extern "C" cvb_status cvb_batch_masks_packed(cvb_batch *b, int invert, uint8_t *bits) {
    return b ? job_masks_packed_all(b, invert, bits) : CVB_ERR_INVALID_ARGUMENT;
}
extern "C" cvb_status cvb_batch_upload_images_smooth(cvb_batch *b, const uint8_t *const *planes, double K, double L, double T,
                                                     int *steps) {
    return b ? job_upload_image_smooth(b, planes, K, L, T, steps) : CVB_ERR_INVALID_ARGUMENT;
}
extern "C" cvb_status cvb_batch_save_images(cvb_batch *b) { return b ? job_save_image(b) : CVB_ERR_INVALID_ARGUMENT; }
extern "C" cvb_status cvb_batch_restore_images(cvb_batch *b) { return b ? job_restore_image(b) : CVB_ERR_INVALID_ARGUMENT; }
extern "C" cvb_status cvb_batch_prefetch_images(cvb_batch *b, const uint8_t *const *planes) {
    return b ? job_prefetch_image(b, planes) : CVB_ERR_INVALID_ARGUMENT;
}
extern "C" cvb_status cvb_batch_release_scratch(cvb_batch *b) {
    return b ? job_release_pm(b) : CVB_ERR_INVALID_ARGUMENT;
}

// ---- one-shot calls on host buffers (the drop-in seams) ------------------------------------------------------
// The one-shot calls run on a whole-image fp64 session owned by the context and reused while the shape stays the same.
static cvb_status oneshot_session(cvb_context *c, int n, int h, int w, cvb_session **out) {
    cvb_session *s = c->oneshot;
    if (s && (s->g.nch != n || s->g.h != h || s->g.w != w)) {
        cvb_session_destroy(s);
        s = c->oneshot = nullptr;
    }
    if (!s) {
        TRY(cvb_session_create(c, n, h, w, CVB_PRECISION_F64, &s));
        c->oneshot = s;
    }
    *out = s;
    return CVB_OK;
}
struct SessionGuard {  // the name is historic: the session is cached, a failed call drops it (its state is unknown)
    cvb_context *c = nullptr;
    cvb_session *s = nullptr;
    bool ok = false;
    ~SessionGuard() {
        if (!ok && c && c->oneshot) cvb_context_trim(c);
    }
};
#define ONESHOT(g, c, n, h, w)  \
    SessionGuard g;             \
    g.c = c;                    \
    TRY(oneshot_session(c, n, h, w, &g.s))

extern "C" cvb_status cvb_perona_malik(cvb_context *c, const uint8_t *const *planes_in, int n, int h, int w, double K,
                                       double L, double T, uint8_t *const *planes_out, int *steps) {
    if (!c) return CVB_ERR_INVALID_ARGUMENT;
    if (!planes_in || !planes_out) return fail(c, CVB_ERR_INVALID_ARGUMENT, "planes is NULL");
    ONESHOT(g, c, n, h, w);
    TRY(cvb_session_upload_image(g.s, planes_in));
    TRY(cvb_session_perona_malik(g.s, K, L, T, steps));
    TRY(cvb_session_download_image(g.s, planes_out));
    g.ok = true;
    return CVB_OK;
}

extern "C" cvb_status cvb_csv_run(cvb_context *c, const uint8_t *const *planes, int n, int h, int w, double *u_inout,
                                  const cvb_csv_params *params, double tol, int max_steps, int *steps_done,
                                  double *last_norm, cvb_frame_fn frame, void *user) {
    if (!c) return CVB_ERR_INVALID_ARGUMENT;
    ONESHOT(g, c, n, h, w);
    TRY(cvb_session_upload_image(g.s, planes));
    TRY(cvb_session_upload_levelset(g.s, u_inout));
    TRY(cvb_session_csv_run(g.s, params, tol, max_steps, steps_done, last_norm, frame, user));
    TRY(cvb_session_download_levelset(g.s, u_inout));
    g.ok = true;
    return CVB_OK;
}

extern "C" cvb_status cvb_csv_run_masks(cvb_context *c, const uint8_t *const *planes, int n, int h, int w, double *u_inout,
                                        const cvb_csv_params *params, double tol, int max_steps, int *steps_done,
                                        double *last_norm, int rule, cvb_mask_fn fn, void *user) {
    if (!c) return CVB_ERR_INVALID_ARGUMENT;
    ONESHOT(g, c, n, h, w);
    TRY(cvb_session_upload_image(g.s, planes));
    TRY(cvb_session_upload_levelset(g.s, u_inout));
    TRY(cvb_session_csv_run_masks(g.s, params, tol, max_steps, steps_done, last_norm, rule, fn, user));
    TRY(cvb_session_download_levelset(g.s, u_inout));
    g.ok = true;
    return CVB_OK;
}

extern "C" cvb_status cvb_segment(cvb_context *c, const uint8_t *const *planes, int n, int h, int w, double *u_inout,
                                  int smooth, double K, double L, double T, uint8_t *const *planes_pm_out,
                                  const cvb_csv_params *params, double tol, int max_steps, int *steps_done,
                                  double *last_norm, int invert, uint8_t *mask_out) {
    if (!c) return CVB_ERR_INVALID_ARGUMENT;
    ONESHOT(g, c, n, h, w);
    TRY(cvb_session_upload_image(g.s, planes));
    TRY(cvb_session_upload_levelset(g.s, u_inout));
    if (smooth) {
        TRY(cvb_session_perona_malik(g.s, K, L, T, nullptr));
        if (planes_pm_out) TRY(cvb_session_download_image(g.s, planes_pm_out));
    }
    TRY(cvb_session_csv_run(g.s, params, tol, max_steps, steps_done, last_norm, nullptr, nullptr));
    TRY(cvb_session_download_levelset(g.s, u_inout));
    if (mask_out) TRY(cvb_session_mask(g.s, invert, mask_out));
    g.ok = true;
    return CVB_OK;
}

extern "C" cvb_status cvb_region_means(cvb_context *c, const uint8_t *const *planes, int n, int h, int w, const double *u,
                                       double eps, double *c1, double *c2) {
    if (!c) return CVB_ERR_INVALID_ARGUMENT;
    ONESHOT(g, c, n, h, w);
    TRY(cvb_session_upload_image(g.s, planes));
    TRY(cvb_session_upload_levelset(g.s, u));
    TRY(cvb_session_region_means(g.s, eps, c1, c2));
    g.ok = true;
    return CVB_OK;
}

extern "C" cvb_status cvb_curvature(cvb_context *c, const double *u, int h, int w, double *kappa) {
    if (!c) return CVB_ERR_INVALID_ARGUMENT;
    if (!u || !kappa) return fail(c, CVB_ERR_INVALID_ARGUMENT, "u / kappa is NULL");
    ONESHOT(g, c, 1, h, w);
    Job *j = g.s;
    TRY(cvb_session_upload_levelset(g.s, u));
    if (!j->d_aux) CU(c, cudaMalloc(&j->d_aux, (size_t)j->g.plane_elems * sizeof(double)));
    CsvArgs A;
    fill_args(j, nullptr, 0.0, A);
    CU(c, launch_csv_kappa(A, c->math == CVB_MATH_STRICT, c->stream));
    c->stats.kernel_launches += 1;
    CU(c, cudaMemcpy2DAsync(kappa, w * sizeof(double), j->d_aux + (size_t)HALO * j->g.pitch, j->g.pitch * sizeof(double),
                            w * sizeof(double), h, cudaMemcpyDeviceToHost, c->stream));
    c->stats.d2h_bytes += (uint64_t)h * w * sizeof(double);
    CU(c, cudaStreamSynchronize(c->stream));
    g.ok = true;
    return CVB_OK;
}

extern "C" cvb_status cvb_delta_map(cvb_context *c, double *data, size_t count, double eps) {
    if (!c) return CVB_ERR_INVALID_ARGUMENT;
    if (!data && count) return fail(c, CVB_ERR_INVALID_ARGUMENT, "data is NULL");
    if (count == 0) return CVB_OK;
    CU(c, cudaSetDevice(c->device));
    double *d = nullptr;
    CU(c, cudaMalloc(&d, count * sizeof(double)));
    cudaError_t e = cudaMemcpyAsync(d, data, count * sizeof(double), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = launch_delta_map(d, count, eps, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(data, d, count * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d);
    c->stats.kernel_launches += 1;
    c->stats.h2d_bytes += count * sizeof(double);
    c->stats.d2h_bytes += count * sizeof(double);
    CU(c, e);
    return CVB_OK;
}

extern "C" cvb_status cvb_stop_condition(cvb_context *c, const uint8_t *const *planes, int n, int h, int w, double tol,
                                         double *stop) {
    if (!c) return CVB_ERR_INVALID_ARGUMENT;
    if (!stop) return fail(c, CVB_ERR_INVALID_ARGUMENT, "stop is NULL");
    ONESHOT(g, c, n, h, w);
    TRY(cvb_session_upload_image(g.s, planes));
    Job *j = g.s;
    cvb_csv_params p{};
    p.eps = 1.0;
    CsvArgs A;
    fill_args(j, &p, tol, A);
    TRY(job_csv_init(j, A, 1));
    TRY(job_fetch_state(j));
    *stop = j->h_state[0].stop;
    g.ok = true;
    return CVB_OK;
}

extern "C" cvb_status cvb_mask(cvb_context *c, const double *u, int h, int w, int invert, uint8_t *mask) {
    if (!c) return CVB_ERR_INVALID_ARGUMENT;
    ONESHOT(g, c, 1, h, w);
    TRY(cvb_session_upload_levelset(g.s, u));
    TRY(cvb_session_mask(g.s, invert, mask));
    g.ok = true;
    return CVB_OK;
}
