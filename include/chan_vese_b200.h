/*
 * chan_vese_b200 -- C ABI of the B200-native (sm_100a) Perona-Malik + Chan-Sandberg-Vese solvers.
 *
 * The reference (ktht/chan_vese) has no plugin / FFI interface: its solvers are free functions and a
 * loop body inside main() (src/main.cpp:583-1008, README.md:106-107).  The entry points below are the
 * seams a maintainer would bind; each one cites the reference code it replaces.  Conventions carried
 * over from the reference: plain C scalars and pointers; host buffers are caller-owned, contiguous,
 * row-major with row stride = w elements (what cv::Mat::data of a fresh / split Mat gives; the
 * reference indexes ptr[i*w+j] everywhere, e.g. src/main.cpp:269-276,507-547); channels are planar in
 * the order cv::split produced them (B,G,R for colour, one plane for -g); the level set u is fp64 and
 * updated in place.  Device memory is library-owned.  Every function returns a cvb_status; no
 * function exits or throws across the ABI (the CLI keeps the reference's msg_exit convention,
 * src/main.cpp:173-178).  Calls are blocking unless stated; a context is thread-compatible, not
 * thread-safe (the reference calls everything from the main thread).
 *
 * There is NO CPU fallback: every compute entry point fails with CVB_ERR_NO_DEVICE when no CUDA
 * device is present.
 */
#ifndef CHAN_VESE_B200_H
#define CHAN_VESE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CVB_MAX_CHANNELS 3

typedef enum cvb_status {
    CVB_OK = 0,
    CVB_ERR_INVALID_ARGUMENT = 1,
    CVB_ERR_NO_DEVICE = 2,
    CVB_ERR_CUDA = 3,
    CVB_ERR_OUT_OF_MEMORY = 4,
    CVB_ERR_STATE = 5,
    CVB_ERR_COMM = 6,
    CVB_ERR_CALLBACK = 7
} cvb_status;

typedef enum cvb_precision {
    CVB_PRECISION_F64 = 0, /* default; the reference's arithmetic type */
    CVB_PRECISION_F32 = 1  /* level set / PM state stored and computed in fp32, reductions in fp64; reported
                              separately (judged on the mask, not on the fp64 level-set tolerance); whole images
                              and batches only, host buffers stay fp64 */
} cvb_precision;

typedef enum cvb_math_mode {
    CVB_MATH_FAST = 0,  /* FMA contraction, rsqrt/rcp + Newton; the production mode */
    CVB_MATH_STRICT = 1 /* the oracle's operation order, IEEE div/sqrt, no FMA: a test mode that is
                           bit-comparable with oracle/cv_oracle.c apart from atan and sum order */
} cvb_math_mode;

/* Parameters of one CSV run: --mu --nu --dt -e --lambda1 --lambda2 (src/main.cpp:759-764). */
typedef struct cvb_csv_params {
    double mu, nu, dt, eps;
    double lambda1[CVB_MAX_CHANNELS];
    double lambda2[CVB_MAX_CHANNELS];
} cvb_csv_params;

/* Counters since context creation (or the last cvb_context_reset_stats). */
typedef struct cvb_stats {
    uint64_t kernel_launches;   /* kernels of this library launched */
    uint64_t csv_step_launches; /* csv_step kernel launches */
    uint64_t pm_step_launches;  /* pm_step kernel launches */
    double csv_ms;              /* device time of csv_step launches (CUDA events on the library's stream) */
    double pm_ms;               /* device time of pm_step launches */
    uint64_t h2d_bytes, d2h_bytes;
    /* multi-GPU slab sessions: time the folding warp of the step kernels spent between publishing this rank's region sums
     * and seeing the last peer's (the slowest rank + the NVLink round trip), and the number of reductions it covers */
    double peer_wait_ms;
    uint64_t peer_waits;
} cvb_stats;

typedef struct cvb_context cvb_context; /* one per GPU (and per rank) */
typedef struct cvb_session cvb_session; /* one image (or one row slab of it) resident in HBM */
typedef struct cvb_batch cvb_batch;     /* a batch of independent equal-sized images resident in HBM */

/* Per-step observer: the seam of vwm.write_frame(u, "t = n") (src/main.cpp:997).  u is a host copy,
 * valid during the call.  Return non-zero to abort the run (CVB_ERR_CALLBACK). */
typedef int (*cvb_frame_fn)(const double *u, int h, int w, int step, void *user);

/* Per-step observer of the SEGMENTATION: the bit-packed mask (numpy.packbits layout along each row, (w+7)/8 bytes per
 * row) of the level set after `step` steps -- what a consumer that draws the contour needs (src/VideoWriterManager.cpp:
 * 57-75), at 1/64 of the bytes of u.  bits is a pinned host buffer, valid during the call.  Non-zero return aborts. */
typedef int (*cvb_mask_fn)(const uint8_t *bits, int h, int w, int step, void *user);
#define CVB_MASK_SEPARATE 0 /* float32(u) > 0: separate(), src/main.cpp:395-400 */
#define CVB_MASK_CONTOUR 1  /* saturate_cast<uchar>(u) > 0, i.e. u > 0.5: VideoWriterManager::draw_contour, :65-68 */

/* ---- library / context ------------------------------------------------------------------------ */
const char *cvb_version(void);
int cvb_device_count(void); /* number of CUDA devices, 0 if none / no driver */
/* device: CUDA ordinal.  stream: a cudaStream_t to launch on, or NULL for a library-owned stream. */
cvb_status cvb_context_create(int device, void *stream, cvb_context **out);
void cvb_context_destroy(cvb_context *ctx);
/* Message of the last error on this context (ctx == NULL: of the last failed context_create). */
const char *cvb_last_error(const cvb_context *ctx);
cvb_status cvb_context_set_math_mode(cvb_context *ctx, cvb_math_mode mode);
/* Tuning knob: rows per tile (0 = automatic). Tiles are keyed to global rows, so results do not
 * depend on the GPU count as long as it is the same.  Applies to jobs created afterwards. */
cvb_status cvb_context_set_tile_rows(cvb_context *ctx, int rows);
cvb_status cvb_context_get_stats(cvb_context *ctx, cvb_stats *out);
cvb_status cvb_context_reset_stats(cvb_context *ctx);
cvb_status cvb_context_synchronize(cvb_context *ctx);
/* The one-shot calls below keep their device buffers (one whole-image session) in the context between calls of the
 * same shape; this frees them (they are also freed by cvb_context_destroy and replaced when the shape changes). */
cvb_status cvb_context_trim(cvb_context *ctx);
/* Pinned host memory helpers (optional; any host pointer is accepted by the calls below). */
cvb_status cvb_host_alloc(size_t bytes, void **out);
void cvb_host_free(void *p);

/* ---- host-side helpers (no GPU needed) --------------------------------------------------------- */
/* Step count of `for (double t = 0; t < T; t += L)`, src/main.cpp:498. */
int cvb_pm_num_steps(double L, double T);
/* levelset_checkerboard, src/main.cpp:221-233 (glibc sin on the host for bit parity). */
cvb_status cvb_levelset_checkerboard(int h, int w, double *u);
/* InteractiveDataRect::get_levelset, src/InteractiveDataRect.cpp:20-27: 1 inside roi, 0 outside. */
cvb_status cvb_levelset_rect(int h, int w, int x, int y, int rw, int rh, double *u);
/* InteractiveDataCirc::get_levelset, src/InteractiveDataCirc.cpp:18-25: one-pixel ring of 1 on 0. */
cvb_status cvb_levelset_circ(int h, int w, int cx, int cy, int radius, double *u);

/* Rows per tile the library would pick for a job of `count` h x w images (what tile_rows = 0 means).  The tiling
 * fixes the order of the fused sums, so the choice does NOT depend on `nranks` (kept in the signature; any value
 * >= 1 gives the same answer): a single image is tiled so that runs on 1, 2, 4 and 8 GPUs are all efficient and
 * give bit-identical results. */
int cvb_auto_tile_rows(int h, int w, int count, int nranks);
/* Row slab [*row_lo, *row_hi) of rank `rank` of `nranks` (1, 2, 4, 8, 16 or 32) for an image of h rows cut
 * into tiles of tile_rows rows: slabs are unions of the library's 32 fixed reduction groups, so a slab run
 * adds the same partial sums in the same order as the single-GPU run (bit-identical results). */
cvb_status cvb_slab_partition(int h, int tile_rows, int nranks, int rank, int *row_lo, int *row_hi);

/* ---- one-shot calls on host buffers (the drop-in seams) ----------------------------------------- */
/* perona_malik(channels, h, w, K, L, T), src/main.cpp:478-560.  n planar uint8 planes in and out
 * (in and out may alias).  *steps (optional) receives the number of diffusion steps run. */
cvb_status cvb_perona_malik(cvb_context *ctx, const uint8_t *const *planes_in, int n, int h, int w, double K,
                            double L, double T, uint8_t *const *planes_out, int *steps);
/* The time-step loop, src/main.cpp:949-1001, including the stop-condition set-up (:949-960).
 * max_steps < 0 means unlimited (:890).  *steps_done = iterations executed (the breaking step's update
 * is applied, :994,:1000), *last_norm = ||du||_2 of the last executed step.  frame may be NULL. */
cvb_status cvb_csv_run(cvb_context *ctx, const uint8_t *const *planes, int n, int h, int w, double *u_inout,
                       const cvb_csv_params *params, double tol, int max_steps, int *steps_done,
                       double *last_norm, cvb_frame_fn frame, void *user);
/* The same loop with the asynchronous mask observer instead of the synchronous level-set observer: masks are produced
 * on the device after every step, travel on a second stream through a ring of pinned buffers and reach fn in step
 * order while later steps are already running (the solver never waits for the host unless 4 masks are in flight). */
cvb_status cvb_csv_run_masks(cvb_context *ctx, const uint8_t *const *planes, int n, int h, int w, double *u_inout,
                             const cvb_csv_params *params, double tol, int max_steps, int *steps_done,
                             double *last_norm, int rule, cvb_mask_fn fn, void *user);
/* PM (optional) followed by CSV with the smoothed planes staying in HBM: main()'s :939-1001.
 * planes_pm_out (optional) receives the uint8 PM result (the "_pm" image, :946); mask_out (optional)
 * receives separate()'s mask (:395-400). */
cvb_status cvb_segment(cvb_context *ctx, const uint8_t *const *planes, int n, int h, int w, double *u_inout,
                       int smooth, double K, double L, double T, uint8_t *const *planes_pm_out,
                       const cvb_csv_params *params, double tol, int max_steps, int *steps_done,
                       double *last_norm, int invert, uint8_t *mask_out);
/* region_variance for both regions of every channel, src/main.cpp:255-281 (+ :188-194). */
cvb_status cvb_region_means(cvb_context *ctx, const uint8_t *const *planes, int n, int h, int w,
                            const double *u, double eps, double *c1, double *c2);
/* curvature(u, h, w), src/main.cpp:342-375. */
cvb_status cvb_curvature(cvb_context *ctx, const double *u, int h, int w, double *kappa);
/* cv::parallel_for_(Range(0, h*w), ParallelPixelFunction(data, w, delta)), src/main.cpp:989 with
 * src/ParallelPixelFunction.cpp:12-17 and regularized_delta (:204-210): data[i] = delta_eps(data[i]). */
cvb_status cvb_delta_map(cvb_context *ctx, double *data, size_t count, double eps);
/* stop_cond = tol * || mean_k I_k ||_2, src/main.cpp:949-960. */
cvb_status cvb_stop_condition(cvb_context *ctx, const uint8_t *const *planes, int n, int h, int w, double tol,
                              double *stop);
/* The mask rule of separate(), src/main.cpp:395-400: float32(u) > 0, optionally inverted. */
cvb_status cvb_mask(cvb_context *ctx, const double *u, int h, int w, int invert, uint8_t *mask);

/* ---- resident sessions (inputs stay in HBM between calls) ---------------------------------------- */
cvb_status cvb_session_create(cvb_context *ctx, int n, int h, int w, cvb_precision prec, cvb_session **out);
/* A row slab [row_lo, row_hi) of an h x w image (multi-GPU row decomposition). Host buffers passed to
 * the upload/download calls of a slab session hold only the slab's own rows. */
cvb_status cvb_session_create_slab(cvb_context *ctx, int n, int h, int w, int row_lo, int row_hi,
                                   cvb_precision prec, cvb_session **out);
void cvb_session_destroy(cvb_session *s);
cvb_status cvb_session_upload_image(cvb_session *s, const uint8_t *const *planes);
cvb_status cvb_session_upload_levelset(cvb_session *s, const double *u);
/* Checkerboard u0 generated on the device from host-computed sign vectors (bit-identical to
 * cvb_levelset_checkerboard, no H2D of the plane). */
cvb_status cvb_session_init_checkerboard(cvb_session *s);
cvb_status cvb_session_perona_malik(cvb_session *s, double K, double L, double T, int *steps);
cvb_status cvb_session_csv_run(cvb_session *s, const cvb_csv_params *params, double tol, int max_steps,
                               int *steps_done, double *last_norm, cvb_frame_fn frame, void *user);
cvb_status cvb_session_csv_run_masks(cvb_session *s, const cvb_csv_params *params, double tol, int max_steps,
                                     int *steps_done, double *last_norm, int rule, cvb_mask_fn fn, void *user);
/* One CSV step.  c1/c2 == NULL: region means of the current u are used (as the loop does);
 * otherwise the given means are used (test hook).  *norm (optional) receives ||du||_2. */
cvb_status cvb_session_csv_step(cvb_session *s, const cvb_csv_params *params, const double *c1,
                                const double *c2, double *norm);
/* Region means of the session's current u and image: c1[n], c2[n]. */
cvb_status cvb_session_region_means(cvb_session *s, double eps, double *c1, double *c2);
cvb_status cvb_session_download_levelset(cvb_session *s, double *u);
cvb_status cvb_session_download_image(cvb_session *s, uint8_t *const *planes);
/* Test hook: the fp64 diffusion state that the last (quantising) step of the newest perona_malik run read, i.e. the
 * planes after steps - 1 diffusion steps (src/main.cpp:498-550 before :551).  n planes of rows * w doubles. */
cvb_status cvb_session_download_pm_state(cvb_session *s, double *const *planes);
cvb_status cvb_session_mask(cvb_session *s, int invert, uint8_t *mask);
/* The same mask bit-packed: 8 pixels per byte, MSB first, every row padded to (w+7)/8 bytes (numpy.packbits
 * layout along the row) -- 1/8 of the device-to-host traffic.  bits: rows * ((w+7)/8) bytes. */
cvb_status cvb_session_mask_packed(cvb_session *s, int invert, uint8_t *bits);
/* upload_image followed by perona_malik, with the host-to-device copies of later planes hidden behind the diffusion
 * of the planes that have already arrived (channels diffuse independently, src/main.cpp:489).  Pinned host memory
 * (cvb_host_alloc) makes the copies truly asynchronous.  Row slabs: overlapped when the peers are mapped (the default), the plain
 * sequence with CVB_COMM=nccl. */
cvb_status cvb_session_upload_image_smooth(cvb_session *s, const uint8_t *const *planes, double K, double L, double T,
                                           int *steps);
/* Perona-Malik smooths the resident planes in place (as the reference re-splits img at src/main.cpp:945).
 * save_image keeps a device copy of the current planes, restore_image brings it back (device to device). */
cvb_status cvb_session_save_image(cvb_session *s);
cvb_status cvb_session_restore_image(cvb_session *s);
/* A stream of images through one session: prefetch_image copies the NEXT image from (pinned) host memory into the saved
 * copy on a second stream and returns at once, while the solver works on the current image; the next restore_image waits
 * for that copy (on the device, not on the host) and makes it current (row slabs: and exchanges the neighbour rows).
 * The host planes must stay untouched until that restore_image has been called.  The upload of image k+1 is thereby
 * hidden behind Perona-Malik + Chan-Vese of image k. */
cvb_status cvb_session_prefetch_image(cvb_session *s, const uint8_t *const *planes);
/* Frees the fp64 Perona-Malik scratch planes (they are re-allocated on demand).  Not for multi-GPU slab sessions:
 * their planes are mapped by the neighbouring ranks (CUDA IPC) for the session's lifetime -> CVB_ERR_STATE. */
cvb_status cvb_session_release_scratch(cvb_session *s);

/* ---- multi-GPU (one process per GPU; row slabs of one image) -------------------------------------- */
#define CVB_COMM_ID_BYTES 128
/* Rank 0 creates the id and the host application broadcasts it (e.g. torch.distributed / MPI). */
cvb_status cvb_comm_create_id(cvb_context *ctx, void *id_out /* CVB_COMM_ID_BYTES */);
cvb_status cvb_comm_init(cvb_context *ctx, const void *id, int nranks, int rank);
cvb_status cvb_comm_destroy(cvb_context *ctx);

/* ---- batches of independent images (no communication) --------------------------------------------- */
cvb_status cvb_batch_create(cvb_context *ctx, int count, int n, int h, int w, cvb_precision prec,
                            cvb_batch **out);
void cvb_batch_destroy(cvb_batch *b);
/* planes: count*n pointers, image-major (image i, channel k at planes[i*n + k]). */
cvb_status cvb_batch_upload_images(cvb_batch *b, const uint8_t *const *planes);
/* One u0 shared by all images (h*w doubles). */
cvb_status cvb_batch_upload_levelset(cvb_batch *b, const double *u0);
cvb_status cvb_batch_init_checkerboard(cvb_batch *b);
cvb_status cvb_batch_perona_malik(cvb_batch *b, double K, double L, double T, int *steps);
/* steps_done[count], last_norm[count] (each optional). */
cvb_status cvb_batch_csv_run(cvb_batch *b, const cvb_csv_params *params, double tol, int max_steps,
                             int *steps_done, double *last_norm);
cvb_status cvb_batch_download_levelset(cvb_batch *b, int index, double *u);
cvb_status cvb_batch_download_image(cvb_batch *b, int index, uint8_t *const *planes);
cvb_status cvb_batch_mask(cvb_batch *b, int index, int invert, uint8_t *mask);
cvb_status cvb_batch_mask_packed(cvb_batch *b, int index, int invert, uint8_t *bits);
/* The packed masks of ALL images with one launch and one copy: bits = count * h * ((w+7)/8) bytes, image-major. */
cvb_status cvb_batch_masks_packed(cvb_batch *b, int invert, uint8_t *bits);
cvb_status cvb_batch_upload_images_smooth(cvb_batch *b, const uint8_t *const *planes, double K, double L, double T, int *steps);
cvb_status cvb_batch_save_images(cvb_batch *b);
cvb_status cvb_batch_restore_images(cvb_batch *b);
cvb_status cvb_batch_prefetch_images(cvb_batch *b, const uint8_t *const *planes);
cvb_status cvb_batch_release_scratch(cvb_batch *b);

#ifdef __cplusplus
}
#endif
#endif /* CHAN_VESE_B200_H */
