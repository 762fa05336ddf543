// Shared definitions for the sm_100a kernels of chan_vese_b200.
//
// Data layout in HBM (library-owned):
//   every plane (level set u, PM state, uint8 image channel) of every image of a job is stored as
//   rows_alloc x pitch elements, rows_alloc = (row_hi - row_lo) + 2*HALO, local row 0 = global row
//   row_lo - HALO.  pitch is a multiple of 16 elements, so a thread's two adjacent columns form one
//   aligned 16-byte (fp64) access and every row starts on a 128-byte line.  Halo rows hold the
//   neighbouring slab's rows (multi-GPU).  INVARIANT at the global image border: the halo rows above
//   row 0 and below row h-1 hold COPIES of the border row (BORDER_REPLICATE / clamped neighbours of
//   the reference, src/main.cpp:351-354, 527-530) -- the production row rings read them instead of
//   clamping indices.  Every path that writes a plane keeps it: the step kernels (replicate_border_rows,
//   pm_replicate_border) and the upload / initialiser / restore paths (launch_replicate_halo).
//   Planes of one kind are contiguous: plane (image m, channel k) = base + (m*nch+k)*plane_elems.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#ifdef CSV_TMA  // build variant: the CSV row ring is filled by the tensor memory accelerator (csv_kernels.cu)
#include <cuda.h>
#endif

namespace cvb {

constexpr int HALO = 4;            // halo rows above and below a slab (two fused PM steps need 4/4, one 2/2, CSV 2/1)
constexpr int TAIL_ROWS = 12;      // rows of slack after the last plane of the u / image buffers (CSV register + L2 prefetch
                                   // run unconditionally up to CSV_PF + 1 rows past the end of a segment)
constexpr int WARPS_PER_CTA = 1;   // a CTA = one warp = one column strip of one row segment: no block barriers,
                                   // a finished warp frees its slot immediately (up to 32 resident CTAs per SM)
constexpr int CTA_THREADS = WARPS_PER_CTA * 32;
constexpr int STRIP_LANES = 64;    // columns touched by a warp: each lane owns 2 adjacent columns
constexpr int CSV_STRIP_OWN = 62;  // lane 0 is a halo lane (it supplies nx of the column to the left)
constexpr int PM_STRIP_OWN = 60;   // lanes 0 and 31 are halo lanes (radius-2 stencil)
constexpr int CSV_CB = WARPS_PER_CTA * CSV_STRIP_OWN;  // columns per CTA
constexpr int PM_CB = WARPS_PER_CTA * PM_STRIP_OWN;    // columns per CTA
constexpr int PM2_CB = 56;         // fused two-step PM: lanes 0, 1, 30, 31 are halo lanes (radius-4 stencil)
constexpr int MAX_CH = 3;
// Accumulator slots of the fused reductions.  a(u) = atan(u/eps)/pi = H(u) - 1/2 (src/main.cpp:188-194):
//   [0] sum a   [1..3] sum I_k*a   [4] sum du^2 (step) or sum mean_k(I)^2 (init)   [5..7] sum I_k (init)
constexpr int NACC = 8;
constexpr int ACC_A = 0, ACC_IA = 1, ACC_SQ = 4, ACC_I = 5;
constexpr int NGROUPS = 32;        // fixed row groups of the deterministic reduction tree

// Geometry of one job (a batch of `count` equal images, each possibly a row slab).
struct Geom {
    int h, w;            // global image size
    int row_lo, row_hi;  // owned global rows [row_lo, row_hi)
    int pitch;           // elements per row (all planes use the same element pitch)
    int rows_alloc;      // row_hi - row_lo + 2*HALO
    int nch;             // channels
    int count;           // images in the job
    int seg_rows;        // rows per segment; segments start at global rows that are multiples of it.  A segment is the
                         // unit of the fused sums (one partial vector per segment and strip): it fixes their order.
    int seg_mult;        // consecutive segments one CTA of the production CSV kernel marches through (it delivers one
                         // partial vector per segment, so the sums -- and the results -- do not depend on it)
    int nseg;            // segments in [row_lo, row_hi)
    int seg0;            // global index of the first local segment (row_lo / seg_rows)
    int nseg_global;     // segments of the whole image
    int ncb_csv, ncb_pm; // column blocks (CTAs per segment)
    int pm_seg_rows, pm_nseg;  // PM segments: [row_lo + s*pm_seg_rows, ...) (PM has no reduction groups to follow)
    int ncb_pm2, pm2_seg_rows, pm2_nseg;  // tiling of the fused two-step PM kernel (56-column strips)
    long long plane_elems;  // rows_alloc * pitch
};

// Per-image state of a CSV run, living in device memory.
struct CsvState {
    double c1[MAX_CH];      // region means used by the next step (src/main.cpp:973-974)
    double c2[MAX_CH];
    double sumI[MAX_CH];    // sum of I_k over the image (exact integers), set by csv_init
    double sums[NACC];      // last reduced sums
    double norm;            // ||du||_2 of the last executed step (src/main.cpp:993)
    double stop;            // tol * || mean_k I_k ||_2 (src/main.cpp:949-960), set by csv_init
    int done;               // 1 once norm <= stop (src/main.cpp:1000)
    int steps_done;         // executed steps; u^n lives in buffer n & 1
    unsigned int final_ticket;
    unsigned int group_ticket[NGROUPS];
    int pad;
};

// ---- multi-GPU row slabs without NCCL in the step loop: peers' memory mapped through CUDA IPC (NVLink) -----------
constexpr int MAX_RANKS = 8;
// One per slab session, in device memory that the peers can write.  All counters only ever grow.
struct CommBox {
    unsigned int arrive[MAX_RANKS];  // arrive[p] = number of reductions whose group sums rank p has pushed to me
    unsigned int produced;           // reductions my own kernels have produced (and pushed)
    unsigned int finalized;          // reductions folded into CsvState (c1/c2, norm, stop flag)
    unsigned int claimed;            // leader election for the next fold
    int pending_mode;                // csv_finalize_image mode of the newest produced reduction
    unsigned int pm_from_above;      // PM launches whose boundary rows the upper neighbour has pushed into my top halo
    unsigned int pm_from_below;
    unsigned int pm_ticket_up, pm_ticket_dn;  // boundary CTAs of the running PM launch that have finished
    unsigned int timed_out;          // sticky: a wait for a peer's flag gave up after SPIN_TIMEOUT_NS (the host reports
                                     // CVB_ERR_COMM instead of hanging; later waits return at once)
    unsigned int wait_count;         // reductions this rank has waited for ...
    unsigned long long wait_ns;      // ... and the time (globaltimer) its folding warp spent between raising its own flags
                                     // and seeing the last peer's: what the slowest rank and the NVLink round trip cost
};
constexpr unsigned long long SPIN_TIMEOUT_NS = 20ull * 1000ull * 1000ull * 1000ull;
struct CommView {
    int p2p;                         // 0: NCCL path (all-gather + send/recv issued by the host between launches)
    int nranks, rank;
    int up_rows, dn_rows;            // rows owned by the neighbours (upper: its bottom halo starts at local row HALO+up_rows)
    unsigned int pm_seq;             // sequence number of this PM launch; boundary CTAs wait for pm_seq-1 (0: no wait)
    CommBox *box;                    // mine
    CommBox *peer_box[MAX_RANKS];    // peer_box[rank] == box
    double *peer_group[MAX_RANKS];   // peers' (double-buffered) group-sum arrays
    double *up_u[2], *dn_u[2];       // neighbours' level-set ping-pong buffers (nullptr at the image border)
    double *up_pm[2], *dn_pm[2];     // neighbours' PM state buffers
};

struct CsvArgs {
    double *u[2];           // ping-pong level-set buffers, count planes each
    const uint8_t *img;     // count * nch planes
    CsvState *state;        // count
    double *partials;       // [count][nseg][ncb][NACC]
    double *seg_sums;       // [count][nseg][NACC]: one sum per row segment (reduce.cuh, seg_level)
    unsigned int *seg_ticket;  // [count][nseg]: CTAs of a segment that have delivered their partial vector
    double *group_sums;     // [2][NGROUPS][count][NACC] (second copy: P2P double buffering)
    double *kappa_out;      // MODE_KAPPA only
    const double *atan_tab; // ATAN_NQ entries, see math.cuh
    double alpha, beta, gamma;  // mu*dt, (1/N)*dt, -nu*dt (src/main.cpp:985 as one addWeighted)
    double eps, inv_eps;
    double lambda1[MAX_CH], lambda2[MAX_CH];
    double tol;
    int multi_rank;         // 1: stop after the group sums; csv_finalize runs after the all-gather
    int par;                // parity of the step counter of every image that is still running (known to the host, so
                            // the first row loads need not wait for the state load); frozen images exit anyway
    int seg_level;          // 1: partial vectors are summed per segment first (jobs with many vectors per group)
    int ngroups_local;      // non-empty groups owned by this rank
    int group_lo, group_hi; // groups owned by this rank
    Geom g;
    CommView cv;
#ifdef CSV_TMA
    // tensor maps of the level-set buffers (2-D: pitch x all rows of all planes) and of the image (3-D: pitch x rows x planes)
    alignas(64) CUtensorMap tm_u[2];
    alignas(64) CUtensorMap tm_img;
#endif
};

struct PmArgs {
    const void *in;   // double* or uint8_t* (first step), count*nch planes
    void *out;        // double* or uint8_t* (last step)
    double K, L, inv_k2;
    int out_buf;      // index of the PM state buffer being written (for the neighbour pushes), -1: uint8 output
    Geom g;
    CommView cv;
};

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
// first / one-past-last global segment of reduction group grp
__host__ __device__ inline int group_seg_begin(int grp, int nseg_global) {
    return (int)(((long long)grp * nseg_global + NGROUPS - 1) / NGROUPS);
}

}  // namespace cvb
