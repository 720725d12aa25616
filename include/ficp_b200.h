/* libficp_b200 - C ABI of the B200-native Fractional-ICP hot path.
 *
 * Drop-in boundary for `FractionalICP` of Silviculturalist/CoRegistrationGame (ficp.py:5-154,
 * called from app.py:658-661).  The reference has no native code, so there is no existing FFI to
 * mirror symbol-for-symbol; each entry point below names the reference method it replaces.  The
 * Python class in coregistrationgame_b200/ficp.py binds these with ctypes (INTEGRATION.md shows the
 * stub a maintainer of the reference would add).
 *
 * Conventions
 *   - plain pointers and sizes only; every function returns 0 on success or a negative status:
 *       -1 invalid argument   -2 non-finite coordinate   -3 CUDA failure   -4 size not supported
 *       -5 no CUDA device
 *     and ficp_last_error() returns a human-readable message for the calling thread.
 *   - "host" pointers are caller-owned CPU buffers (numpy arrays); functions ending in _device
 *     take device pointers and only enqueue work on `stream` (a cudaStream_t passed as void*,
 *     NULL = default stream).
 *   - point arrays are row-major float64 with `ld` doubles per row; columns 0,1 are X,Y and
 *     column 2 (when use_z != 0) is the height Z.  Z enters distances only, it is never moved.
 *   - there is no CPU fallback: without a CUDA device every compute call fails with -5/-3.
 */
#ifndef FICP_B200_H
#define FICP_B200_H

#include <stdint.h>

#if defined(__GNUC__)
#define FICP_API __attribute__((visibility("default")))
#else
#define FICP_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ficp_target ficp_target; /* built index over the Layer-2 (CHM) points */
typedef struct ficp_batch ficp_batch;   /* plots x start-pose hypotheses resident on the device */

typedef struct {
    int64_t m;            /* target points */
    int32_t has_z;
    int32_t grid_w, grid_h;
    double cell;          /* cell edge */
    double x0, y0;        /* grid origin */
    double bbox[4];       /* xmin xmax ymin ymax */
    double build_ms;      /* device time of the build kernels */
    int32_t clamped;      /* 1: skewed target - the grid spans a robust core extent, points outside sit in the border cells */
    int32_t max_cell_pts; /* points in the heaviest cell */
} ficp_target_info;

typedef struct {
    int32_t n_stages;          /* 1 or 2 (run() = 2 stages, _iterate() = 1) */
    int32_t max_iterations;    /* per stage, ficp.py:12 */
    int32_t allow_reflection;  /* ficp.py:13 */
    int32_t min_k;             /* hypotheses ending with fewer inliers are not ranked */
    double threshold;          /* ficp.py:11 */
    double window_margin;      /* metres of slack around a plot's footprint staged on-chip; <0 = auto */
    int32_t warps_per_cta;     /* 0 = auto */
    int32_t ctas_per_sm;       /* 0 = auto */
    int32_t disable_window;    /* 1 = never stage a window (every query walks the global grid); for tests */
    int32_t team_warps;        /* warps per ICP at launch: 0 = auto (1 unless the batch has fewer ICPs than the GPU
                                  has warp slots), 1, 2, 4 or 8; results do not depend on it */
    int32_t no_helpers;        /* warps that run out of hypotheses help the ICPs still in flight in their CTA (elastic
                                  kernel): 0 = auto (on for batches below ~10 ICPs per warp slot), 1 = off, 2 = on;
                                  results are identical */
    int32_t trace_passes;      /* > 0: record the first `trace_passes` passes of EVERY ICP of the batch (nearest-neighbour
                                  rows, squared distances, trimmed-subset flags, k, FRMSD) for ficp_batch_trace; 0 = off.
                                  Test instrument: direct parity with ficp.py:69-71 (indices) and :62-63,:133 (inlier set) */
    int32_t cta_per_icp;       /* kernel shape: 0 = auto, 1 = one WARP per ICP (throughput shape, icp_persistent.cu),
                                  2 = one CTA per ICP (latency shape for batches smaller than the machine, icp_team.cu:
                                  every phase of a pass is cooperative).  Results are bit-identical. */
    int32_t reserved;
} ficp_batch_params;

typedef struct {
    double m00, m01, m10, m11; /* final = M (p - centre) + c */
    double cx, cy;
    double frmsd, rmse;        /* FRMSD and trimmed RMSE of the last pass */
    int32_t k;                 /* trimmed subset size of the last pass */
    int32_t passes;            /* NN passes executed = hypothesis-iterations */
    int32_t flags;             /* bit0: some query used the global grid, bit1: window disabled */
    int32_t pad;
} ficp_hyp_result;

typedef struct {
    int32_t n_plots, n_hyp, n_hyp_local;
    int32_t elems_per_lane, match_z, warps_per_cta, ctas, ctas_per_sm, slices_per_plot;
    int32_t window_pts_cap, window_cells_cap;
    int32_t team_warps;        /* warps per ICP at launch chosen by the planner */
    int32_t helpers;           /* 1 = elastic kernel (idle warps help) */
    int32_t trace_passes;      /* passes recorded per ICP (0: trace off) */
    int64_t smem_bytes;
    int64_t rows;
    int32_t trace_stride;      /* entries per pass record of ficp_batch_trace */
    int32_t cta_per_icp;       /* 1 = the CTA-per-ICP kernel was chosen */
    int32_t rows_direct;       /* 1 = src_host was page-locked: uploaded as it is and split into (u, z) on the device */
    int32_t reserved;
} ficp_batch_info;

FICP_API const char* ficp_last_error(void);
FICP_API int ficp_device_count(int32_t* n);
FICP_API int ficp_set_device(int32_t device);
/* sms, L2 bytes, opt-in shared memory per block, SM clock kHz */
FICP_API int ficp_device_props(int32_t* sms, int64_t* l2_bytes, int64_t* smem_optin, int32_t* clock_khz);

/* measurement aid for bench.py: read bandwidth of a buffer of `bytes` that stays in L2, GB/s */
FICP_API int ficp_measure_l2_read_gbs(int64_t bytes, int32_t iters, double* gbs);

/* ---- kernel 1a: grid build.  Replaces cKDTree(target) (ficp.py:69), hoisted out of the loop.
 * pts_per_cell: mean target points per grid cell (<= 0: 2).  Measured optima on B200: 6 (XYZ) / 3 (XY) when the index
 * feeds ficp_batch_* (cells wide enough that a query's 3x3 block settles its search), 3 / 2 for bulk ficp_nn_query;
 * the Python host (TargetIndex(purpose=...)) passes these.  Results never depend on it. */
FICP_API int ficp_target_create(const double* pts_host, int64_t m, int32_t ld, int32_t use_z, double pts_per_cell,
                       void* stream, ficp_target** out);
FICP_API int ficp_target_create_device(const double* pts_dev, int64_t m, int32_t ld, int32_t use_z, double pts_per_cell,
                              void* stream, ficp_target** out);
FICP_API int ficp_target_get_info(const ficp_target* t, ficp_target_info* info);
FICP_API void ficp_target_destroy(ficp_target* t);

/* ---- kernel 1b: NN query.  Replaces tree.query(source, k=1) (ficp.py:70).
 * idx_out: original target row (lowest index among exact ties); dist_out: Euclidean distance. */
FICP_API int ficp_nn_query(const ficp_target* t, const double* q_host, int64_t n, int32_t ld, int32_t use_z,
                  int64_t* idx_out, double* dist_out, void* stream);
FICP_API int ficp_nn_query_device(const ficp_target* t, const double* q_dev, int64_t n, int32_t ld, int32_t use_z,
                         int32_t* idx_dev, double* dist_dev, void* stream);
/* The same with the kernel chosen by the caller (tests, benchmarks): kernel = 0 auto (what the two calls above do: the bulk
 * kernel for batches of >= 65536 queries that are at least half as many as the grid has cells), 1 = one thread per query
 * over the L2-resident grid, 2 = bulk (queries brought into cell order, windows of cells staged in shared memory with
 * cp.async.bulk).  Both return identical bits.  counters_out (optional, 3 x u64, bulk kernel only; makes the call
 * synchronous): queries resolved from a shared-memory window / on the global grid / that needed rings >= 2. */
FICP_API int ficp_nn_query_ex(const ficp_target* t, const double* q_host, int64_t n, int32_t ld, int32_t use_z,
                     int64_t* idx_out, double* dist_out, int32_t kernel, uint64_t* counters_out, void* stream);
FICP_API int ficp_nn_query_device_ex(const ficp_target* t, const double* q_dev, int64_t n, int32_t ld, int32_t use_z,
                            int32_t* idx_dev, double* dist_dev, int32_t kernel, uint64_t* counters_out, void* stream);

/* ---- greedy match-and-remove.  Replaces CHMPlot.remove_matches (chm_plot.py:223-285).
 * For every plot (rows [offsets[p], offsets[p+1]) of trees_host) the trees are visited in order; a tree takes its
 * nearest REMAINING target point (lowest index among ties) and removes it iff distance < thr[row].
 * matched_out[row] = original target index removed by that tree, or -1.  Plots are independent of each other. */
FICP_API int ficp_match_remove(const ficp_target* t, const double* trees_host, const int64_t* offsets, int64_t n_plots,
                      int32_t ld, int32_t use_z, const double* thr_host, int64_t* matched_out, void* stream);

/* ---- radial crop.  Replaces `cdist(coordinates, centre) <= dist` (chm_plot.py:144-148, :306-311).
 * mask_out[i] = 1 iff target row i lies within `dist` of (cx, cy) in XY (Euclidean distance, <=).  m bytes. */
FICP_API int ficp_radial_crop(const ficp_target* t, double cx, double cy, double dist, uint8_t* mask_out, void* stream);

/* ---- kernel 2: trimming.  Replaces find_optimal_fraction / get_n_first_elements (ficp.py:62-63,73-86).
 * weights[k-1] = 1/((k/n)**lambda) (computed by the caller with the reference's own expression).
 * fixed_k > 0 selects a fixed subset size instead of the FRMSD-optimal one.  src/corr may be NULL
 * to obtain only the stable (distance, index) order.  One CTA in shared memory for n <= 8192, a chain of launches over
 * global scratch above (n <= 2^24). */
FICP_API int ficp_select_fraction(const double* src_host, int32_t ld_s, const double* corr_host, int32_t ld_c,
                         const double* dist_host, int64_t n, int32_t md, const double* weights_host,
                         int64_t fixed_k, int64_t* k_out, double* frmsd_out, int64_t* order_out);

/* ---- kernel 3: closed-form rigid 2-D fit.  Replaces compute_optimal_transform_2d (ficp.py:89-110).
 * T9: row-major 3x3 homogeneous transform. */
FICP_API int ficp_fit_rigid2d(const double* src_host, int32_t ld_s, const double* tgt_host, int32_t ld_t, int64_t k,
                     int32_t allow_reflection, double* T9);
/* Replaces apply_transform_2d_xy_only (ficp.py:112-119): XY moved, other columns copied bit-identically. */
FICP_API int ficp_apply_xy(const double* in_host, double* out_host, int64_t n, int32_t ld, const double* T9);
/* Sum of squared coordinate differences over md columns: the sum inside frmsd (ficp.py:58-59). */
FICP_API int ficp_sumsq(const double* a_host, int32_t ld_a, const double* b_host, int32_t ld_b, int64_t k, int32_t md,
               double* out);

/* ---- plots above the persistent kernels' 1024-tree limit: the loop of ficp.py:122-147 driven pass by pass from the host
 * (which keeps the convergence test in the reference's own expressions) over arrays that stay on the device.  Each step
 * runs the stage kernels above on the same values in the same order as the host-buffer entry points - same bits.
 *   create        uploads columns 0..md-1 of the plot; `target[idx]` is gathered from the index's own copy of the target
 *   set_weights   FRMSD weights of the stage (n doubles, 1/((k/n)**lambda), as for ficp_select_fraction)
 *   pass          NN of the current positions + `target[idx]` + trim order + k (ficp.py:65-86, :123-124); sumsq_out = sum
 *                 of squared residuals over the k trimmed rows (the sum inside frmsd, ficp.py:58-59); k_out = 0: none kept
 *   fit_apply     compute_optimal_transform_2d on those rows + apply_transform_2d_xy_only (ficp.py:137-139); T9 = the step
 *   read_xy       current XY of the plot (n x 2) */
typedef struct ficp_stepper ficp_stepper;
FICP_API int ficp_stepper_create(const ficp_target* t, const double* src_host, int64_t n, int32_t ld_s, int32_t md,
                        ficp_stepper** out);
FICP_API int ficp_stepper_set_weights(ficp_stepper* s, const double* weights_host);
FICP_API int ficp_stepper_pass(ficp_stepper* s, int64_t fixed_k, int64_t* k_out, double* sumsq_out);
FICP_API int ficp_stepper_fit_apply(ficp_stepper* s, int32_t allow_reflection, double* T9);
FICP_API int ficp_stepper_read_xy(ficp_stepper* s, double* xy_out);
FICP_API void ficp_stepper_destroy(ficp_stepper* s);

/* ---- host-side plot geometry (no device needed).  What a caller of ficp_batch_create needs for thousands of plots without
 * thousands of numpy calls.  One thread up to 512 K rows (starting a thread costs as much as 50 K rows of the pass, measured), one
 * more per further 512 K rows, at most 4; FICP_HOST_THREADS=n allows up to n threads, one per 32 K rows (1 = always serial).
 *   ficp_plot_centres    centres_out[2p..2p+1] = mean of columns 0,1 of plot p, rows added in order - the bits of
 *                        `rows[:, :2].mean(axis=0)`, the point `Plot.rotate_plot` / `coordinate_flip` turn about
 *                        (trees.py:201-222)
 *   ficp_plot_geometry   the pass ficp_batch_create itself makes over the rows (exported so that it can be checked without a
 *                        GPU): u_out [rows*2] = row - centre, z_out [rows] (use_z), ubar_out [n_plots*2] = mean of u,
 *                        rho_out [n_plots] >= max |u - ubar| (sizes the on-chip window only).  -2 on a non-finite coordinate.
 *                        u_out = NULL: the per-plot values only (the read-only pass of the page-locked route). */
FICP_API int ficp_plot_centres(const double* src_host, int32_t ld, const int64_t* plot_offsets, int64_t n_plots,
                      double* centres_out);
FICP_API int ficp_plot_geometry(const double* src_host, int32_t ld, int32_t use_z, const int64_t* plot_offsets, int64_t n_plots,
                      const double* centres, double* u_out, double* z_out, double* ubar_out, double* rho_out);

/* ---- kernel 4: persistent batched ICP.  Replaces _iterate()/run() (ficp.py:122-154), batched over
 * plots and start-pose hypotheses.
 *   src_host        concatenated plot rows; plot p owns rows [plot_offsets[p], plot_offsets[p+1]).  Page-locked memory
 *                   (cudaHostAlloc / cudaHostRegister, a torch pinned tensor) with ld <= 4 is sent by one DMA as it is and
 *                   split on the device; anything else is staged through a page-locked block of the library - same results
 *   centres         n_plots x 2: the point each hypothesis rotates about (trees.py:165-222); NULL = the mean of each plot's
 *                   first two columns, rows added in order (what ficp_plot_centres returns)
 *   hyp             n_hyp x 6: m00 m01 m10 m11 dx dy ; start pose = M (p - centre) + centre + d
 *   hyp_begin/stride  this process runs hypotheses hyp_begin, hyp_begin+stride, ... (multi-GPU sharding)
 *   weights         for table t and stage s: weights[weight_offsets[t] + s*n_t + (k-1)] = 1/((k/n_t)**lambda_s)
 *   plot_tab        table index of each plot (plots with equal n share a table)
 *   fixed_k         per-plot fixed subset size or NULL */
FICP_API int ficp_batch_create(const ficp_target* t, const double* src_host, int32_t ld, int32_t use_z,
                      const int64_t* plot_offsets, int64_t n_plots, const double* centres, const double* hyp,
                      int64_t n_hyp, int32_t hyp_begin, int32_t hyp_stride, const double* weights,
                      const int64_t* weight_offsets, const int32_t* plot_tab, int32_t n_tabs,
                      const int32_t* fixed_k, const ficp_batch_params* params, int32_t want_final_xy, void* stream,
                      ficp_batch** out);
FICP_API int ficp_batch_get_info(const ficp_batch* b, ficp_batch_info* info);
FICP_API int ficp_batch_run(ficp_batch* b, void* stream);  /* enqueue only */
/* waits for `stream`, then copies out whatever is non-NULL: results [n_plots*n_hyp_local], best_keys
 * [n_plots] ((fp32 score bits << 32) | hypothesis id, min = best), final_xy [rows*2] (only when created
 * with want_final_xy and n_hyp_local == 1), stats [8]: passes, global-path queries, windows disabled,
 * fix-up rounds, queries, queries that needed a search (the others passed the skip test: their previous
 * neighbour was proved to still be the nearest), searched queries that the 3x3 block of cells did not settle
 * (finished by the ring loop or on the global grid), passes whose trim order the CTA-per-ICP kernel rebuilt with its
 * block sort (the others reused / repaired the previous pass's order; 0 for the warp-per-ICP kernel). */
FICP_API int ficp_batch_results(ficp_batch* b, ficp_hyp_result* results, uint64_t* best_keys, double* final_xy,
                       uint64_t* stats, void* stream);
/* device-to-device copy of the per-plot best keys into caller memory (e.g. a torch tensor that is then
 * all-reduced with MIN over NCCL). */
FICP_API int ficp_batch_copy_best_keys_device(ficp_batch* b, void* dst_dev, void* stream);
/* enqueue only: packs this GPU's best registration per plot into caller DEVICE memory, FICP_PACK_WORDS x 8 bytes per plot:
 * word 0 the packed key, words 1..10 the ficp_hyp_result row (80 bytes) of that hypothesis, word 11 the
 * hypothesis-iterations this GPU ran, words 12..13 the translation b of the world-frame transform final = M p + b
 * (b = c - M centre, doubles).  One NCCL all_gather of these records is the whole multi-GPU exchange
 * (coregistrationgame_b200/dist.py); the winner per plot is the record with the smallest key. */
#define FICP_PACK_WORDS 14
FICP_API int ficp_batch_pack_best_device(ficp_batch* b, void* dst_dev, void* stream);
/* waits for `stream` and returns ONLY the best registration per plot (host buffers): packed_out[n_plots * FICP_PACK_WORDS]
 * in the layout of ficp_batch_pack_best_device, stats[8] as in ficp_batch_results.
 * The per-hypothesis table (80 B x plots x hypotheses) stays on the device: what app.py:658-661 needs is the winner. */
FICP_API int ficp_batch_best(ficp_batch* b, uint64_t* packed_out, uint64_t* stats, void* stream);
/* per-pass trace of a batch created with trace_passes > 0 (waits for `stream`).  For ICP c = plot * n_hyp_local + j and
 * pass p < min(passes of that ICP, trace_passes), entry t = (c * trace_passes + p) * trace_stride + tree:
 *   idx_out[t]    original target row of the tree's nearest neighbour (`tree.query`, ficp.py:70; lowest row among ties)
 *   d2_out[t]     its squared distance (sqrt = the distance ficp.py:70 returns, bit for bit)
 *   inlier_out[t] 1 iff the tree is in the trimmed subset `argsort(d)[:k]` of that pass (ficp.py:62-63, :133)
 * and k_out / frmsd_out [c * trace_passes + p] = subset size and FRMSD of the pass (ficp.py:73-86).  NULL = skip. */
FICP_API int ficp_batch_trace(ficp_batch* b, int32_t* idx_out, double* d2_out, uint8_t* inlier_out, int32_t* k_out,
                     double* frmsd_out, void* stream);
FICP_API void ficp_batch_destroy(ficp_batch* b);

#ifdef __cplusplus
}
#endif
#endif /* FICP_B200_H */
