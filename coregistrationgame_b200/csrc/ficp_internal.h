// Internal (non-ABI) declarations shared by the translation units of libficp_b200.so.
#pragma once
#include <cstdint>
#include <string>
#include <cuda_runtime.h>
#include "ficp_common.cuh"

namespace ficp {

// ---- error plumbing ---------------------------------------------------------------------------
enum Status : int {
    kOk = 0,
    kErrInvalid = -1,    // bad argument (maps to ValueError in the Python shim)
    kErrNonFinite = -2,  // NaN/Inf coordinate (the reference raises ValueError via scipy)
    kErrCuda = -3,       // CUDA runtime failure
    kErrTooLarge = -4,   // size outside what the kernels support
    kErrNoDevice = -5,
};
void set_error(const std::string& msg);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);
#define FICP_CUDA(call)                                                          \
    do {                                                                         \
        cudaError_t _e = (call);                                                 \
        if (_e != cudaSuccess) return ::ficp::cuda_fail(_e, #call, __FILE__, __LINE__); \
    } while (0)

// ---- device memory: stream-ordered allocator with a retaining pool -----------------------------------------
// cudaMalloc/cudaFree of the tens-of-MB index buffers cost 30-150 ms per call pair on B200 (measured, driver
// 580); a PRIVATE cudaMemPool per device (release threshold = never trim; the process-wide default pool is left
// alone) makes repeated index / batch creation reuse memory.  Allocation and release are ordered on the stream the
// caller works on: `stream` is the stream whose subsequent work uses the memory (alloc) / whose prior work was the
// last to use it (free).  Handles that outlive a call (Target, Batch) record an event after every enqueue and wait
// for it before releasing their buffers (see UseEvent), so destroy-after-enqueue on a non-blocking stream is safe.
cudaError_t dev_alloc(void** p, size_t bytes, cudaStream_t stream);
void dev_free(void* p, cudaStream_t stream);
// Stream on which a handle releases its buffers once nothing is in flight (after UseEvent::wait).  Blocks freed on the stream
// they were allocated on are reusable by the next allocation at once; freed on another stream the pool can only reuse them
// opportunistically and now and then grows instead (measured: one end-to-end step in ~25 took 14-40 ms longer inside
// ficp_target_create).  The legacy default stream - what callers pass unless they bring their own - always exists, so a handle
// built on it frees on it; a handle built on a caller's stream (which may be gone by then) frees on the per-thread stream.
inline cudaStream_t release_stream(cudaStream_t alloc_stream) { return alloc_stream == nullptr ? nullptr : cudaStreamPerThread; }
template <class T>
inline cudaError_t dev_alloc_t(T** p, size_t count, cudaStream_t stream) {
    return dev_alloc(reinterpret_cast<void**>(p), sizeof(T) * (count ? count : 1), stream);
}
// "last use" marker of a handle: record(stream) after enqueuing work that touches the handle's buffers, wait()
// before they are released.  After wait() nothing is in flight, so the release may be ordered on any stream.
struct UseEvent {
    cudaEvent_t ev = nullptr;
    void record(cudaStream_t s) {
        if (!ev && cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) { ev = nullptr; cudaStreamSynchronize(s); return; }
        if (cudaEventRecord(ev, s) != cudaSuccess) cudaStreamSynchronize(s);
    }
    void wait() { if (ev) cudaEventSynchronize(ev); }
    ~UseEvent() { if (ev) cudaEventDestroy(ev); }
};

// ---- target index -----------------------------------------------------------------------------
struct Target {
    int device = 0;
    long long m = 0;
    int has_z = 0;
    GridView view{};
    double bbox[4] = {0, 0, 0, 0};  // xmin xmax ymin ymax
    double pts_per_cell = 2.0;
    // owned device buffers
    double2* d_xy = nullptr;   // XY layout
    double4* d_rec = nullptr;  // XYZ layout
    int* d_orig = nullptr;     // XY layout
    unsigned* d_cell_start = nullptr;
    float build_ms = 0.f;  // device time of the build kernels (CUDA events)
    long long max_cell_pts = 0;  // heaviest cell (diagnostic: skew of the target)
    mutable UseEvent used;  // last enqueued work that reads the index (queries, ICP batches)
    cudaStream_t alloc_stream = nullptr;  // stream the buffers were allocated on (see release_stream)
};

// Builds the grid from row-major points (ld doubles per row; columns 0,1[,2]).  `pts` is a host
// pointer unless on_device != 0.
int target_build(const double* pts, int on_device, long long m, int ld, int use_z, double pts_per_cell,
                 cudaStream_t stream, Target** out);
void target_free(Target* t);

// ---- standalone stage kernels (host launchers; all pointers are DEVICE pointers) ----------------
int launch_nn_query(const GridView& v, bool z3, const double* d_q, long long n, int ld, int* d_idx, double* d_dist,
                    double* d_d2, cudaStream_t stream);
// Bulk form (nn_bulk.cu): queries brought into cell order, windows of target cells staged in shared memory by
// cp.async.bulk.  Same outputs, same bits.  `d_counters` (optional, 3 x u64, zeroed by the caller): queries resolved
// from a shared-memory window / on the global grid / that needed rings >= 2.
constexpr long long kBulkMinQueries = 1 << 16;
bool nn_bulk_applies(const GridView& v, long long n);
int launch_nn_query_bulk(const GridView& v, bool z3, const double* d_q, long long n, int ld, int* d_idx, double* d_dist,
                         double* d_d2, unsigned long long* d_counters, cudaStream_t stream);
int launch_radial_crop(const GridView& v, double cx, double cy, double dist, unsigned char* d_mask, cudaStream_t stream);
int launch_match_remove(const GridView& v, bool z3, const double* d_trees, const long long* d_offsets, int n_plots,
                        int ld, const double* d_thr, long long* d_out, int* d_scratch, cudaStream_t stream);
int measure_l2_read_gbs(size_t bytes, int iters, double* gbs);
// Sort by (dist, index) + FRMSD prefix scan + first-minimum k (auto) or fixed k.  One CTA in shared memory up to
// kSelectMaxN rows, a chain of launches over global scratch above (up to kSelectLargeMaxN).
constexpr int kSelectMaxN = 8192;
constexpr int kSelectLargeMaxN = 1 << 24;
int launch_select_fraction(const double* d_src, int ld_s, const double* d_corr, int ld_c, const double* d_dist,
                           int n, int md, const double* d_weights, int fixed_k, long long* d_k_out,
                           double* d_frmsd_out, int* d_order_out, cudaStream_t stream);
int launch_inverse_perm(const GridView& v, int* d_inv, cudaStream_t stream);
int launch_gather_grid_rows(const GridView& v, bool z3, const int* d_inv, const int* d_idx, long long n, double* d_out,
                            cudaStream_t stream);
int launch_fit_rigid2d(const double* d_src, int ld_s, const double* d_tgt, int ld_t, const int* d_sel, int k,
                       int allow_reflection, double* d_T9, cudaStream_t stream);
int launch_apply_xy(const double* d_in, double* d_out, long long n, int ld, const double* d_T9, cudaStream_t stream);
int launch_sumsq(const double* d_a, int ld_a, const double* d_b, int ld_b, const int* d_sel, int k, int md,
                 double* d_out, cudaStream_t stream);

// ---- persistent batched ICP ---------------------------------------------------------------------
constexpr int kMaxStages = 2;
constexpr int kWindowRowsCap = 256;   // grid rows a shared-memory window may span (both persistent kernels)

struct PlotMeta {
    long long off;       // first row of this plot in the concatenated source arrays
    int n;               // trees in the plot
    int tab;             // index of its FRMSD weight table
    double cinx, ciny;   // centre the hypotheses rotate about (source rows are stored as u = p - cin)
    double ubx, uby;     // mean of u (shift point for the cross-covariance sums)
    int wx0, wy0, wx1, wy1;  // window of grid cells staged in shared memory ([x0,x1) x [y0,y1))
    int fixed_k;         // 0 = FRMSD-optimal subset size (reference behaviour); >0 = fixed trim size
    int pad;
};

struct HypResult {
    double m00, m01, m10, m11;  // linear part:  final = M (p - cin) + c
    double cx, cy;              // c
    double frmsd, rmse;         // FRMSD and trimmed RMSE of the final pass
    int k;                      // trimmed subset size of the final pass
    int passes;                 // NN passes executed (= hypothesis-iterations)
    int flags;
    int pad;
};

struct IcpParams {
    GridView grid;
    const double2* src_u;
    const double* src_z;
    const PlotMeta* plots;
    int n_plots;
    const double* hyp;  // [n_hyp][6] m00 m01 m10 m11 dx dy
    int n_hyp, hyp_begin, hyp_stride, n_hyp_local;
    const double* tabs;       // per table: [stage][2][NPAD] (g = c*c/k, then c), lane-permuted
    int n_stages;
    double threshold;
    int max_iter, allow_reflection, min_k;
    HypResult* results;             // [n_plots][n_hyp_local]
    unsigned long long* best_key;   // [n_plots]
    double* final_xy;               // optional [rows][2], only when n_hyp_local == 1
    int* slice_counter;
    int* hyp_counter;               // [n_plots]
    int slices_per_plot, n_slices;
    int wcap_pts, wcap_cells, wcap_rows;
    int slots;                      // elastic kernel: lead warps (ICPs in flight) per CTA; the rest help
    int dyn_leads;                  // rounds are handed out to helpers while at most this many leads are active
    unsigned long long* stats;      // [0] passes [1] queries resolved on the global path [2] window disabled
                                    // [3] trim-order fix-up rounds [4] queries [5] searched [6] deferred
                                    // [7] CTA-per-ICP kernel: passes whose trim order was rebuilt by the block sort
    // optional per-pass trace (tests: direct parity of NN indices / inlier sets, ficp.py:69-71,62-63,133); off: cap 0
    int trace_cap;                  // passes recorded per ICP
    int trace_stride;               // entries per pass record (= NPAD of the launch)
    int* tr_idx;                    // [icp][cap][stride] original target row of every tree's nearest neighbour
    double* tr_d2;                  // [icp][cap][stride] its squared distance
    unsigned char* tr_in;           // [icp][cap][stride] 1 = tree is in the trimmed subset of the pass
    int* tr_k;                      // [icp][cap]
    double* tr_f;                   // [icp][cap] FRMSD of the pass
};

struct IcpLaunch {
    int e;            // elements per lane (NPAD = 32*e)
    bool z3;
    int warps;        // warps per CTA
    int slots;        // lead warps per CTA (= warps unless elastic)
    bool elastic;     // idle warps help the ICPs in flight (icp_kernel<..., true>)
    bool cta_per_icp; // icp_team_kernel: the whole CTA (32 e threads) works on one ICP
    int ctas;         // grid size
    size_t smem;      // dynamic shared memory per CTA
};
// CTA-per-ICP kernel (icp_team.cu): T = 32 e threads work on one ICP
size_t icp_team_smem_bytes(int e, bool z3, int wcap_pts, int wcap_cells, int wcap_rows);
int icp_team_max_ctas_per_sm(int e, bool z3, size_t smem, int* out);
int icp_team_threads(int e);   // threads of its CTA for plots of 32 e tree slots (one or two trees per thread)
int launch_icp_team(const IcpParams& p, int e, bool z3, int ctas, size_t smem, cudaStream_t stream);
int icp_max_warps(int e);
size_t icp_smem_bytes(int e, bool z3, int slots, int wcap_pts, int wcap_cells, int wcap_rows);
int icp_max_ctas_per_sm(int e, bool z3, int warps, bool elastic, size_t smem, int* out);
int launch_icp(const IcpParams& p, const IcpLaunch& l, cudaStream_t stream);
constexpr int kPackWords = 14;   // 8-byte words per plot record of the exchange (== FICP_PACK_WORDS of the C ABI)
int launch_pack_best(const unsigned long long* d_best, const HypResult* d_results, const PlotMeta* d_plots, int n_plots,
                     int n_hyp_local, int hyp_begin, int hyp_stride, const unsigned long long* d_stats,
                     unsigned long long* d_dst, cudaStream_t stream);

int launch_split_rows(const double* d_raw, int ld, const PlotMeta* d_plots, int n_plots, bool z3, double2* d_u, double* d_z,
                      cudaStream_t stream);

}  // namespace ficp
