// extern "C" entry points of libficp_b200.so (declared in include/ficp_b200.h).
// Host-side orchestration only: argument checks, device buffers, launch configuration.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <vector>
#include "../../include/ficp_b200.h"
#include "ficp_internal.h"
#include "batch_prep.h"

namespace ficp {

static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }
int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    char buf[512];
    snprintf(buf, sizeof buf, "CUDA error %d (%s) at %s:%d in %s", (int)e, cudaGetErrorString(e), file, line, what);
    g_last_error = buf;
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) return kErrNoDevice;
    return kErrCuda;
}

namespace {
std::mutex g_pool_mu;
cudaMemPool_t g_pools[64] = {};
bool g_pool_tried[64] = {};
// private pool of the current device (nullptr: creation failed, fall back to the device's default pool untouched)
cudaMemPool_t private_pool() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    std::lock_guard<std::mutex> lk(g_pool_mu);
    if (!g_pool_tried[dev]) {
        g_pool_tried[dev] = true;
        cudaMemPoolProps props{};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = dev;
        cudaMemPool_t pool = nullptr;
        if (cudaMemPoolCreate(&pool, &props) == cudaSuccess) {
            unsigned long long keep = ~0ull;  // never trim: freed blocks stay in the pool for the next index / batch
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
            g_pools[dev] = pool;
        } else {
            cudaGetLastError();
        }
    }
    return g_pools[dev];
}
}  // namespace

cudaError_t dev_alloc(void** p, size_t bytes, cudaStream_t stream) {
    cudaMemPool_t pool = private_pool();
    if (pool) return cudaMallocFromPoolAsync(p, bytes ? bytes : 1, pool, stream);
    return cudaMallocAsync(p, bytes ? bytes : 1, stream);
}
void dev_free(void* p, cudaStream_t stream) {
    if (p) cudaFreeAsync(p, stream);
}

// Page-locked staging blocks for the host->device copies of ficp_batch_create: the rows are written once (u = p - centre)
// and leave by DMA from where they were written, instead of from pageable memory through the driver's own staging copy.
// Blocks are kept for the life of the process (a handful: one per concurrent caller) and handed out under a mutex; a block
// is returned only after the stream that read it has been synchronised.  Above 256 MB, or when pinning fails, the caller
// gets ordinary heap memory.
namespace {
struct StageBlock { void* p; size_t cap; };
std::mutex g_stage_mu;
std::vector<StageBlock> g_stage_free;
}  // namespace
struct HostStage {
    void* p = nullptr;
    size_t cap = 0;
    bool pinned = false;
    cudaStream_t pending_on = nullptr;
    bool pending = false;
    explicit HostStage(size_t bytes) {
        if (bytes == 0) bytes = 8;
        if (bytes <= ((size_t)256 << 20)) {
            {
                std::lock_guard<std::mutex> lk(g_stage_mu);
                for (size_t i = 0; i < g_stage_free.size(); ++i)
                    if (g_stage_free[i].cap >= bytes) {
                        p = g_stage_free[i].p; cap = g_stage_free[i].cap; pinned = true;
                        g_stage_free.erase(g_stage_free.begin() + (long)i);
                        break;
                    }
                if (!p && !g_stage_free.empty()) {   // too small: replace the largest kept block instead of hoarding
                    cudaFreeHost(g_stage_free.back().p);
                    g_stage_free.pop_back();
                }
            }
            if (!p) {
                const size_t want = std::max<size_t>(bytes + bytes / 4, (size_t)1 << 20);
                if (cudaHostAlloc(&p, want, cudaHostAllocPortable) == cudaSuccess) { cap = want; pinned = true; }
                else { cudaGetLastError(); p = nullptr; }
            }
        }
        if (!p) { p = ::operator new(bytes); cap = bytes; pinned = false; }
    }
    ~HostStage() {
        if (pending) cudaStreamSynchronize(pending_on);   // an early error return: the copies may still be reading
        if (!pinned) { ::operator delete(p); return; }
        std::lock_guard<std::mutex> lk(g_stage_mu);
        g_stage_free.push_back({p, cap});
    }
    HostStage(const HostStage&) = delete;
    HostStage& operator=(const HostStage&) = delete;
};

static_assert(sizeof(ficp_hyp_result) == sizeof(HypResult), "ABI struct mismatch");
static_assert(FICP_PACK_WORDS == kPackWords, "ABI constant mismatch");

// RAII device buffer for the host-buffer convenience calls
// (allocated and released on the stream the call works on: stream-ordered, so an early error return never frees
// memory a kernel enqueued before it still uses)
template <class T>
struct DevBuf {
    T* p = nullptr;
    cudaStream_t s;
    explicit DevBuf(cudaStream_t stream = nullptr) : s(stream) {}
    ~DevBuf() { dev_free(p, s); }
    int alloc(size_t n) {
        FICP_CUDA(dev_alloc_t(&p, n, s));
        return kOk;
    }
};

struct Batch {
    const Target* tgt = nullptr;
    int n_plots = 0, n_hyp = 0, n_hyp_local = 0;
    long long rows = 0;
    bool z3 = false;
    bool want_final = false;
    bool rows_direct = false;   // the rows went to the device as the caller holds them and were split there
    IcpParams params{};
    IcpLaunch launch{};
    int ctas_per_sm = 0;
    double2* d_src_u = nullptr;
    double* d_src_z = nullptr;
    PlotMeta* d_plots = nullptr;
    double* d_hyp = nullptr;
    double* d_tabs = nullptr;
    HypResult* d_results = nullptr;
    unsigned long long* d_best = nullptr;
    double* d_final = nullptr;
    int* d_counters = nullptr;  // [0] slice counter, [1..n_plots] hypothesis counters
    unsigned long long* d_stats = nullptr;
    int trace_cap = 0, trace_stride = 0;
    int* d_tr_idx = nullptr;
    double* d_tr_d2 = nullptr;
    unsigned char* d_tr_in = nullptr;
    int* d_tr_k = nullptr;
    double* d_tr_f = nullptr;
    UseEvent used;  // last enqueued run / copy that touches the buffers below
    cudaStream_t alloc_stream = nullptr;
    ~Batch() {
        used.wait();  // nothing in flight any more: the release below may be ordered on any stream
        cudaStream_t s = release_stream(alloc_stream);
        dev_free(d_tr_idx, s); dev_free(d_tr_d2, s); dev_free(d_tr_in, s); dev_free(d_tr_k, s); dev_free(d_tr_f, s);
        dev_free(d_src_u, s); dev_free(d_src_z, s); dev_free(d_plots, s); dev_free(d_hyp, s); dev_free(d_tabs, s);
        dev_free(d_results, s); dev_free(d_best, s); dev_free(d_final, s); dev_free(d_counters, s); dev_free(d_stats, s);
    }
};

static int pick_e(int max_n) {
    int e = 1;
    while (32 * e < max_n) e <<= 1;
    return e;
}

}  // namespace ficp

using namespace ficp;

extern "C" {

const char* ficp_last_error(void) { return g_last_error.c_str(); }

int ficp_device_count(int32_t* n) {
    int c = 0;
    cudaError_t e = cudaGetDeviceCount(&c);
    if (e != cudaSuccess) {
        *n = 0;
        return cuda_fail(e, "cudaGetDeviceCount", __FILE__, __LINE__);
    }
    *n = c;
    return kOk;
}

int ficp_set_device(int32_t device) {
    FICP_CUDA(cudaSetDevice(device));
    return kOk;
}

int ficp_device_props(int32_t* sms, int64_t* l2_bytes, int64_t* smem_optin, int32_t* clock_khz) {
    int dev = 0, v = 0;
    FICP_CUDA(cudaGetDevice(&dev));
    FICP_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev));
    if (sms) *sms = v;
    FICP_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrL2CacheSize, dev));
    if (l2_bytes) *l2_bytes = v;
    FICP_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    if (smem_optin) *smem_optin = v;
    FICP_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrClockRate, dev));
    if (clock_khz) *clock_khz = v;
    return kOk;
}

int ficp_measure_l2_read_gbs(int64_t bytes, int32_t iters, double* gbs) {
    if (!gbs || bytes < 4096 || iters < 1) { set_error("ficp_measure_l2_read_gbs: bad arguments"); return kErrInvalid; }
    return measure_l2_read_gbs((size_t)bytes, iters, gbs);
}

// ------------------------------------------------------------------------------------------ target
int ficp_target_create(const double* pts_host, int64_t m, int32_t ld, int32_t use_z, double pts_per_cell,
                       void* stream, ficp_target** out) {
    if (!out || (m > 0 && !pts_host)) { set_error("ficp_target_create: null pointer"); return kErrInvalid; }
    Target* t = nullptr;
    const int rc = target_build(pts_host, 0, m, ld, use_z, pts_per_cell, (cudaStream_t)stream, &t);
    *out = reinterpret_cast<ficp_target*>(t);
    return rc;
}

int ficp_target_create_device(const double* pts_dev, int64_t m, int32_t ld, int32_t use_z, double pts_per_cell,
                              void* stream, ficp_target** out) {
    if (!out || (m > 0 && !pts_dev)) { set_error("ficp_target_create_device: null pointer"); return kErrInvalid; }
    Target* t = nullptr;
    const int rc = target_build(pts_dev, 1, m, ld, use_z, pts_per_cell, (cudaStream_t)stream, &t);
    *out = reinterpret_cast<ficp_target*>(t);
    return rc;
}

int ficp_target_get_info(const ficp_target* th, ficp_target_info* info) {
    if (!th || !info) { set_error("ficp_target_get_info: null pointer"); return kErrInvalid; }
    const Target* t = reinterpret_cast<const Target*>(th);
    info->m = t->m;
    info->has_z = t->has_z;
    info->grid_w = t->view.g.gw;
    info->grid_h = t->view.g.gh;
    info->cell = t->view.g.h;
    info->x0 = t->view.g.x0;
    info->y0 = t->view.g.y0;
    for (int i = 0; i < 4; ++i) info->bbox[i] = t->bbox[i];
    info->build_ms = t->build_ms;
    info->clamped = t->view.g.clamped;
    info->max_cell_pts = (int32_t)std::min<long long>(t->max_cell_pts, 2147483647LL);
    return kOk;
}

void ficp_target_destroy(ficp_target* t) { target_free(reinterpret_cast<Target*>(t)); }

// ------------------------------------------------------------------------------------------ NN query
int ficp_nn_query_device_ex(const ficp_target* th, const double* q_dev, int64_t n, int32_t ld, int32_t use_z,
                            int32_t* idx_dev, double* dist_dev, int32_t kernel, uint64_t* counters_out, void* stream) {
    if (!th) { set_error("ficp_nn_query: null target"); return kErrInvalid; }
    const Target* t = reinterpret_cast<const Target*>(th);
    if (n <= 0) return kOk;
    if (!q_dev || !idx_dev) { set_error("ficp_nn_query: null pointer"); return kErrInvalid; }
    if (ld < 2 || (use_z && ld < 3)) { set_error("ficp_nn_query: not enough columns"); return kErrInvalid; }
    if (use_z && !t->has_z) { set_error("ficp_nn_query: target was built without Z"); return kErrInvalid; }
    if (kernel < 0 || kernel > 2) { set_error("ficp_nn_query: kernel must be 0 (auto), 1 (thread per query) or 2 (bulk)"); return kErrInvalid; }
    cudaStream_t s = (cudaStream_t)stream;
    const bool bulk = (kernel == 2) || (kernel == 0 && nn_bulk_applies(t->view, n));
    if (bulk && n > 0x7FFFFFFFLL) { set_error("ficp_nn_query: bulk kernel takes at most 2^31 - 1 queries per call"); return kErrTooLarge; }
    int rc;
    if (bulk && counters_out) {
        DevBuf<unsigned long long> dc(s);
        if ((rc = dc.alloc(4))) return rc;
        FICP_CUDA(cudaMemsetAsync(dc.p, 0, sizeof(unsigned long long) * 4, s));
        rc = launch_nn_query_bulk(t->view, use_z != 0, q_dev, n, ld, idx_dev, dist_dev, nullptr, dc.p, s);
        t->used.record(s);
        if (rc) return rc;
        FICP_CUDA(cudaMemcpyAsync(counters_out, dc.p, sizeof(uint64_t) * 3, cudaMemcpyDeviceToHost, s));
        FICP_CUDA(cudaStreamSynchronize(s));
        return kOk;
    }
    if (counters_out) counters_out[0] = counters_out[1] = counters_out[2] = 0;
    rc = bulk ? launch_nn_query_bulk(t->view, use_z != 0, q_dev, n, ld, idx_dev, dist_dev, nullptr, nullptr, s)
              : launch_nn_query(t->view, use_z != 0, q_dev, n, ld, idx_dev, dist_dev, nullptr, s);
    t->used.record(s);
    return rc;
}

int ficp_nn_query_device(const ficp_target* th, const double* q_dev, int64_t n, int32_t ld, int32_t use_z,
                         int32_t* idx_dev, double* dist_dev, void* stream) {
    return ficp_nn_query_device_ex(th, q_dev, n, ld, use_z, idx_dev, dist_dev, 0, nullptr, stream);
}

int ficp_nn_query_ex(const ficp_target* th, const double* q_host, int64_t n, int32_t ld, int32_t use_z,
                     int64_t* idx_out, double* dist_out, int32_t kernel, uint64_t* counters_out, void* stream) {
    if (n <= 0) return kOk;
    if (!th || !q_host || !idx_out) { set_error("ficp_nn_query: null pointer"); return kErrInvalid; }
    if (ld < 2 || (use_z && ld < 3)) { set_error("ficp_nn_query: not enough columns"); return kErrInvalid; }
    for (int64_t i = 0; i < n; ++i) {
        const double* r = q_host + (size_t)i * ld;
        if (!std::isfinite(r[0]) || !std::isfinite(r[1]) || (use_z && !std::isfinite(r[2]))) {
            set_error("query contains non-finite coordinates ('x' must be finite)");
            return kErrNonFinite;
        }
    }
    cudaStream_t s = (cudaStream_t)stream;
    DevBuf<double> dq(s), dd(s);
    DevBuf<int> di(s);
    int rc;
    if ((rc = dq.alloc((size_t)n * ld)) || (rc = dd.alloc(n)) || (rc = di.alloc(n))) return rc;
    FICP_CUDA(cudaMemcpyAsync(dq.p, q_host, sizeof(double) * (size_t)n * ld, cudaMemcpyHostToDevice, s));
    if ((rc = ficp_nn_query_device_ex(th, dq.p, n, ld, use_z, di.p, dd.p, kernel, counters_out, stream))) return rc;
    std::vector<int> hi((size_t)n);
    FICP_CUDA(cudaMemcpyAsync(hi.data(), di.p, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, s));
    if (dist_out) FICP_CUDA(cudaMemcpyAsync(dist_out, dd.p, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, s));
    FICP_CUDA(cudaStreamSynchronize(s));
    for (int64_t i = 0; i < n; ++i) idx_out[i] = hi[(size_t)i];
    return kOk;
}

int ficp_nn_query(const ficp_target* th, const double* q_host, int64_t n, int32_t ld, int32_t use_z,
                  int64_t* idx_out, double* dist_out, void* stream) {
    return ficp_nn_query_ex(th, q_host, n, ld, use_z, idx_out, dist_out, 0, nullptr, stream);
}

// ------------------------------------------------------------------------------------------ match & remove
int ficp_match_remove(const ficp_target* th, const double* trees_host, const int64_t* offsets, int64_t n_plots,
                      int32_t ld, int32_t use_z, const double* thr_host, int64_t* matched_out, void* stream) {
    if (n_plots <= 0) return kOk;
    if (!th || !trees_host || !offsets || !thr_host || !matched_out) { set_error("ficp_match_remove: null pointer"); return kErrInvalid; }
    const Target* t = reinterpret_cast<const Target*>(th);
    if (ld < 2 || (use_z && (ld < 3 || !t->has_z))) { set_error("ficp_match_remove: Z requested but not available on both sides"); return kErrInvalid; }
    if (n_plots > 100000000) { set_error("ficp_match_remove: too many plots"); return kErrTooLarge; }
    const long long rows = offsets[n_plots];
    for (int64_t p = 0; p < n_plots; ++p)
        if (offsets[p + 1] < offsets[p]) { set_error("ficp_match_remove: offsets must be non-decreasing"); return kErrInvalid; }
    if (rows <= 0) return kOk;
    if (t->m <= 0) { for (long long i = 0; i < rows; ++i) matched_out[i] = -1; return kOk; }
    cudaStream_t s = (cudaStream_t)stream;
    DevBuf<double> dt(s), dthr(s);
    DevBuf<long long> doff(s), dout(s);
    DevBuf<int> dscr(s);
    int rc;
    if ((rc = dt.alloc((size_t)rows * ld)) || (rc = dthr.alloc(rows)) || (rc = doff.alloc(n_plots + 1)) || (rc = dout.alloc(rows)) ||
        (rc = dscr.alloc(rows)))
        return rc;
    static_assert(sizeof(long long) == sizeof(int64_t), "int64 layout");
    FICP_CUDA(cudaMemcpyAsync(dt.p, trees_host, sizeof(double) * (size_t)rows * ld, cudaMemcpyHostToDevice, s));
    FICP_CUDA(cudaMemcpyAsync(dthr.p, thr_host, sizeof(double) * (size_t)rows, cudaMemcpyHostToDevice, s));
    FICP_CUDA(cudaMemcpyAsync(doff.p, offsets, sizeof(int64_t) * (size_t)(n_plots + 1), cudaMemcpyHostToDevice, s));
    if ((rc = launch_match_remove(t->view, use_z != 0, dt.p, doff.p, (int)n_plots, ld, dthr.p, dout.p, dscr.p, s))) return rc;
    FICP_CUDA(cudaMemcpyAsync(matched_out, dout.p, sizeof(int64_t) * (size_t)rows, cudaMemcpyDeviceToHost, s));
    FICP_CUDA(cudaStreamSynchronize(s));
    return kOk;
}

// ------------------------------------------------------------------------------------------ radial crop
int ficp_radial_crop(const ficp_target* th, double cx, double cy, double dist, uint8_t* mask_out, void* stream) {
    if (!th || !mask_out) { set_error("ficp_radial_crop: null pointer"); return kErrInvalid; }
    const Target* t = reinterpret_cast<const Target*>(th);
    if (t->m <= 0) return kOk;
    if (!std::isfinite(cx) || !std::isfinite(cy) || std::isnan(dist)) { set_error("ficp_radial_crop: non-finite centre or radius"); return kErrNonFinite; }
    cudaStream_t s = (cudaStream_t)stream;
    DevBuf<unsigned char> dm(s);
    int rc;
    if ((rc = dm.alloc((size_t)t->m))) return rc;
    FICP_CUDA(cudaMemsetAsync(dm.p, 0, (size_t)t->m, s));
    if ((rc = launch_radial_crop(t->view, cx, cy, dist, dm.p, s))) return rc;
    FICP_CUDA(cudaMemcpyAsync(mask_out, dm.p, (size_t)t->m, cudaMemcpyDeviceToHost, s));
    FICP_CUDA(cudaStreamSynchronize(s));
    return kOk;
}

// ------------------------------------------------------------------------------------------ trimming
int ficp_select_fraction(const double* src_host, int32_t ld_s, const double* corr_host, int32_t ld_c,
                         const double* dist_host, int64_t n, int32_t md, const double* weights_host, int64_t fixed_k,
                         int64_t* k_out, double* frmsd_out, int64_t* order_out) {
    if (k_out) *k_out = 0;
    if (frmsd_out) *frmsd_out = HUGE_VAL;
    if (n <= 0) return kOk;
    if (!dist_host) { set_error("ficp_select_fraction: null distances"); return kErrInvalid; }
    if (n > kSelectLargeMaxN) { set_error("ficp_select_fraction: more than 2^24 points per plot are not supported"); return kErrTooLarge; }
    const bool want_k = (src_host != nullptr);
    if (want_k && (!corr_host || !weights_host || ld_s < md || ld_c < md || (md != 2 && md != 3))) {
        set_error("ficp_select_fraction: bad source/correspondence arguments");
        return kErrInvalid;
    }
    if (fixed_k < 0 || fixed_k > n) { set_error("ficp_select_fraction: fixed_k out of range"); return kErrInvalid; }
    DevBuf<double> ds, dc, dd, dw, df;
    DevBuf<long long> dk;
    DevBuf<int> dord;
    int rc;
    if ((rc = dd.alloc(n)) || (rc = dk.alloc(1)) || (rc = df.alloc(1)) || (rc = dord.alloc(n))) return rc;
    FICP_CUDA(cudaMemcpy(dd.p, dist_host, sizeof(double) * n, cudaMemcpyHostToDevice));
    if (want_k) {
        if ((rc = ds.alloc((size_t)n * ld_s)) || (rc = dc.alloc((size_t)n * ld_c)) || (rc = dw.alloc(n))) return rc;
        FICP_CUDA(cudaMemcpy(ds.p, src_host, sizeof(double) * n * ld_s, cudaMemcpyHostToDevice));
        FICP_CUDA(cudaMemcpy(dc.p, corr_host, sizeof(double) * n * ld_c, cudaMemcpyHostToDevice));
        FICP_CUDA(cudaMemcpy(dw.p, weights_host, sizeof(double) * n, cudaMemcpyHostToDevice));
    }
    rc = launch_select_fraction(want_k ? ds.p : nullptr, ld_s, dc.p, ld_c, dd.p, (int)n, md, dw.p, (int)fixed_k, dk.p,
                                df.p, order_out ? dord.p : nullptr, 0);
    if (rc) return rc;
    FICP_CUDA(cudaDeviceSynchronize());
    if (want_k) {
        long long k = 0;
        FICP_CUDA(cudaMemcpy(&k, dk.p, sizeof k, cudaMemcpyDeviceToHost));
        if (k_out) *k_out = k;
        if (frmsd_out) FICP_CUDA(cudaMemcpy(frmsd_out, df.p, sizeof(double), cudaMemcpyDeviceToHost));
    }
    if (order_out) {
        std::vector<int> ho((size_t)n);
        FICP_CUDA(cudaMemcpy(ho.data(), dord.p, sizeof(int) * n, cudaMemcpyDeviceToHost));
        for (int64_t i = 0; i < n; ++i) order_out[i] = ho[(size_t)i];
    }
    return kOk;
}

// ------------------------------------------------------------------------------------------ fit / apply
int ficp_fit_rigid2d(const double* src_host, int32_t ld_s, const double* tgt_host, int32_t ld_t, int64_t k,
                     int32_t allow_reflection, double* T9) {
    if (!src_host || !tgt_host || !T9 || k <= 0 || ld_s < 2 || ld_t < 2 || k > 100000000LL) {
        set_error("ficp_fit_rigid2d: bad arguments (need k >= 1 rows with >= 2 columns)");
        return kErrInvalid;
    }
    DevBuf<double> ds, dt, dT;
    int rc;
    if ((rc = ds.alloc((size_t)k * ld_s)) || (rc = dt.alloc((size_t)k * ld_t)) || (rc = dT.alloc(9))) return rc;
    FICP_CUDA(cudaMemcpy(ds.p, src_host, sizeof(double) * k * ld_s, cudaMemcpyHostToDevice));
    FICP_CUDA(cudaMemcpy(dt.p, tgt_host, sizeof(double) * k * ld_t, cudaMemcpyHostToDevice));
    if ((rc = launch_fit_rigid2d(ds.p, ld_s, dt.p, ld_t, nullptr, (int)k, allow_reflection, dT.p, 0))) return rc;
    FICP_CUDA(cudaMemcpy(T9, dT.p, sizeof(double) * 9, cudaMemcpyDeviceToHost));
    return kOk;
}

int ficp_apply_xy(const double* in_host, double* out_host, int64_t n, int32_t ld, const double* T9) {
    if (n <= 0) return kOk;
    if (!in_host || !out_host || !T9 || ld < 2) { set_error("ficp_apply_xy: bad arguments"); return kErrInvalid; }
    DevBuf<double> di, dout, dT;
    int rc;
    if ((rc = di.alloc((size_t)n * ld)) || (rc = dout.alloc((size_t)n * ld)) || (rc = dT.alloc(9))) return rc;
    FICP_CUDA(cudaMemcpy(di.p, in_host, sizeof(double) * n * ld, cudaMemcpyHostToDevice));
    FICP_CUDA(cudaMemcpy(dT.p, T9, sizeof(double) * 9, cudaMemcpyHostToDevice));
    if ((rc = launch_apply_xy(di.p, dout.p, n, ld, dT.p, 0))) return rc;
    FICP_CUDA(cudaMemcpy(out_host, dout.p, sizeof(double) * n * ld, cudaMemcpyDeviceToHost));
    return kOk;
}

int ficp_sumsq(const double* a_host, int32_t ld_a, const double* b_host, int32_t ld_b, int64_t k, int32_t md,
               double* out) {
    if (!out) { set_error("ficp_sumsq: null output"); return kErrInvalid; }
    *out = 0.0;
    if (k <= 0) return kOk;
    if (!a_host || !b_host || ld_a < md || ld_b < md || (md != 2 && md != 3)) { set_error("ficp_sumsq: bad arguments"); return kErrInvalid; }
    DevBuf<double> da, db, dout;
    int rc;
    if ((rc = da.alloc((size_t)k * ld_a)) || (rc = db.alloc((size_t)k * ld_b)) || (rc = dout.alloc(1))) return rc;
    FICP_CUDA(cudaMemcpy(da.p, a_host, sizeof(double) * k * ld_a, cudaMemcpyHostToDevice));
    FICP_CUDA(cudaMemcpy(db.p, b_host, sizeof(double) * k * ld_b, cudaMemcpyHostToDevice));
    if ((rc = launch_sumsq(da.p, ld_a, db.p, ld_b, nullptr, (int)k, md, dout.p, 0))) return rc;
    FICP_CUDA(cudaMemcpy(out, dout.p, sizeof(double), cudaMemcpyDeviceToHost));
    return kOk;
}

// ------------------------------------------------------------------------------------------ stepper
// Device-resident state of ONE plot above the persistent kernels' 1024-tree limit: the host drives the loop of
// ficp.py:122-147 pass by pass (it owns the convergence test, in the reference's own expressions), the arrays stay here.
// Every step runs exactly the stage kernels the host-buffer entry points above run, on the same values in the same order -
// same bits, without their six host round trips per pass.
struct Stepper {
    const Target* t = nullptr;
    long long n = 0, m = 0;
    int md = 2;
    long long k_last = 0;
    double *d_src = nullptr, *d_src2 = nullptr, *d_corr = nullptr, *d_dist = nullptr, *d_w = nullptr;
    double *d_f = nullptr, *d_sum = nullptr, *d_T = nullptr;
    int *d_idx = nullptr, *d_order = nullptr, *d_inv = nullptr;   // d_inv: grid position of every original target row
    long long* d_k = nullptr;
    ~Stepper() {
        cudaDeviceSynchronize();
        void* all[] = {d_inv, d_src, d_src2, d_corr, d_dist, d_w, d_f, d_sum, d_T, d_idx, d_order, d_k};
        for (void* q : all) dev_free(q, 0);
    }
};

int ficp_stepper_create(const ficp_target* th, const double* src_host, int64_t n, int32_t ld_s, int32_t md, ficp_stepper** out) {
    if (out) *out = nullptr;
    if (!th || !src_host || !out) { set_error("ficp_stepper_create: null pointer"); return kErrInvalid; }
    const Target* t = reinterpret_cast<const Target*>(th);
    if ((md != 2 && md != 3) || ld_s < md || n <= 0 || n > kSelectLargeMaxN || t->m <= 0 || t->m > 0x7FFFFFFFLL) {
        set_error("ficp_stepper_create: bad shape (need md in {2, 3} columns, 1 <= n <= 2^24, a non-empty target)");
        return kErrInvalid;
    }
    if ((md == 3) != (t->has_z != 0)) { set_error("ficp_stepper_create: the index must be built on the md matched columns"); return kErrInvalid; }
    Stepper* S = new Stepper();
    S->t = t; S->n = n; S->m = t->m; S->md = md;
    cudaError_t e = cudaSuccess;
    auto A = [&](auto** q, size_t count) { if (e == cudaSuccess) e = dev_alloc_t(q, count, 0); };
    A(&S->d_inv, (size_t)t->m); A(&S->d_src, (size_t)n * md); A(&S->d_src2, (size_t)n * md); A(&S->d_corr, (size_t)n * md);
    A(&S->d_dist, (size_t)n); A(&S->d_w, (size_t)n); A(&S->d_f, 1); A(&S->d_sum, 1); A(&S->d_T, 9);
    A(&S->d_idx, (size_t)n); A(&S->d_order, (size_t)n); A(&S->d_k, 1);
    if (e == cudaSuccess) {
        if (ld_s == md) {
            e = cudaMemcpy(S->d_src, src_host, sizeof(double) * (size_t)n * md, cudaMemcpyHostToDevice);
        } else {   // pack the matched columns first: a strided copy of short rows is one transfer per row
            std::vector<double> packed((size_t)n * md);
            for (int64_t i = 0; i < n; ++i)
                for (int c = 0; c < md; ++c) packed[(size_t)i * md + c] = src_host[(size_t)i * ld_s + c];
            e = cudaMemcpy(S->d_src, packed.data(), sizeof(double) * packed.size(), cudaMemcpyHostToDevice);
        }
    }
    if (e != cudaSuccess) { delete S; return cuda_fail(e, "ficp_stepper_create", __FILE__, __LINE__); }
    const int rc = launch_inverse_perm(t->view, S->d_inv, 0);
    t->used.record(0);
    if (rc) { delete S; return rc; }
    *out = reinterpret_cast<ficp_stepper*>(S);
    return kOk;
}

int ficp_stepper_set_weights(ficp_stepper* sh, const double* weights_host) {
    if (!sh || !weights_host) { set_error("ficp_stepper_set_weights: null pointer"); return kErrInvalid; }
    Stepper* S = reinterpret_cast<Stepper*>(sh);
    FICP_CUDA(cudaMemcpy(S->d_w, weights_host, sizeof(double) * S->n, cudaMemcpyHostToDevice));
    return kOk;
}

// One pass (ficp.py:133-135 / :123-124): nearest neighbours of the current positions, their rows, the trim order and the
// FRMSD-optimal k (or fixed_k), and the sum of squared residuals over the first k rows of that order.
int ficp_stepper_pass(ficp_stepper* sh, int64_t fixed_k, int64_t* k_out, double* sumsq_out) {
    if (!sh || !k_out || !sumsq_out) { set_error("ficp_stepper_pass: null pointer"); return kErrInvalid; }
    Stepper* S = reinterpret_cast<Stepper*>(sh);
    *k_out = 0; *sumsq_out = 0.0;
    if (fixed_k < 0 || fixed_k > S->n) { set_error("ficp_stepper_pass: fixed_k out of range"); return kErrInvalid; }
    const bool z3 = S->md == 3;
    int rc = nn_bulk_applies(S->t->view, S->n)
                 ? launch_nn_query_bulk(S->t->view, z3, S->d_src, S->n, S->md, S->d_idx, S->d_dist, nullptr, nullptr, 0)
                 : launch_nn_query(S->t->view, z3, S->d_src, S->n, S->md, S->d_idx, S->d_dist, nullptr, 0);
    S->t->used.record(0);
    if (rc) return rc;
    if ((rc = launch_gather_grid_rows(S->t->view, z3, S->d_inv, S->d_idx, S->n, S->d_corr, 0))) return rc;
    if ((rc = launch_select_fraction(S->d_src, S->md, S->d_corr, S->md, S->d_dist, (int)S->n, S->md, S->d_w, (int)fixed_k, S->d_k,
                                     S->d_f, S->d_order, 0))) return rc;
    long long k = 0;
    FICP_CUDA(cudaMemcpy(&k, S->d_k, sizeof k, cudaMemcpyDeviceToHost));
    S->k_last = k;
    *k_out = k;
    if (k > 0) {
        if ((rc = launch_sumsq(S->d_src, S->md, S->d_corr, S->md, S->d_order, (int)k, S->md, S->d_sum, 0))) return rc;
        FICP_CUDA(cudaMemcpy(sumsq_out, S->d_sum, sizeof(double), cudaMemcpyDeviceToHost));
    }
    return kOk;
}

// Fit on the trimmed rows of the last pass and move the plot (ficp.py:137-139); T9 = the step's 3x3 transform.
int ficp_stepper_fit_apply(ficp_stepper* sh, int32_t allow_reflection, double* T9) {
    if (!sh || !T9) { set_error("ficp_stepper_fit_apply: null pointer"); return kErrInvalid; }
    Stepper* S = reinterpret_cast<Stepper*>(sh);
    if (S->k_last <= 0) { set_error("ficp_stepper_fit_apply: the last pass kept no rows"); return kErrInvalid; }
    int rc;
    if ((rc = launch_fit_rigid2d(S->d_src, S->md, S->d_corr, S->md, S->d_order, (int)S->k_last, allow_reflection, S->d_T, 0))) return rc;
    if ((rc = launch_apply_xy(S->d_src, S->d_src2, S->n, S->md, S->d_T, 0))) return rc;
    std::swap(S->d_src, S->d_src2);
    FICP_CUDA(cudaMemcpy(T9, S->d_T, sizeof(double) * 9, cudaMemcpyDeviceToHost));
    return kOk;
}

int ficp_stepper_read_xy(ficp_stepper* sh, double* xy_out) {
    if (!sh || !xy_out) { set_error("ficp_stepper_read_xy: null pointer"); return kErrInvalid; }
    Stepper* S = reinterpret_cast<Stepper*>(sh);
    if (S->md == 2) {
        FICP_CUDA(cudaMemcpy(xy_out, S->d_src, sizeof(double) * 2 * (size_t)S->n, cudaMemcpyDeviceToHost));
        return kOk;
    }
    std::vector<double> rows((size_t)S->n * S->md);
    FICP_CUDA(cudaMemcpy(rows.data(), S->d_src, sizeof(double) * rows.size(), cudaMemcpyDeviceToHost));
    for (long long i = 0; i < S->n; ++i) { xy_out[2 * i] = rows[(size_t)i * S->md]; xy_out[2 * i + 1] = rows[(size_t)i * S->md + 1]; }
    return kOk;
}

void ficp_stepper_destroy(ficp_stepper* sh) { delete reinterpret_cast<Stepper*>(sh); }

// ------------------------------------------------------------------------------------------ host-side plot geometry
static int check_offsets(const char* who, const int64_t* plot_offsets, int64_t n_plots) {
    if (n_plots <= 0) { set_error(std::string(who) + ": need n_plots >= 1"); return kErrInvalid; }
    for (int64_t p = 0; p < n_plots; ++p)
        if (plot_offsets[p + 1] <= plot_offsets[p] || plot_offsets[p] < 0) { set_error(std::string(who) + ": empty plot or unsorted offsets"); return kErrInvalid; }
    return kOk;
}

int ficp_plot_centres(const double* src_host, int32_t ld, const int64_t* plot_offsets, int64_t n_plots, double* centres_out) {
    if (!src_host || !plot_offsets || !centres_out || ld < 2) { set_error("ficp_plot_centres: null pointer or ld < 2"); return kErrInvalid; }
    if (int rc = check_offsets("ficp_plot_centres", plot_offsets, n_plots)) return rc;
    plot_centres_host(src_host, ld, plot_offsets, n_plots, centres_out, host_threads_for(plot_offsets[n_plots] - plot_offsets[0]));
    return kOk;
}

int ficp_plot_geometry(const double* src_host, int32_t ld, int32_t use_z, const int64_t* plot_offsets, int64_t n_plots,
                       const double* centres, double* u_out, double* z_out, double* ubar_out, double* rho_out) {
    if (!src_host || !plot_offsets || !centres || !ubar_out || !rho_out || (u_out && use_z && !z_out) || ld < 2 || (use_z && ld < 3)) {
        set_error("ficp_plot_geometry: null pointer or too few columns");
        return kErrInvalid;
    }
    if (int rc = check_offsets("ficp_plot_geometry", plot_offsets, n_plots)) return rc;
    if (plot_offsets[0] != 0) { set_error("ficp_plot_geometry: offsets must start at 0"); return kErrInvalid; }
    if (!plot_geometry_host(src_host, ld, use_z != 0, plot_offsets, n_plots, centres, u_out, z_out, ubar_out, rho_out,
                            host_threads_for(plot_offsets[n_plots]))) {
        set_error("source contains non-finite coordinates ('x' must be finite)");
        return kErrNonFinite;
    }
    return kOk;
}

// ------------------------------------------------------------------------------------------ batch
int ficp_batch_create(const ficp_target* th, const double* src_host, int32_t ld, int32_t use_z,
                      const int64_t* plot_offsets, int64_t n_plots, const double* centres, const double* hyp,
                      int64_t n_hyp, int32_t hyp_begin, int32_t hyp_stride, const double* weights,
                      const int64_t* weight_offsets, const int32_t* plot_tab, int32_t n_tabs, const int32_t* fixed_k,
                      const ficp_batch_params* prm, int32_t want_final_xy, void* stream, ficp_batch** out) {
    if (out) *out = nullptr;
    if (!th || !src_host || !plot_offsets || !hyp || !weights || !weight_offsets || !plot_tab || !prm || !out) {
        set_error("ficp_batch_create: null pointer");
        return kErrInvalid;
    }
    const Target* t = reinterpret_cast<const Target*>(th);
    if (t->m <= 0) { set_error("ficp_batch_create: empty target (callers return the source unchanged, ficp.py:66-68)"); return kErrInvalid; }
    if (n_plots <= 0 || n_hyp <= 0 || hyp_stride <= 0 || hyp_begin < 0 || hyp_begin >= hyp_stride + n_hyp) {
        set_error("ficp_batch_create: need n_plots >= 1, n_hyp >= 1, hyp_stride >= 1");
        return kErrInvalid;
    }
    if (ld < 2 || (use_z && (ld < 3 || !t->has_z))) { set_error("ficp_batch_create: Z requested but not available on both sides"); return kErrInvalid; }
    if (prm->n_stages < 1 || prm->n_stages > kMaxStages) { set_error("ficp_batch_create: n_stages must be 1 or 2"); return kErrInvalid; }
    if (n_plots > 50000000 || n_hyp > 100000000) { set_error("ficp_batch_create: batch too large"); return kErrTooLarge; }
    cudaStream_t s = (cudaStream_t)stream;
    const GridGeom& g = t->view.g;

    const long long rows = plot_offsets[n_plots];
    int max_n = 0;
    for (int64_t p = 0; p < n_plots; ++p) {
        const long long n = plot_offsets[p + 1] - plot_offsets[p];
        if (n <= 0) { set_error("ficp_batch_create: empty plot (callers return an empty plot unchanged)"); return kErrInvalid; }
        if (n > 1024) { set_error("ficp_batch_create: more than 1024 trees per plot are not supported by the persistent kernel"); return kErrTooLarge; }
        if (plot_tab[p] < 0 || plot_tab[p] >= n_tabs) { set_error("ficp_batch_create: plot_tab out of range"); return kErrInvalid; }
        if (weight_offsets[plot_tab[p] + 1] - weight_offsets[plot_tab[p]] != (long long)prm->n_stages * n) {
            set_error("ficp_batch_create: weight table length does not match n_stages * plot size");
            return kErrInvalid;
        }
        if (fixed_k && (fixed_k[p] < 0 || fixed_k[p] > n)) { set_error("ficp_batch_create: fixed_k out of range"); return kErrInvalid; }
        max_n = std::max<int>(max_n, (int)n);
    }
    const int e = pick_e(max_n);
    const int npad = 32 * e;
    const bool z3 = use_z != 0;
    const int n_hyp_local = (hyp_begin < n_hyp) ? (int)((n_hyp - hyp_begin + hyp_stride - 1) / hyp_stride) : 0;
    if (n_hyp_local <= 0) { set_error("ficp_batch_create: this shard owns no hypothesis"); return kErrInvalid; }

    Batch* b = new Batch();
    struct Guard { Batch* b; bool armed = true; ~Guard() { if (armed) delete b; } } guard{b};
    b->alloc_stream = s;
    b->tgt = t; b->n_plots = (int)n_plots; b->n_hyp = (int)n_hyp; b->n_hyp_local = n_hyp_local;
    b->rows = rows; b->z3 = z3; b->want_final = (want_final_xy != 0) && n_hyp_local == 1;

    // ---- hypothesis translation range (for the window footprint)
    double dxmin = HUGE_VAL, dxmax = -HUGE_VAL, dymin = HUGE_VAL, dymax = -HUGE_VAL;
    for (int64_t h = 0; h < n_hyp; ++h) {
        for (int c = 0; c < 6; ++c)
            if (!std::isfinite(hyp[h * 6 + c])) { set_error("hypothesis table contains non-finite values"); return kErrNonFinite; }
        dxmin = std::min(dxmin, hyp[h * 6 + 4]); dxmax = std::max(dxmax, hyp[h * 6 + 4]);
        dymin = std::min(dymin, hyp[h * 6 + 5]); dymax = std::max(dymax, hyp[h * 6 + 5]);
    }

    // ---- per-plot geometry: local coordinates u = p - centre, shift point, footprint of the start poses
    // (batch_prep.h: a few host threads over the plots - with one ICP per plot this pass is the end-to-end step)
    std::vector<PlotMeta> plots((size_t)n_plots);
    // Two routes for the rows.  (a) The caller's array is page-locked (a torch pinned tensor, cudaHostRegister'ed memory) and
    // narrow: it goes to the device AS IT IS - one DMA, started before the host pass below and running under it - and a small
    // kernel splits it into u = (x, y) - centre and z (launch_split_rows: the same single subtraction, the same bits); the host
    // only READS the rows once (centres, shift points, radii, finiteness).  (b) Otherwise the host writes u and z into a
    // page-locked block of the library and that block is copied, in up to four slices of plots, each as soon as it is written.
    // FICP_HOST_STAGING=1 forces (b) (tests compare the two).
    bool direct = false;
    if (ld <= 4 && !std::getenv("FICP_HOST_STAGING")) {
        cudaPointerAttributes at{};
        if (cudaPointerGetAttributes(&at, src_host) == cudaSuccess) direct = (at.type == cudaMemoryTypeHost);
        else cudaGetLastError();
    }
    const size_t per_row = direct ? 0 : (z3 ? 3 : 2);
    HostStage stage(sizeof(double) * ((size_t)rows * per_row + 3 * (size_t)n_plots));
    double* const h_u = direct ? nullptr : static_cast<double*>(stage.p);   // (ux, uy) per row = the device's double2
    double* const h_z = (!direct && z3) ? h_u + 2 * (size_t)rows : nullptr;
    double* const h_ubar = static_cast<double*>(stage.p) + (size_t)rows * per_row;
    double* const h_rho = h_ubar + 2 * (size_t)n_plots;
    FICP_CUDA(dev_alloc(reinterpret_cast<void**>(&b->d_src_u), sizeof(double2) * (size_t)rows, s));
    if (z3) FICP_CUDA(dev_alloc(reinterpret_cast<void**>(&b->d_src_z), sizeof(double) * (size_t)rows, s));
    DevBuf<double> d_raw(s);
    // centres == NULL: every plot turns about the in-order mean of its first two columns (ficp_plot_centres: the bits of
    // `rows[:, :2].mean(axis=0)`), taken here right before the geometry pass reads the same rows
    std::vector<double> own_centres(centres ? 0 : 2 * (size_t)n_plots);
    if (!centres) centres = own_centres.data();
    if (direct) {
        if (int rc = d_raw.alloc((size_t)rows * ld)) return rc;
        stage.pending_on = s; stage.pending = true;   // from here on the stream reads the CALLER's array: never return before it is done
        FICP_CUDA(cudaMemcpyAsync(d_raw.p, src_host, sizeof(double) * (size_t)rows * ld, cudaMemcpyHostToDevice, s));
        if (!own_centres.empty()) plot_centres_host(src_host, ld, plot_offsets, n_plots, own_centres.data(), host_threads_for(rows));
        if (!plot_geometry_host(src_host, ld, z3, plot_offsets, n_plots, centres, nullptr, nullptr, h_ubar, h_rho, host_threads_for(rows))) {
            set_error("source contains non-finite coordinates ('x' must be finite)");
            return kErrNonFinite;
        }
    } else {
        const int n_slices_up = (stage.pinned && rows >= 65536) ? (int)std::min<int64_t>(4, n_plots) : 1;
        int64_t p0 = 0;
        for (int sl = 1; sl <= n_slices_up; ++sl) {
            int64_t p1 = n_plots;
            if (sl < n_slices_up) {
                p1 = std::lower_bound(plot_offsets, plot_offsets + n_plots, rows * sl / n_slices_up) - plot_offsets;
                p1 = std::max(p0, std::min(p1, n_plots));
            }
            if (p1 == p0) continue;
            const long long r0 = plot_offsets[p0], r1 = plot_offsets[p1];
            if (!own_centres.empty())
                plot_centres_host(src_host, ld, plot_offsets + p0, p1 - p0, own_centres.data() + 2 * p0, host_threads_for(r1 - r0));
            if (!plot_geometry_host(src_host, ld, z3, plot_offsets + p0, p1 - p0, centres + 2 * p0, h_u, h_z, h_ubar + 2 * p0, h_rho + p0,
                                    host_threads_for(r1 - r0))) {
                set_error("source contains non-finite coordinates ('x' must be finite)");
                return kErrNonFinite;
            }
            stage.pending_on = s; stage.pending = true;
            FICP_CUDA(cudaMemcpyAsync(b->d_src_u + r0, h_u + 2 * (size_t)r0, sizeof(double2) * (size_t)(r1 - r0), cudaMemcpyHostToDevice, s));
            if (z3) FICP_CUDA(cudaMemcpyAsync(b->d_src_z + r0, h_z + (size_t)r0, sizeof(double) * (size_t)(r1 - r0), cudaMemcpyHostToDevice, s));
            p0 = p1;
        }
    }
    b->rows_direct = direct;
    struct Foot { double rho, fx0, fx1, fy0, fy1; };
    std::vector<Foot> foot((size_t)n_plots);
    for (int64_t p = 0; p < n_plots; ++p) {
        PlotMeta& pm = plots[(size_t)p];
        pm.off = plot_offsets[p];
        pm.n = (int)(plot_offsets[p + 1] - plot_offsets[p]);
        pm.tab = plot_tab[p];
        pm.cinx = centres[2 * p]; pm.ciny = centres[2 * p + 1];
        pm.fixed_k = fixed_k ? fixed_k[p] : 0;
        pm.pad = 0;
        pm.wx0 = pm.wy0 = pm.wx1 = pm.wy1 = 0;
        pm.ubx = h_ubar[2 * (size_t)p]; pm.uby = h_ubar[2 * (size_t)p + 1];
        Foot& f = foot[(size_t)p];
        f.rho = h_rho[(size_t)p];
        // footprint of the plot centroid over all start poses: centre_h = M_h ubar + cin + d_h
        if (std::hypot(pm.ubx, pm.uby) <= 1e-9 * (f.rho + 1.0)) {
            f.fx0 = pm.cinx + dxmin; f.fx1 = pm.cinx + dxmax; f.fy0 = pm.ciny + dymin; f.fy1 = pm.ciny + dymax;
        } else {
            f.fx0 = f.fy0 = HUGE_VAL; f.fx1 = f.fy1 = -HUGE_VAL;
            for (int64_t h = 0; h < n_hyp; ++h) {
                const double* hr = hyp + h * 6;
                const double cx = hr[0] * pm.ubx + hr[1] * pm.uby + pm.cinx + hr[4];
                const double cy = hr[2] * pm.ubx + hr[3] * pm.uby + pm.ciny + hr[5];
                f.fx0 = std::min(f.fx0, cx); f.fx1 = std::max(f.fx1, cx); f.fy0 = std::min(f.fy0, cy); f.fy1 = std::max(f.fy1, cy);
            }
        }
    }
    // window of grid cells a plot can reach with `margin` of slack: cell rectangle + estimated point count
    const double mean_cell_pts = (double)t->m / ((double)g.gw * g.gh);
    const double margin0 = prm->window_margin >= 0.0 ? prm->window_margin : std::min(3.0 * g.h + 5.0, g.h + 20.0);
    auto window_of = [&](const Foot& f, double margin, int* rect, long long* cells, double* est) -> bool {
        const double r = f.rho + margin;
        const double bx0 = f.fx0 - r, bx1 = f.fx1 + r, by0 = f.fy0 - r, by1 = f.fy1 + r;
        if (bx1 < g.x0 || by1 < g.y0 || bx0 > g.x0 + g.gw * g.h || by0 > g.y0 + g.gh * g.h) return false;  // off the map
        rect[0] = clamp_cell((bx0 - g.x0) * g.inv_h, g.gw); rect[1] = clamp_cell((bx1 - g.x0) * g.inv_h, g.gw) + 1;
        rect[2] = clamp_cell((by0 - g.y0) * g.inv_h, g.gh); rect[3] = clamp_cell((by1 - g.y0) * g.inv_h, g.gh) + 1;
        *cells = (long long)(rect[1] - rect[0]) * (rect[3] - rect[2]);
        *est = (double)*cells * mean_cell_pts * 1.15 + 96;
        return true;
    };

    // ---- launch shape: warps per CTA, CTAs per SM, shared-memory budget, window capacity
    int dev = 0, sms = 0, smem_optin = 0;
    FICP_CUDA(cudaGetDevice(&dev));
    FICP_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    FICP_CUDA(cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    int warps = prm->warps_per_cta > 0 ? prm->warps_per_cta : std::min(icp_max_warps(e), 16);
    warps = std::max(1, std::min(warps, icp_max_warps(e)));
    if (prm->no_helpers < 0 || prm->no_helpers > 2) { set_error("ficp_batch_create: no_helpers must be 0 (auto), 1 (off) or 2 (on)"); return kErrInvalid; }
    if (prm->cta_per_icp < 0 || prm->cta_per_icp > 2) { set_error("ficp_batch_create: cta_per_icp must be 0 (auto), 1 (off) or 2 (on)"); return kErrInvalid; }
    const long long n_icps_all = (long long)n_plots * n_hyp_local;
    // Kernel shape.  One WARP per ICP (icp_persistent.cu) has the higher throughput when the batch fills the machine; a
    // batch smaller than that is bounded by the serial chain of its longest ICP, and the CTA-per-ICP kernel (icp_team.cu:
    // every phase of a pass cooperative across 32 e threads) runs that chain several times faster.  Plots of <= 32
    // trees (e == 1) are one warp either way.  Auto: below kCtaAutoIcpsPerSm ICPs per SM (measured, profiles/r02_summary.md).
    // One start pose per plot (config 4) leaves the warp-per-ICP kernel one warp per CTA: the CTA-per-ICP kernel is ahead a
    // little further there (measured at 8.4 ICPs per SM: 40.7 vs 37.6 M hyp-iter/s, profiles/r02_bench_c4_*.json).
    // (round 2, after the latency work on the CTA kernel: 2048 ICPs = 13.8 per SM 3.06 vs 3.26 ms, 1366 ICPs 2.31 vs 3.18 ms,
    // 4096 ICPs = 27.7 per SM 5.31 vs 4.63 ms - profiles/r02_strong_scaling_probe_crossover.jsonl)
    const long long kCtaAutoIcpsPerSm = 14;
    const bool cta_mode = (e >= 2) && (prm->cta_per_icp == 2 || (prm->cta_per_icp == 0 && prm->team_warps == 0 && prm->no_helpers == 0 &&
                                                                 prm->warps_per_cta == 0 && n_icps_all <= kCtaAutoIcpsPerSm * sms));
    // Elastic kernel: warps without an ICP of their own help the ICPs in flight in their CTA.  It pays while the
    // tail (the last ICP of every warp) is a visible share of the launch - measured: 1.7 ICPs per warp slot +8..19 %,
    // 28 per slot -2 % (its one-warp-per-ICP path compiles slightly worse) - so by default it is used below 10 ICPs per
    // warp slot.  `team` = warps per ICP at launch: one, unless the batch is smaller than the machine (fewer ICPs
    // than warp slots) - then each CTA starts with fewer leads than warps and the rest help from the first pass on.
    // Results are bit-identical in every mode.
    const bool elastic = !cta_mode && (prm->no_helpers == 2 || (prm->no_helpers == 0 && (prm->team_warps > 1 || n_icps_all < 10LL * sms * 16)));
    int team = prm->team_warps;
    if (team != 0 && team != 1 && team != 2 && team != 4 && team != 8) { set_error("ficp_batch_create: team_warps must be 0 (auto), 1, 2, 4 or 8"); return kErrInvalid; }
    if (!cta_mode && !elastic && team > 1) { set_error("ficp_batch_create: team_warps > 1 needs the elastic kernel (no_helpers = 0)"); return kErrInvalid; }
    if (team == 0) {
        team = 1;
        const long long n_icps = (long long)n_plots * n_hyp_local, slots16 = (long long)sms * 16;
        while (elastic && team < 8 && n_icps * team * 2 <= slots16) team *= 2;
    }
    while (team > 1 && (team > e || team > warps)) team >>= 1;  // at least one round per warp of the team
    int slots_per_cta = std::max(1, std::min(warps / team, n_hyp_local));
    const size_t sm_total = 228 * 1024;  // per-SM shared memory; each resident CTA also reserves 1 KB
    const int wcap_rows = kWindowRowsCap;
    // per-ICP state (distances, neighbours, search list, trim order, slack: 18 B per tree) must leave room for a window
    while (!cta_mode && slots_per_cta > 1 && icp_smem_bytes(e, z3, slots_per_cta, 0, 0, wcap_rows) + 16384 > (size_t)smem_optin) --slots_per_cta;
    warps = slots_per_cta * team;
    if (cta_mode) { warps = icp_team_threads(e) / 32; team = warps; slots_per_cta = 1; }
    auto smem_of = [&](int wp, int wc) -> size_t {
        return cta_mode ? icp_team_smem_bytes(e, z3, wp, wc, wcap_rows) : icp_smem_bytes(e, z3, slots_per_cta, wp, wc, wcap_rows);
    };
    const size_t fixed_bytes = smem_of(0, 0);
    const size_t per_pt = 16 + (z3 ? 8 : 0), per_cell = 4;
    // bytes left for the window (point records + cell table) with `ctas` CTAs resident per SM
    auto window_bytes = [&](int ctas) -> size_t {
        const size_t budget = std::min<size_t>((size_t)smem_optin, sm_total / ctas - 1024);
        if (prm->disable_window || budget <= fixed_bytes + 4096) return 0;
        return budget - fixed_bytes - 256;
    };
    auto fits = [&](double est, long long cells, size_t bytes) -> bool {
        return est <= 32767.0 && (size_t)std::ceil(est) * per_pt + (size_t)cells * per_cell + 64 <= bytes;
    };
    // a window with a modest margin (one ring of cells + a few metres of drift) must fit for the worst plot
    int want_ctas_per_sm = prm->ctas_per_sm > 0 ? prm->ctas_per_sm
                           : cta_mode ? std::max(1, std::min(16, 32 / e))       // as many CTAs as the kernel's register budget allows (occupancy query below)
                                      : std::max(1, std::min(8, 16 / warps));
    if (prm->ctas_per_sm <= 0) {  // fewer resident CTAs when that is what it takes for the windows to fit on-chip
        auto modest_fits = [&](int ctas) -> bool {
            const size_t bytes = window_bytes(ctas);
            for (int64_t p = 0; p < n_plots; ++p) {
                int rect[4]; long long cells; double est;
                if (window_of(foot[(size_t)p], std::min(margin0, 1.5 * g.h + 3.0), rect, &cells, &est) && !fits(est, cells, bytes)) return false;
            }
            return true;
        };
        while (want_ctas_per_sm > 1 && !modest_fits(want_ctas_per_sm)) --want_ctas_per_sm;
    }
    const size_t wbytes = window_bytes(want_ctas_per_sm);

    // ---- window rectangles: the largest margin (up to margin0) that fits the capacity
    int wcap_pts = 0;
    long long need_cells = 0;
    for (int64_t p = 0; p < n_plots && wbytes > 0; ++p) {
        PlotMeta& pm = plots[(size_t)p];
        double margin = margin0;
        for (int attempt = 0; attempt < 12; ++attempt, margin *= 0.7) {
            int rect[4]; long long cells; double est;
            if (!window_of(foot[(size_t)p], margin, rect, &cells, &est)) break;
            // the largest window of any plot sets the reservation: a plot must fit next to the others' maxima
            const double est_all = std::max(est, (double)wcap_pts);
            const long long cells_all = std::max(cells, need_cells);
            if ((rect[3] - rect[2]) <= wcap_rows && fits(est_all, cells_all, wbytes)) {
                pm.wx0 = rect[0]; pm.wx1 = rect[1]; pm.wy0 = rect[2]; pm.wy1 = rect[3];
                wcap_pts = (int)std::ceil(est_all);
                need_cells = cells_all;
                break;
            }
        }
    }
    const int wcap_cells = (int)need_cells;
    const size_t smem = smem_of(wcap_pts, wcap_cells);
    if (smem > (size_t)smem_optin) { set_error("ficp_batch_create: shared-memory plan exceeds the device limit"); return kErrTooLarge; }
    int occ = 0;
    int rc = cta_mode ? icp_team_max_ctas_per_sm(e, z3, smem, &occ) : icp_max_ctas_per_sm(e, z3, warps, elastic, smem, &occ);
    if (rc) return rc;
    if (occ < 1) { set_error("ficp_batch_create: kernel does not fit on an SM with this configuration"); return kErrTooLarge; }
    const int ctas_per_sm = prm->ctas_per_sm > 0 ? std::min(occ, prm->ctas_per_sm) : occ;
    const long long resident = (long long)sms * ctas_per_sm;
    // tickets: each plot can be worked on by up to `slices_per_plot` CTAs at once (all of them when there are
    // few plots); tickets are dealt round-robin over the plots by the kernel
    const int max_slices_per_plot = (n_hyp_local + slots_per_cta - 1) / slots_per_cta;
    int slices_per_plot = (int)std::min<long long>(max_slices_per_plot, resident);
    const long long n_slices = (long long)n_plots * slices_per_plot;
    if (n_slices > 2000000000LL) { set_error("ficp_batch_create: too many work slices"); return kErrTooLarge; }

    // ---- FRMSD weight tables: [table][stage][g | c][NPAD], lane-permuted (k-1 = lane*E + r  ->  r*32 + lane)
    std::vector<double> h_tabs((size_t)n_tabs * prm->n_stages * 2 * npad, HUGE_VAL);
    for (int tb = 0; tb < n_tabs; ++tb) {
        const long long len = weight_offsets[tb + 1] - weight_offsets[tb];
        if (len % prm->n_stages) { set_error("ficp_batch_create: malformed weight table"); return kErrInvalid; }
        const int nt = (int)(len / prm->n_stages);
        if (nt > npad) { set_error("ficp_batch_create: weight table larger than the plot size class"); return kErrInvalid; }
        for (int st = 0; st < prm->n_stages; ++st) {
            double* gdst = h_tabs.data() + ((size_t)(tb * prm->n_stages + st) * 2) * npad;
            double* cdst = gdst + npad;
            for (int k = 1; k <= nt; ++k) {
                const double c = weights[weight_offsets[tb] + (long long)st * nt + (k - 1)];
                const int lane = (k - 1) / e, r = (k - 1) % e;
                cdst[r * 32 + lane] = c;
                gdst[r * 32 + lane] = c * c / (double)k;
            }
        }
    }

    // ---- device buffers
    FICP_CUDA(dev_alloc(reinterpret_cast<void**>(&b->d_plots), sizeof(PlotMeta) * (size_t)n_plots, s));
    FICP_CUDA(dev_alloc(reinterpret_cast<void**>(&b->d_hyp), sizeof(double) * 6 * (size_t)n_hyp, s));
    FICP_CUDA(dev_alloc(reinterpret_cast<void**>(&b->d_tabs), sizeof(double) * h_tabs.size(), s));
    FICP_CUDA(dev_alloc(reinterpret_cast<void**>(&b->d_results), sizeof(HypResult) * (size_t)n_plots * n_hyp_local, s));
    FICP_CUDA(dev_alloc(reinterpret_cast<void**>(&b->d_best), sizeof(unsigned long long) * (size_t)n_plots, s));
    if (b->want_final) FICP_CUDA(dev_alloc(reinterpret_cast<void**>(&b->d_final), sizeof(double) * 2 * (size_t)rows, s));
    FICP_CUDA(dev_alloc(reinterpret_cast<void**>(&b->d_counters), sizeof(int) * (size_t)(n_plots + 1), s));
    FICP_CUDA(dev_alloc(reinterpret_cast<void**>(&b->d_stats), sizeof(unsigned long long) * 8, s));
    if (prm->trace_passes < 0) { set_error("ficp_batch_create: trace_passes must be >= 0"); return kErrInvalid; }
    if (prm->trace_passes > 0) {
        const size_t recs = (size_t)n_plots * n_hyp_local * (size_t)prm->trace_passes;
        if (recs * (size_t)npad * 13 > ((size_t)4 << 30)) { set_error("ficp_batch_create: trace buffers above 4 GB (trace fewer ICPs / passes)"); return kErrTooLarge; }
        b->trace_cap = prm->trace_passes; b->trace_stride = npad;
        FICP_CUDA(dev_alloc(reinterpret_cast<void**>(&b->d_tr_idx), sizeof(int) * recs * npad, s));
        FICP_CUDA(dev_alloc(reinterpret_cast<void**>(&b->d_tr_d2), sizeof(double) * recs * npad, s));
        FICP_CUDA(dev_alloc(reinterpret_cast<void**>(&b->d_tr_in), recs * npad, s));
        FICP_CUDA(dev_alloc(reinterpret_cast<void**>(&b->d_tr_k), sizeof(int) * recs, s));
        FICP_CUDA(dev_alloc(reinterpret_cast<void**>(&b->d_tr_f), sizeof(double) * recs, s));
    }
    FICP_CUDA(cudaMemcpyAsync(b->d_plots, plots.data(), sizeof(PlotMeta) * (size_t)n_plots, cudaMemcpyHostToDevice, s));
    if (direct) { if (int rc = launch_split_rows(d_raw.p, ld, b->d_plots, (int)n_plots, z3, b->d_src_u, b->d_src_z, s)) return rc; }
    FICP_CUDA(cudaMemcpyAsync(b->d_hyp, hyp, sizeof(double) * 6 * (size_t)n_hyp, cudaMemcpyHostToDevice, s));
    FICP_CUDA(cudaMemcpyAsync(b->d_tabs, h_tabs.data(), sizeof(double) * h_tabs.size(), cudaMemcpyHostToDevice, s));
    FICP_CUDA(cudaStreamSynchronize(s));  // the staging block and vectors go out of scope below
    stage.pending = false;

    IcpParams& P = b->params;
    P.grid = t->view;
    P.src_u = b->d_src_u; P.src_z = b->d_src_z; P.plots = b->d_plots; P.n_plots = (int)n_plots;
    P.hyp = b->d_hyp; P.n_hyp = (int)n_hyp; P.hyp_begin = hyp_begin; P.hyp_stride = hyp_stride; P.n_hyp_local = n_hyp_local;
    P.tabs = b->d_tabs; P.n_stages = prm->n_stages; P.threshold = prm->threshold;
    P.max_iter = prm->max_iterations; P.allow_reflection = prm->allow_reflection; P.min_k = prm->min_k;
    P.results = b->d_results; P.best_key = b->d_best; P.final_xy = b->d_final;
    P.slice_counter = b->d_counters; P.hyp_counter = b->d_counters + 1;
    P.slices_per_plot = slices_per_plot; P.n_slices = (int)n_slices;
    P.wcap_pts = wcap_pts; P.wcap_cells = wcap_cells; P.wcap_rows = wcap_rows;
    P.slots = slots_per_cta;
    P.dyn_leads = warps / 2;  // rounds are handed out once the helpers are at least as many as the leads (8 vs 12: same)
    P.stats = b->d_stats;
    P.trace_cap = b->trace_cap; P.trace_stride = b->trace_stride;
    P.tr_idx = b->d_tr_idx; P.tr_d2 = b->d_tr_d2; P.tr_in = b->d_tr_in; P.tr_k = b->d_tr_k; P.tr_f = b->d_tr_f;
    b->launch.e = e; b->launch.z3 = z3; b->launch.warps = warps; b->launch.slots = slots_per_cta; b->launch.elastic = elastic; b->launch.smem = smem;
    b->launch.ctas = (int)std::min<long long>(cta_mode ? n_icps_all : n_slices, resident);
    b->launch.cta_per_icp = cta_mode;
    b->ctas_per_sm = ctas_per_sm;
    guard.armed = false;
    *out = reinterpret_cast<ficp_batch*>(b);
    return kOk;
}

int ficp_batch_get_info(const ficp_batch* bh, ficp_batch_info* info) {
    if (!bh || !info) { set_error("ficp_batch_get_info: null pointer"); return kErrInvalid; }
    const Batch* b = reinterpret_cast<const Batch*>(bh);
    info->n_plots = b->n_plots; info->n_hyp = b->n_hyp; info->n_hyp_local = b->n_hyp_local;
    info->elems_per_lane = b->launch.e; info->match_z = b->z3 ? 1 : 0; info->warps_per_cta = b->launch.warps;
    info->ctas = b->launch.ctas; info->ctas_per_sm = b->ctas_per_sm; info->slices_per_plot = b->params.slices_per_plot;
    info->window_pts_cap = b->params.wcap_pts; info->window_cells_cap = b->params.wcap_cells; info->team_warps = b->launch.warps / b->launch.slots; info->helpers = b->launch.elastic ? 1 : 0;
    info->smem_bytes = (int64_t)b->launch.smem; info->rows = b->rows;
    info->trace_passes = b->trace_cap; info->trace_stride = b->trace_stride; info->cta_per_icp = b->launch.cta_per_icp ? 1 : 0;
    info->rows_direct = b->rows_direct ? 1 : 0; info->reserved = 0;
    return kOk;
}

int ficp_batch_run(ficp_batch* bh, void* stream) {
    if (!bh) { set_error("ficp_batch_run: null batch"); return kErrInvalid; }
    Batch* b = reinterpret_cast<Batch*>(bh);
    cudaStream_t s = (cudaStream_t)stream;
    FICP_CUDA(cudaMemsetAsync(b->d_counters, 0, sizeof(int) * (size_t)(b->n_plots + 1), s));
    FICP_CUDA(cudaMemsetAsync(b->d_stats, 0, sizeof(unsigned long long) * 8, s));
    FICP_CUDA(cudaMemsetAsync(b->d_best, 0xFF, sizeof(unsigned long long) * (size_t)b->n_plots, s));
    const int rc = b->launch.cta_per_icp ? launch_icp_team(b->params, b->launch.e, b->launch.z3, b->launch.ctas, b->launch.smem, s)
                                         : launch_icp(b->params, b->launch, s);
    b->used.record(s);
    b->tgt->used.record(s);
    return rc;
}

int ficp_batch_results(ficp_batch* bh, ficp_hyp_result* results, uint64_t* best_keys, double* final_xy, uint64_t* stats,
                       void* stream) {
    if (!bh) { set_error("ficp_batch_results: null batch"); return kErrInvalid; }
    Batch* b = reinterpret_cast<Batch*>(bh);
    cudaStream_t s = (cudaStream_t)stream;
    if (results)
        FICP_CUDA(cudaMemcpyAsync(results, b->d_results, sizeof(HypResult) * (size_t)b->n_plots * b->n_hyp_local,
                                  cudaMemcpyDeviceToHost, s));
    if (best_keys)
        FICP_CUDA(cudaMemcpyAsync(best_keys, b->d_best, sizeof(uint64_t) * (size_t)b->n_plots, cudaMemcpyDeviceToHost, s));
    if (final_xy) {
        if (!b->want_final) { set_error("ficp_batch_results: batch was not created with want_final_xy (or owns > 1 hypothesis per plot)"); return kErrInvalid; }
        FICP_CUDA(cudaMemcpyAsync(final_xy, b->d_final, sizeof(double) * 2 * (size_t)b->rows, cudaMemcpyDeviceToHost, s));
    }
    if (stats) FICP_CUDA(cudaMemcpyAsync(stats, b->d_stats, sizeof(uint64_t) * 8, cudaMemcpyDeviceToHost, s));
    FICP_CUDA(cudaStreamSynchronize(s));
    return kOk;
}

int ficp_batch_copy_best_keys_device(ficp_batch* bh, void* dst_dev, void* stream) {
    if (!bh || !dst_dev) { set_error("ficp_batch_copy_best_keys_device: null pointer"); return kErrInvalid; }
    Batch* b = reinterpret_cast<Batch*>(bh);
    FICP_CUDA(cudaMemcpyAsync(dst_dev, b->d_best, sizeof(uint64_t) * (size_t)b->n_plots, cudaMemcpyDeviceToDevice,
                              (cudaStream_t)stream));
    b->used.record((cudaStream_t)stream);
    return kOk;
}

int ficp_batch_pack_best_device(ficp_batch* bh, void* dst_dev, void* stream) {
    if (!bh || !dst_dev) { set_error("ficp_batch_pack_best_device: null pointer"); return kErrInvalid; }
    Batch* b = reinterpret_cast<Batch*>(bh);
    const int rc = launch_pack_best(b->d_best, b->d_results, b->d_plots, b->n_plots, b->n_hyp_local, b->params.hyp_begin,
                                    b->params.hyp_stride, b->d_stats, reinterpret_cast<unsigned long long*>(dst_dev), (cudaStream_t)stream);
    b->used.record((cudaStream_t)stream);
    return rc;
}

int ficp_batch_best(ficp_batch* bh, uint64_t* packed_out, uint64_t* stats, void* stream) {
    if (!bh || !packed_out) { set_error("ficp_batch_best: null pointer"); return kErrInvalid; }
    Batch* b = reinterpret_cast<Batch*>(bh);
    cudaStream_t s = (cudaStream_t)stream;
    DevBuf<unsigned long long> tmp(s);
    int rc;
    if ((rc = tmp.alloc((size_t)b->n_plots * kPackWords))) return rc;
    if ((rc = launch_pack_best(b->d_best, b->d_results, b->d_plots, b->n_plots, b->n_hyp_local, b->params.hyp_begin,
                               b->params.hyp_stride, b->d_stats, tmp.p, s)))
        return rc;
    FICP_CUDA(cudaMemcpyAsync(packed_out, tmp.p, sizeof(uint64_t) * kPackWords * (size_t)b->n_plots, cudaMemcpyDeviceToHost, s));
    if (stats) FICP_CUDA(cudaMemcpyAsync(stats, b->d_stats, sizeof(uint64_t) * 8, cudaMemcpyDeviceToHost, s));
    FICP_CUDA(cudaStreamSynchronize(s));
    return kOk;
}

int ficp_batch_trace(ficp_batch* bh, int32_t* idx_out, double* d2_out, uint8_t* inlier_out, int32_t* k_out,
                     double* frmsd_out, void* stream) {
    if (!bh) { set_error("ficp_batch_trace: null batch"); return kErrInvalid; }
    Batch* b = reinterpret_cast<Batch*>(bh);
    if (b->trace_cap <= 0) { set_error("ficp_batch_trace: batch was created with trace_passes = 0"); return kErrInvalid; }
    cudaStream_t s = (cudaStream_t)stream;
    const size_t recs = (size_t)b->n_plots * b->n_hyp_local * (size_t)b->trace_cap, ent = recs * (size_t)b->trace_stride;
    if (idx_out) FICP_CUDA(cudaMemcpyAsync(idx_out, b->d_tr_idx, sizeof(int) * ent, cudaMemcpyDeviceToHost, s));
    if (d2_out) FICP_CUDA(cudaMemcpyAsync(d2_out, b->d_tr_d2, sizeof(double) * ent, cudaMemcpyDeviceToHost, s));
    if (inlier_out) FICP_CUDA(cudaMemcpyAsync(inlier_out, b->d_tr_in, ent, cudaMemcpyDeviceToHost, s));
    if (k_out) FICP_CUDA(cudaMemcpyAsync(k_out, b->d_tr_k, sizeof(int) * recs, cudaMemcpyDeviceToHost, s));
    if (frmsd_out) FICP_CUDA(cudaMemcpyAsync(frmsd_out, b->d_tr_f, sizeof(double) * recs, cudaMemcpyDeviceToHost, s));
    FICP_CUDA(cudaStreamSynchronize(s));
    return kOk;
}

void ficp_batch_destroy(ficp_batch* b) { delete reinterpret_cast<Batch*>(b); }

}  // extern "C"
