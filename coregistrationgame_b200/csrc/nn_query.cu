// Kernel 1b (standalone form): bulk nearest-neighbour query against a built target index.
//
// Replaces `tree.query(self._xyz_or_xy(source), k=1)` + `target[idx]` of
// /root/reference/ficp.py:70-71 for callers that use FractionalICP.find_correspondences directly.
// One thread per query over the global (L2-resident) grid; the persistent ICP kernel uses the same
// search routine through a shared-memory window instead (icp_persistent.cu).
//
// Outputs per query: original target index (lowest index among exact ties), Euclidean distance
// sqrt(d2) (IEEE, identical bits to scipy) and optionally d2 itself.
#include "ficp_internal.h"
#include "nn_search.cuh"

namespace ficp {

namespace {

template <bool Z3>
__global__ void __launch_bounds__(128) nn_query_kernel(GridView v, const double* __restrict__ q, long long n, int ld,
                                                       int* __restrict__ idx, double* __restrict__ dist,
                                                       double* __restrict__ d2out) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double qx = q[i * ld], qy = q[i * ld + 1];
    const double qz = Z3 ? q[i * ld + 2] : 0.0;
    const GlobalAcc acc = make_global_acc(v);
    double best;
    int pos;
    nn_search_stream<Z3>(acc, v.g, qx, qy, qz, -1, best, pos);
    // a non-finite query has no nearest neighbour (every comparison fails): index -1, distance NaN - never an
    // out-of-range read.  (The reference raises in scipy, ficp.py:70; the host entry point checks and returns -2.)
    idx[i] = (pos >= 0) ? grid_orig(v, pos) : -1;
    if (dist) dist[i] = (pos >= 0) ? sqrt(best) : __longlong_as_double(0x7FF8000000000000LL);
    if (d2out) d2out[i] = (pos >= 0) ? best : __longlong_as_double(0x7FF8000000000000LL);
}

// position of every original target row in the cell-sorted grid (the stepper gathers `target[idx]` from the grid's own
// copies of the coordinates - the same bits - instead of keeping a second copy of the target on the device)
__global__ void __launch_bounds__(256) inverse_perm_kernel(GridView v, int* __restrict__ inv) {
    const long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (p < v.m) inv[grid_orig(v, p)] = (int)p;
}
template <bool Z3>
__global__ void __launch_bounds__(256) gather_grid_rows_kernel(GridView v, const int* __restrict__ inv, const int* __restrict__ idx,
                                                               long long n, double* __restrict__ out) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int p = inv[idx[i]];
    const double2 xy = grid_xy(v, p);
    constexpr int md = Z3 ? 3 : 2;
    out[i * md] = xy.x;
    out[i * md + 1] = xy.y;
    if (Z3) out[i * md + 2] = grid_z(v, p);
}

// Greedy match-and-remove (replaces CHMPlot.remove_matches, chm_plot.py:223-285): the plot's
// trees are visited IN ORDER; each takes its nearest remaining CHM point and removes it when the distance is below
// the tree's threshold.  The order dependence makes it sequential per plot: one thread per plot, plots in parallel.
template <bool Z3>
__global__ void match_remove_kernel(GridView v, const double* __restrict__ trees, const long long* __restrict__ offsets,
                                    int n_plots, int ld, const double* __restrict__ thr, long long* __restrict__ out,
                                    int* __restrict__ scratch) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_plots) return;
    const long long lo = offsets[p], hi = offsets[p + 1];
    MaskedGlobalAcc acc;
    acc.xy = v.xy; acc.rec = v.rec; acc.org = v.orig; acc.cell_start = v.cell_start; acc.gw = v.g.gw;
    acc.removed = scratch + lo;
    acc.n_removed = 0;
    int* rem = scratch + lo;
    for (long long t = lo; t < hi; ++t) {
        out[t] = -1;
        if ((long long)acc.n_removed >= v.m) continue;  // nothing left (the reference breaks out of its loop)
        const double qx = trees[t * ld], qy = trees[t * ld + 1];
        const double qz = Z3 ? trees[t * ld + 2] : 0.0;
        double best;
        int pos;
        nn_search_stream<Z3>(acc, v.g, qx, qy, qz, -1, best, pos);
        if (pos >= 0 && sqrt(best) < thr[t]) {
            rem[acc.n_removed++] = pos;
            out[t] = grid_orig(v, pos);
        }
    }
}

// Radial crop (replaces `cdist(coords, centre) <= dist`, chm_plot.py:144-148, :306-311): marks the
// points within `dist` of (cx, cy).  Only the grid rows / cell runs the disc can touch are visited (one CTA per row,
// the row's run of cells is one contiguous range), i.e. O(points near the disc), not O(M).
__global__ void __launch_bounds__(128) radial_crop_kernel(GridView v, double cx, double cy, double dist, int row0, int col0,
                                                          int col1, unsigned char* __restrict__ mask) {
    const int y = row0 + blockIdx.x;
    const unsigned* row = v.cell_start + (size_t)y * v.g.gw;
    const unsigned s = row[col0], e = row[col1 + 1];
    for (unsigned j = s + threadIdx.x; j < e; j += blockDim.x) {
        const double2 p = grid_xy(v, j);
        const double dx = dsub(p.x, cx), dy = dsub(p.y, cy);
        if (sqrt(dadd(dmul(dx, dx), dmul(dy, dy))) <= dist) mask[grid_orig(v, j)] = 1;
    }
}

// Measurement aid: read-only sweep over a buffer that fits in L2, 16 B per thread per step.
__global__ void __launch_bounds__(256) l2_read_kernel(const uint4* __restrict__ buf, size_t n_vec, int iters,
                                                      unsigned* __restrict__ sink) {
    unsigned acc = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int it = 0; it < iters; ++it)
        for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n_vec; i += stride) {
            const uint4 v = __ldcg(buf + i);
            acc += v.x ^ v.y ^ v.z ^ v.w;
        }
    if (acc == 0x12345678u) *sink = acc;  // keeps the loads alive
}

}  // namespace

int measure_l2_read_gbs(size_t bytes, int iters, double* gbs) {
    struct Res {
        uint4* buf = nullptr; unsigned* sink = nullptr; cudaEvent_t a = nullptr, b = nullptr;
        ~Res() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); dev_free(buf, nullptr); dev_free(sink, nullptr); }
    } r;
    const size_t n_vec = bytes / sizeof(uint4);
    FICP_CUDA(dev_alloc(reinterpret_cast<void**>(&r.buf), n_vec * sizeof(uint4), nullptr));
    FICP_CUDA(dev_alloc(reinterpret_cast<void**>(&r.sink), sizeof(unsigned), nullptr));
    FICP_CUDA(cudaMemset(r.buf, 1, n_vec * sizeof(uint4)));
    FICP_CUDA(cudaEventCreate(&r.a));
    FICP_CUDA(cudaEventCreate(&r.b));
    uint4* buf = r.buf; unsigned* sink = r.sink; cudaEvent_t a = r.a, b = r.b;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    l2_read_kernel<<<sms * 8, 256>>>(buf, n_vec, 2, sink);  // warm the L2
    FICP_CUDA(cudaEventRecord(a));
    l2_read_kernel<<<sms * 8, 256>>>(buf, n_vec, iters, sink);
    FICP_CUDA(cudaEventRecord(b));
    FICP_CUDA(cudaEventSynchronize(b));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    *gbs = (double)n_vec * sizeof(uint4) * iters / (ms * 1e-3) / 1e9;
    return kOk;
}

int launch_inverse_perm(const GridView& v, int* d_inv, cudaStream_t stream) {
    if (v.m <= 0) return kOk;
    inverse_perm_kernel<<<(unsigned)((v.m + 255) / 256), 256, 0, stream>>>(v, d_inv);
    FICP_CUDA(cudaGetLastError());
    return kOk;
}

int launch_gather_grid_rows(const GridView& v, bool z3, const int* d_inv, const int* d_idx, long long n, double* d_out,
                            cudaStream_t stream) {
    if (n <= 0) return kOk;
    const unsigned nb = (unsigned)((n + 255) / 256);
    if (z3) gather_grid_rows_kernel<true><<<nb, 256, 0, stream>>>(v, d_inv, d_idx, n, d_out);
    else gather_grid_rows_kernel<false><<<nb, 256, 0, stream>>>(v, d_inv, d_idx, n, d_out);
    FICP_CUDA(cudaGetLastError());
    return kOk;
}

int launch_radial_crop(const GridView& v, double cx, double cy, double dist, unsigned char* d_mask, cudaStream_t stream) {
    if (v.m <= 0 || !(dist >= 0.0)) return kOk;
    const GridGeom& g = v.g;
    const double pad = dist + g.eps + g.h * 1e-6;
    if (cx + pad < g.tx0 || cy + pad < g.ty0 || cx - pad > g.tx1 || cy - pad > g.ty1) return kOk;   // disc misses the points' bounding box
    const int col0 = clamp_cell((cx - pad - g.x0) * g.inv_h, g.gw), col1 = clamp_cell((cx + pad - g.x0) * g.inv_h, g.gw);
    const int row0 = clamp_cell((cy - pad - g.y0) * g.inv_h, g.gh), row1 = clamp_cell((cy + pad - g.y0) * g.inv_h, g.gh);
    radial_crop_kernel<<<row1 - row0 + 1, 128, 0, stream>>>(v, cx, cy, dist, row0, col0, col1, d_mask);
    FICP_CUDA(cudaGetLastError());
    return kOk;
}

int launch_match_remove(const GridView& v, bool z3, const double* d_trees, const long long* d_offsets, int n_plots,
                        int ld, const double* d_thr, long long* d_out, int* d_scratch, cudaStream_t stream) {
    if (n_plots <= 0) return kOk;
    const int t = 32;
    const unsigned nb = (unsigned)((n_plots + t - 1) / t);
    if (z3)
        match_remove_kernel<true><<<nb, t, 0, stream>>>(v, d_trees, d_offsets, n_plots, ld, d_thr, d_out, d_scratch);
    else
        match_remove_kernel<false><<<nb, t, 0, stream>>>(v, d_trees, d_offsets, n_plots, ld, d_thr, d_out, d_scratch);
    FICP_CUDA(cudaGetLastError());
    return kOk;
}

int launch_nn_query(const GridView& v, bool z3, const double* d_q, long long n, int ld, int* d_idx, double* d_dist,
                    double* d_d2, cudaStream_t stream) {
    if (n <= 0) return kOk;
    if (v.m <= 0) {
        set_error("nn_query: empty target");
        return kErrInvalid;
    }
    const int t = 128;
    const unsigned nb = (unsigned)((n + t - 1) / t);
    if (z3)
        nn_query_kernel<true><<<nb, t, 0, stream>>>(v, d_q, n, ld, d_idx, d_dist, d_d2);
    else
        nn_query_kernel<false><<<nb, t, 0, stream>>>(v, d_q, n, ld, d_idx, d_dist, d_d2);
    FICP_CUDA(cudaGetLastError());
    return kOk;
}

}  // namespace ficp
