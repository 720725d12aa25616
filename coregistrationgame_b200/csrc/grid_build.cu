// Kernel 1a: uniform-grid spatial hash over the Layer-2 (CHM) points, built by counting sort.
//
// Replaces the kd-tree construction the reference repeats on every ICP pass
// (/root/reference/ficp.py:69 `cKDTree(self._xyz_or_xy(target))`); here it is built ONCE per target.
//
// Passes (all HBM-streaming, DESIGN.md "grid build": ~48 B of traffic per point):
//   bbox      read xyz                     -> min/max + finiteness flag
//   bin_count read xy, write cell id       -> per-cell histogram (L2 atomics)
//   scan      exclusive prefix of the histogram -> cell_start (CSR)
//   scatter   read xyz + cell id           -> cell-sorted records (XY: double2 + index; XYZ: one 32 B record)
//   cell_sort orders each cell by original index, so the layout is deterministic (stable sort)
#include <cmath>
#include <cstdio>
#include <vector>
#include "ficp_internal.h"

namespace ficp {

namespace {

constexpr int kT = 256;

struct BBox {
    double xmin, xmax, ymin, ymax;
    int nonfinite;
};

__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xFFFFFFFFu, v, o));
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xFFFFFFFFu, v, o));
    return v;
}

// grid-stride partial bounding boxes; one BBox per block
__global__ void __launch_bounds__(kT) bbox_kernel(const double* __restrict__ pts, long long m, int ld, int use_z,
                                                  BBox* __restrict__ partial) {
    double xmin = kInf, xmax = -kInf, ymin = kInf, ymax = -kInf;
    int bad = 0;
    for (long long i = blockIdx.x * (long long)kT + threadIdx.x; i < m; i += (long long)gridDim.x * kT) {
        const double x = pts[i * ld], y = pts[i * ld + 1];
        bad |= !(isfinite(x) && isfinite(y));
        if (use_z) bad |= !isfinite(pts[i * ld + 2]);
        xmin = fmin(xmin, x); xmax = fmax(xmax, x);
        ymin = fmin(ymin, y); ymax = fmax(ymax, y);
    }
    __shared__ BBox sh[kT / 32];
    xmin = warp_min(xmin); xmax = warp_max(xmax); ymin = warp_min(ymin); ymax = warp_max(ymax);
    bad = __any_sync(0xFFFFFFFFu, bad);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) sh[w] = BBox{xmin, xmax, ymin, ymax, bad};
    __syncthreads();
    if (threadIdx.x == 0) {
        BBox b = sh[0];
        for (int i = 1; i < kT / 32; ++i) {
            b.xmin = fmin(b.xmin, sh[i].xmin); b.xmax = fmax(b.xmax, sh[i].xmax);
            b.ymin = fmin(b.ymin, sh[i].ymin); b.ymax = fmax(b.ymax, sh[i].ymax);
            b.nonfinite |= sh[i].nonfinite;
        }
        partial[blockIdx.x] = b;
    }
}

__global__ void bbox_final_kernel(const BBox* __restrict__ partial, int n, BBox* __restrict__ out) {
    double xmin = kInf, xmax = -kInf, ymin = kInf, ymax = -kInf;
    int bad = 0;
    for (int i = threadIdx.x; i < n; i += 32) {
        const BBox b = partial[i];
        xmin = fmin(xmin, b.xmin); xmax = fmax(xmax, b.xmax);
        ymin = fmin(ymin, b.ymin); ymax = fmax(ymax, b.ymax);
        bad |= b.nonfinite;
    }
    xmin = warp_min(xmin); xmax = warp_max(xmax); ymin = warp_min(ymin); ymax = warp_max(ymax);
    bad = __any_sync(0xFFFFFFFFu, bad);
    if (threadIdx.x == 0) *out = BBox{xmin, xmax, ymin, ymax, bad};
}

__global__ void __launch_bounds__(kT) bin_count_kernel(const double* __restrict__ pts, long long m, int ld, GridGeom g,
                                                       unsigned* __restrict__ cellid, unsigned* __restrict__ counts) {
    const long long i = blockIdx.x * (long long)kT + threadIdx.x;
    if (i >= m) return;
    const double x = pts[i * ld], y = pts[i * ld + 1];
    const int cx = clamp_cell((x - g.x0) * g.inv_h, g.gw);
    const int cy = clamp_cell((y - g.y0) * g.inv_h, g.gh);
    const unsigned c = (unsigned)cy * (unsigned)g.gw + (unsigned)cx;
    cellid[i] = c;
    atomicAdd(counts + c, 1u);
}

// ---- 3-kernel exclusive scan over the histogram ------------------------------------------------
constexpr int kScanPer = 8;
constexpr int kScanChunk = kT * kScanPer;  // 2048 cells per block

__device__ __forceinline__ unsigned block_exclusive_scan(unsigned v, unsigned* total) {
    // exclusive scan of one value per thread across a kT-thread block
    __shared__ unsigned wsum[kT / 32];
    const int l = threadIdx.x & 31, w = threadIdx.x >> 5;
    unsigned inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(0xFFFFFFFFu, inc, o);
        if (l >= o) inc += t;
    }
    if (l == 31) wsum[w] = inc;
    __syncthreads();
    if (w == 0) {
        unsigned s = (l < kT / 32) ? wsum[l] : 0u;
#pragma unroll
        for (int o = 1; o < kT / 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(0xFFFFFFFFu, s, o);
            if (l >= o) s += t;
        }
        if (l < kT / 32) wsum[l] = s;  // inclusive over warps
    }
    __syncthreads();
    const unsigned base = (w > 0) ? wsum[w - 1] : 0u;
    if (total) *total = wsum[kT / 32 - 1];
    const unsigned r = base + inc - v;
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(kT) scan_block_sums_kernel(const unsigned* __restrict__ counts, long long nc,
                                                             unsigned* __restrict__ block_sums) {
    const long long base = blockIdx.x * (long long)kScanChunk;
    unsigned s = 0;
#pragma unroll
    for (int j = 0; j < kScanPer; ++j) {
        const long long i = base + j * kT + threadIdx.x;
        if (i < nc) s += counts[i];
    }
    unsigned tot;
    block_exclusive_scan(s, &tot);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(kT) scan_partials_kernel(unsigned* __restrict__ block_sums, int nb) {
    // single block: exclusive scan in place, chunk by chunk with a running carry
    __shared__ unsigned carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < nb; base += kT) {
        const int i = base + threadIdx.x;
        const unsigned v = (i < nb) ? block_sums[i] : 0u;
        unsigned tot;
        const unsigned ex = block_exclusive_scan(v, &tot);
        const unsigned carry = carry_s;
        if (i < nb) block_sums[i] = carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + tot;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kT) scan_apply_kernel(const unsigned* __restrict__ counts, long long nc,
                                                        const unsigned* __restrict__ block_offsets,
                                                        unsigned* __restrict__ cell_start) {
    // thread t owns kScanPer CONSECUTIVE cells of the block's chunk
    const long long base = blockIdx.x * (long long)kScanChunk + (long long)threadIdx.x * kScanPer;
    unsigned v[kScanPer];
    unsigned s = 0;
#pragma unroll
    for (int j = 0; j < kScanPer; ++j) {
        v[j] = (base + j < nc) ? counts[base + j] : 0u;
        s += v[j];
    }
    unsigned tot;
    unsigned run = block_exclusive_scan(s, &tot) + block_offsets[blockIdx.x];
#pragma unroll
    for (int j = 0; j < kScanPer; ++j) {
        if (base + j < nc) cell_start[base + j] = run;
        run += v[j];
    }
    if (base <= nc - 1 && nc - 1 < base + kScanPer) cell_start[nc] = run;  // total, written by the owner of the last cell
}

__global__ void __launch_bounds__(kT) scatter_kernel(const double* __restrict__ pts, long long m, int ld, int use_z,
                                                     const unsigned* __restrict__ cellid,
                                                     const unsigned* __restrict__ cell_start,
                                                     unsigned* __restrict__ fill, double2* __restrict__ xy,
                                                     double4* __restrict__ rec, int* __restrict__ orig) {
    const long long i = blockIdx.x * (long long)kT + threadIdx.x;
    if (i >= m) return;
    const unsigned c = cellid[i];
    const unsigned p = cell_start[c] + atomicAdd(fill + c, 1u);
    if (use_z) {
        rec[p] = make_double4(pts[i * ld], pts[i * ld + 1], pts[i * ld + 2], index_to_bits((int)i));
    } else {
        xy[p] = make_double2(pts[i * ld], pts[i * ld + 1]);
        orig[p] = (int)i;
    }
}

// Orders every cell by original index (insertion sort; cells hold ~2-3 points).  Makes the
// cell-sorted layout independent of atomic ordering, i.e. a stable counting sort.
__global__ void __launch_bounds__(kT) cell_sort_kernel(long long nc, const unsigned* __restrict__ cell_start,
                                                       double2* __restrict__ xy, double4* __restrict__ rec,
                                                       int* __restrict__ orig, int use_z) {
    const long long c = blockIdx.x * (long long)kT + threadIdx.x;
    if (c >= nc) return;
    const unsigned s = cell_start[c], e = cell_start[c + 1];
    if (use_z) {
        for (unsigned a = s + 1; a < e; ++a) {
            const double4 kr = rec[a];
            const int key = bits_to_index(kr.w);
            unsigned b = a;
            while (b > s && bits_to_index(rec[b - 1].w) > key) {
                rec[b] = rec[b - 1];
                --b;
            }
            rec[b] = kr;
        }
    } else {
        for (unsigned a = s + 1; a < e; ++a) {
            const int key = orig[a];
            const double2 kxy = xy[a];
            unsigned b = a;
            while (b > s && orig[b - 1] > key) {
                orig[b] = orig[b - 1];
                xy[b] = xy[b - 1];
                --b;
            }
            orig[b] = key;
            xy[b] = kxy;
        }
    }
}

}  // namespace

void target_free(Target* t) {
    if (!t) return;
    t->used.wait();  // queries / batches enqueued on any stream have finished reading the index
    dev_free(t->d_xy, cudaStreamPerThread);
    dev_free(t->d_rec, cudaStreamPerThread);
    dev_free(t->d_orig, cudaStreamPerThread);
    dev_free(t->d_cell_start, cudaStreamPerThread);
    delete t;
}

int target_build(const double* pts, int on_device, long long m, int ld, int use_z, double pts_per_cell,
                 cudaStream_t stream, Target** out) {
    *out = nullptr;
    if (m < 0 || ld < 2 || (use_z && ld < 3)) {
        set_error("target_build: need m >= 0 and at least 2 (3 with z) columns");
        return kErrInvalid;
    }
    if (m > 2000000000LL) {
        set_error("target_build: more than 2e9 target points are not supported");
        return kErrTooLarge;
    }
    if (!(pts_per_cell > 0.0)) pts_per_cell = 2.0;
    Target* t = new Target();
    t->m = m;
    t->has_z = use_z ? 1 : 0;
    t->pts_per_cell = pts_per_cell;
    cudaGetDevice(&t->device);
    struct Guard {
        Target* t; bool armed = true;
        double* raw = nullptr; unsigned* cellid = nullptr; unsigned* counts = nullptr; unsigned* fill = nullptr;
        unsigned* bsum = nullptr; BBox* part = nullptr;
        cudaEvent_t ev0 = nullptr, ev1 = nullptr;
        cudaStream_t s = nullptr;
        ~Guard() {
            if (ev0) cudaEventDestroy(ev0);
            if (ev1) cudaEventDestroy(ev1);
            // scratch is released in stream order: kernels enqueued before an early return may still use it
            dev_free(cellid, s); dev_free(counts, s); dev_free(fill, s); dev_free(bsum, s); dev_free(part, s);
            dev_free(raw, s);
            if (armed) { t->used.record(s); target_free(t); }
        }
    } g{t};
    g.s = stream;

    if (m == 0) {  // empty target: valid handle, no grid (callers short-circuit like ficp.py:66-68)
        t->view.m = 0;
        g.armed = false;
        *out = t;
        return kOk;
    }

    const double* d_pts = pts;
    if (!on_device) {
        FICP_CUDA(dev_alloc(reinterpret_cast<void**>(&g.raw), sizeof(double) * (size_t)m * ld, stream));
        FICP_CUDA(cudaMemcpyAsync(g.raw, pts, sizeof(double) * (size_t)m * ld, cudaMemcpyHostToDevice, stream));
        d_pts = g.raw;
    }
    FICP_CUDA(cudaEventCreate(&g.ev0));
    FICP_CUDA(cudaEventCreate(&g.ev1));
    cudaEvent_t ev0 = g.ev0, ev1 = g.ev1;

    // ---- bounding box + finiteness
    const int nb_bbox = (int)std::min<long long>((m + kT - 1) / kT, 148 * 8);
    FICP_CUDA(dev_alloc(reinterpret_cast<void**>(&g.part), sizeof(BBox) * (nb_bbox + 1), stream));
    FICP_CUDA(cudaEventRecord(ev0, stream));
    bbox_kernel<<<nb_bbox, kT, 0, stream>>>(d_pts, m, ld, use_z, g.part);
    bbox_final_kernel<<<1, 32, 0, stream>>>(g.part, nb_bbox, g.part + nb_bbox);
    FICP_CUDA(cudaEventRecord(ev1, stream));
    BBox bb;
    FICP_CUDA(cudaMemcpyAsync(&bb, g.part + nb_bbox, sizeof(BBox), cudaMemcpyDeviceToHost, stream));
    FICP_CUDA(cudaStreamSynchronize(stream));
    float bbox_ms = 0.f;
    cudaEventElapsedTime(&bbox_ms, ev0, ev1);
    if (bb.nonfinite) {
        set_error("target contains non-finite coordinates ('x' must be finite)");
        return kErrNonFinite;
    }
    t->bbox[0] = bb.xmin; t->bbox[1] = bb.xmax; t->bbox[2] = bb.ymin; t->bbox[3] = bb.ymax;

    // ---- grid geometry: ~pts_per_cell points per cell on average
    GridGeom gg{};
    const double ex = bb.xmax - bb.xmin, ey = bb.ymax - bb.ymin;
    const double big = std::max(ex, ey);
    double h;
    if (!(big > 0.0)) {
        h = 1.0;
    } else {
        const double exx = std::max(ex, big * 1e-6), eyy = std::max(ey, big * 1e-6);
        h = std::sqrt(pts_per_cell * exx * eyy / (double)m);
        if (!(h > 0.0) || !std::isfinite(h)) h = big;
    }
    const double max_cells = std::max(4.0 * (double)m, 1024.0);
    for (;;) {
        const double gw = std::floor(ex / h) + 1.0, gh = std::floor(ey / h) + 1.0;
        if (gw * gh <= max_cells && gw < 2.0e9 && gh < 2.0e9) {
            gg.gw = (int)gw;
            gg.gh = (int)gh;
            break;
        }
        h *= 1.25;
    }
    gg.x0 = bb.xmin; gg.y0 = bb.ymin; gg.h = h; gg.inv_h = 1.0 / h;
    gg.eps = h * 1e-9 + (std::fabs(bb.xmin) + std::fabs(bb.ymin) + big) * 8e-16;
    const long long nc = (long long)gg.gw * gg.gh;

    // ---- counting sort
    FICP_CUDA(dev_alloc(reinterpret_cast<void**>(&g.cellid), sizeof(unsigned) * (size_t)m, stream));
    FICP_CUDA(dev_alloc(reinterpret_cast<void**>(&g.counts), sizeof(unsigned) * (size_t)nc, stream));
    FICP_CUDA(dev_alloc(reinterpret_cast<void**>(&g.fill), sizeof(unsigned) * (size_t)nc, stream));
    FICP_CUDA(dev_alloc(reinterpret_cast<void**>(&t->d_cell_start), sizeof(unsigned) * (size_t)(nc + 1), stream));
    if (use_z) {
        FICP_CUDA(dev_alloc(reinterpret_cast<void**>(&t->d_rec), sizeof(double4) * (size_t)m, stream));
    } else {
        FICP_CUDA(dev_alloc(reinterpret_cast<void**>(&t->d_xy), sizeof(double2) * (size_t)m, stream));
        FICP_CUDA(dev_alloc(reinterpret_cast<void**>(&t->d_orig), sizeof(int) * (size_t)m, stream));
    }
    const int nb_scan = (int)((nc + kScanChunk - 1) / kScanChunk);
    FICP_CUDA(dev_alloc(reinterpret_cast<void**>(&g.bsum), sizeof(unsigned) * (size_t)nb_scan, stream));
    FICP_CUDA(cudaEventRecord(ev0, stream));   // device time of the build = bbox kernels + everything from here
    FICP_CUDA(cudaMemsetAsync(g.counts, 0, sizeof(unsigned) * (size_t)nc, stream));
    FICP_CUDA(cudaMemsetAsync(g.fill, 0, sizeof(unsigned) * (size_t)nc, stream));
    const unsigned nb_pts = (unsigned)((m + kT - 1) / kT);
    bin_count_kernel<<<nb_pts, kT, 0, stream>>>(d_pts, m, ld, gg, g.cellid, g.counts);
    scan_block_sums_kernel<<<nb_scan, kT, 0, stream>>>(g.counts, nc, g.bsum);
    scan_partials_kernel<<<1, kT, 0, stream>>>(g.bsum, nb_scan);
    scan_apply_kernel<<<nb_scan, kT, 0, stream>>>(g.counts, nc, g.bsum, t->d_cell_start);
    scatter_kernel<<<nb_pts, kT, 0, stream>>>(d_pts, m, ld, use_z, g.cellid, t->d_cell_start, g.fill, t->d_xy, t->d_rec,
                                              t->d_orig);
    const unsigned nb_cells = (unsigned)((nc + kT - 1) / kT);
    cell_sort_kernel<<<nb_cells, kT, 0, stream>>>(nc, t->d_cell_start, t->d_xy, t->d_rec, t->d_orig, use_z);
    FICP_CUDA(cudaGetLastError());
    FICP_CUDA(cudaEventRecord(ev1, stream));
    FICP_CUDA(cudaStreamSynchronize(stream));
    cudaEventElapsedTime(&t->build_ms, ev0, ev1);
    t->build_ms += bbox_ms;

    t->view.g = gg;
    t->view.xy = t->d_xy;
    t->view.rec = t->d_rec;
    t->view.orig = t->d_orig;
    t->view.cell_start = t->d_cell_start;
    t->view.m = m;
    g.armed = false;
    *out = t;
    return kOk;
}

}  // namespace ficp
