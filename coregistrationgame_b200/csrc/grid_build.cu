// Kernel 1a: uniform-grid spatial hash over the Layer-2 (CHM) points, built by counting sort.
//
// Replaces the kd-tree construction the reference repeats on every ICP pass
// (/root/reference/ficp.py:69 `cKDTree(self._xyz_or_xy(target))`); here it is built ONCE per target.
//
// One stream-ordered chain, no host round-trip before the end (round 1 synchronised twice mid-build):
//   geometry   read xyz: bounding box + finiteness; the last CTA to finish reduces the partials, takes the robust
//              extent from a 1024-point sample (skewed targets, see below) and writes the grid geometry to DEVICE memory
//   bin        read xy + geometry: cell id and arrival rank of every point (one L2 atomic per point)
//   scan       single-pass chained scan (decoupled look-back) of the per-cell counts -> cell_start (CSR)
//   scatter    read xyz + cell id + rank -> cell-sorted records (XY: double2 + index; XYZ: one 32 B record)
//   order      cells of 2..32 points are put in original-index order by one thread each (a stable counting sort:
//              the layout does not depend on atomic timing); heavier cells are queued ...
//   heavy      ... and ordered by one warp each (odd-even transposition, cells up to 4096 points).  Cells above that
//              keep arrival order: search RESULTS never depend on the order inside a cell (exact distance ties are
//              settled by original index, tests/hostcheck), only the layout would.
// HBM traffic ~116 B per point (DESIGN.md "grid build"); sizes that depend on the geometry (cell count) are bounded by
// max(m, 1024) cells up front, so nothing has to come back to the host before the last kernel is enqueued.
//
// Skewed targets: the cell edge follows from extent and mean density, so ONE stray coordinate (a (0, 0)
// placeholder row in UTM data) or long thin tails would put almost every point into a handful of cells.  The grid
// therefore spans a ROBUST extent per axis - the central 99 % of a 1024-point sample, widened by a quarter - whenever the
// full extent is more than 1.5x wider than that; points outside are clamped into the border cells (GridGeom::clamped),
// which the search treats as unbounded outward.  Clean data (uniform stands) keeps the full bounding box, unclamped.
#include <cmath>
#include <cstdio>
#include <vector>
#include "ficp_internal.h"
#include "chained_scan.cuh"

namespace ficp {

namespace {

constexpr int kT = kScanT;
constexpr int kSample = 256;           // points sampled for the robust extent (one per thread of the last CTA)
constexpr int kHeavyCap = 4096;        // heavy-cell queue entries
constexpr int kHeavyMaxPts = 4096;     // cells above this keep arrival order

struct BBox {
    double xmin, xmax, ymin, ymax;
    int nonfinite;
};

// device-resident build state (read back once, at the end)
struct BuildState {
    BBox bb;                  // true bounding box of all points + finiteness
    GridGeom g;
    long long nc;             // cells
    unsigned ticket;          // geometry kernel: CTAs finished
    unsigned scan_ticket;     // scan kernel: dynamic block ids
    unsigned n_heavy;         // cells queued for the heavy-cell kernel
    unsigned max_cell;        // largest cell (diagnostic)
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

// Sorts the kSample (= kT) sample values of one axis by counting: thread i ranks its value among all (ties by index)
// and drops it into its slot - no barriers inside, every read a shared-memory broadcast.
__device__ void rank_sort_sample(const double* in, double* out) {
    const double v = in[threadIdx.x];
    int rank = 0;
#pragma unroll 8
    for (int j = 0; j < kSample; ++j) {
        const double o = in[j];
        rank += (o < v || (o == v && j < (int)threadIdx.x)) ? 1 : 0;
    }
    out[rank] = v;
}

// Robust extent of one axis from the sorted sample (see the file header).  Returns clamped?.
__device__ bool robust_axis(const double* sorted, int ns, double lo_all, double hi_all, double& lo, double& hi) {
    lo = lo_all; hi = hi_all;
    if (ns < 200) return false;
    const int r = (ns + 127) / 128;                          // ~0.8 % from either end
    const double a = sorted[r], b = sorted[ns - 1 - r], w = b - a;
    if (!(hi_all - lo_all > 1.5 * w) || !(w >= 0.0)) return false;
    if (w == 0.0) return false;                              // > 99 % of the sample on one coordinate: keep the full extent
    lo = fmax(lo_all, a - 0.25 * w);
    hi = fmin(hi_all, b + 0.25 * w);
    return (lo > lo_all) || (hi < hi_all);
}

// grid-stride partial bounding boxes; the last CTA finishes the geometry
__global__ void __launch_bounds__(kT) geometry_kernel(const double* __restrict__ pts, long long m, int ld, int use_z,
                                                      double pts_per_cell, BBox* __restrict__ partial,
                                                      BuildState* __restrict__ st) {
    __shared__ BBox sh[kT / 32];
    __shared__ double sx[kSample], sy[kSample], tx[kSample], ty[kSample];
    static_assert(kSample == kT, "one sample per thread");
    __shared__ bool last;
    double xmin = kInf, xmax = -kInf, ymin = kInf, ymax = -kInf;
    int bad = 0;
    for (long long i = blockIdx.x * (long long)kT + threadIdx.x; i < m; i += (long long)gridDim.x * kT) {
        const double x = pts[i * ld], y = pts[i * ld + 1];
        bad |= !(isfinite(x) && isfinite(y));
        if (use_z) bad |= !isfinite(pts[i * ld + 2]);
        xmin = fmin(xmin, x); xmax = fmax(xmax, x);
        ymin = fmin(ymin, y); ymax = fmax(ymax, y);
    }
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
        __threadfence();
        last = (atomicAdd(&st->ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    // ---- last CTA: reduce the partials
    xmin = kInf; xmax = -kInf; ymin = kInf; ymax = -kInf; bad = 0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += kT) {
        BBox b;
        {
            const double* pv = reinterpret_cast<const double*>(partial + i);   // written by other CTAs: read through L2
            b.xmin = __ldcg(pv); b.xmax = __ldcg(pv + 1); b.ymin = __ldcg(pv + 2); b.ymax = __ldcg(pv + 3);
            b.nonfinite = __ldcg(reinterpret_cast<const int*>(pv + 4));
        }
        xmin = fmin(xmin, b.xmin); xmax = fmax(xmax, b.xmax);
        ymin = fmin(ymin, b.ymin); ymax = fmax(ymax, b.ymax);
        bad |= b.nonfinite;
    }
    xmin = warp_min(xmin); xmax = warp_max(xmax); ymin = warp_min(ymin); ymax = warp_max(ymax);
    bad = __any_sync(0xFFFFFFFFu, bad);
    __syncthreads();
    if (l == 0) sh[w] = BBox{xmin, xmax, ymin, ymax, bad};
    __syncthreads();
    BBox bb = sh[0];
    for (int i = 1; i < kT / 32; ++i) {
        bb.xmin = fmin(bb.xmin, sh[i].xmin); bb.xmax = fmax(bb.xmax, sh[i].xmax);
        bb.ymin = fmin(bb.ymin, sh[i].ymin); bb.ymax = fmax(bb.ymax, sh[i].ymax);
        bb.nonfinite |= sh[i].nonfinite;
    }
    // ---- robust extent from an evenly strided sample
    const int ns = (int)((m < kSample) ? m : kSample);
    {
        const int k = threadIdx.x;
        if (k < ns && !bb.nonfinite) {
            const long long i = (long long)(((unsigned long long)k * (unsigned long long)m) / (unsigned long long)ns);
            tx[k] = pts[i * ld]; ty[k] = pts[i * ld + 1];
        } else {
            tx[k] = kInf; ty[k] = kInf;      // sorts to the end
        }
    }
    __syncthreads();
    rank_sort_sample(tx, sx);
    rank_sort_sample(ty, sy);
    __syncthreads();
    if (threadIdx.x != 0) return;
    double x0, x1, y0, y1;
    const bool cx = robust_axis(sx, ns, bb.xmin, bb.xmax, x0, x1);
    const bool cy = robust_axis(sy, ns, bb.ymin, bb.ymax, y0, y1);
    // ---- grid geometry: ~pts_per_cell points per cell on average over the (robust) extent
    GridGeom gg;
    const double ex = x1 - x0, ey = y1 - y0;
    const double big = fmax(ex, ey);
    double h;
    if (!(big > 0.0)) {
        h = 1.0;
    } else {
        const double exx = fmax(ex, big * 1e-6), eyy = fmax(ey, big * 1e-6);
        h = sqrt(pts_per_cell * exx * eyy / (double)m);
        if (!(h > 0.0) || !isfinite(h)) h = big;
    }
    const double max_cells = fmax((double)m, 1024.0);
    for (;;) {
        const double gw = floor(ex / h) + 1.0, gh = floor(ey / h) + 1.0;
        if (gw * gh <= max_cells && gw < 2.0e9 && gh < 2.0e9) {
            gg.gw = (int)gw;
            gg.gh = (int)gh;
            break;
        }
        h *= 1.25;
    }
    gg.x0 = x0; gg.y0 = y0; gg.h = h; gg.inv_h = 1.0 / h;
    gg.eps = h * 1e-9 + (fabs(x0) + fabs(y0) + big) * 8e-16;
    gg.tx0 = bb.xmin; gg.tx1 = bb.xmax; gg.ty0 = bb.ymin; gg.ty1 = bb.ymax;
    gg.clamped = (cx || cy) ? 1 : 0;
    gg.pad = 0;
    if (bb.nonfinite) { gg.gw = 1; gg.gh = 1; gg.x0 = gg.y0 = 0.0; gg.h = gg.inv_h = 1.0; gg.eps = 0.0; gg.clamped = 0; }
    st->bb = bb;
    st->g = gg;
    st->nc = (long long)gg.gw * gg.gh;
}

__global__ void __launch_bounds__(kT) bin_kernel(const double* __restrict__ pts, long long m, int ld,
                                                 const BuildState* __restrict__ st, unsigned* __restrict__ cellid,
                                                 unsigned* __restrict__ rank, unsigned* __restrict__ counts) {
    const long long i = blockIdx.x * (long long)kT + threadIdx.x;
    if (i >= m) return;
    const GridGeom g = st->g;
    const double x = pts[i * ld], y = pts[i * ld + 1];
    const int cx = clamp_cell((x - g.x0) * g.inv_h, g.gw);
    const int cy = clamp_cell((y - g.y0) * g.inv_h, g.gh);
    const unsigned c = (unsigned)cy * (unsigned)g.gw + (unsigned)cx;
    cellid[i] = c;
    rank[i] = atomicAdd(counts + c, 1u);
}

// ---- single-pass chained scan (decoupled look-back) over the per-cell counts: chained_scan.cuh
__global__ void __launch_bounds__(kT) scan_kernel(const unsigned* __restrict__ counts, BuildState* __restrict__ st,
                                                  unsigned long long* __restrict__ desc, unsigned* __restrict__ cell_start) {
    chained_scan_block(counts, st->nc, &st->scan_ticket, &st->max_cell, desc, cell_start);
}

__global__ void __launch_bounds__(kT) scatter_kernel(const double* __restrict__ pts, long long m, int ld, int use_z,
                                                     const unsigned* __restrict__ cellid, const unsigned* __restrict__ rank,
                                                     const unsigned* __restrict__ cell_start, double2* __restrict__ xy,
                                                     double4* __restrict__ rec, int* __restrict__ orig) {
    const long long i = blockIdx.x * (long long)kT + threadIdx.x;
    if (i >= m) return;
    const unsigned p = cell_start[cellid[i]] + rank[i];
    if (use_z) {
        rec[p] = make_double4(pts[i * ld], pts[i * ld + 1], pts[i * ld + 2], index_to_bits((int)i));
    } else {
        xy[p] = make_double2(pts[i * ld], pts[i * ld + 1]);
        orig[p] = (int)i;
    }
}

// Orders every cell by original index: the arrival ranks of the bin kernel depend on atomic timing, the final layout
// must not.  One thread per cell for cells of 2..32 points (insertion sort, at most 496 moves); heavier cells are queued.
__global__ void __launch_bounds__(kT) cell_order_kernel(BuildState* __restrict__ st, const unsigned* __restrict__ cell_start,
                                                        double2* __restrict__ xy, double4* __restrict__ rec,
                                                        int* __restrict__ orig, int use_z, unsigned* __restrict__ heavy) {
    const long long c = blockIdx.x * (long long)kT + threadIdx.x;
    if (c >= st->nc) return;
    const unsigned s = cell_start[c], e = cell_start[c + 1];
    if (e - s < 2) return;
    if (e - s > 32) {
        if (e - s <= (unsigned)kHeavyMaxPts) {
            const unsigned k = atomicAdd(&st->n_heavy, 1u);
            if (k < (unsigned)kHeavyCap) heavy[k] = (unsigned)c;
        }
        return;
    }
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

// One warp per queued heavy cell (33..4096 points: clustered targets, duplicates): odd-even transposition on the
// original index, in place.  O(c^2 / 32) warp steps - milliseconds for the largest cell, and only for skewed input.
__global__ void __launch_bounds__(kT) heavy_cell_kernel(const BuildState* __restrict__ st, const unsigned* __restrict__ cell_start,
                                                        double2* __restrict__ xy, double4* __restrict__ rec,
                                                        int* __restrict__ orig, int use_z, const unsigned* __restrict__ heavy) {
    const int lane = threadIdx.x & 31;
    const unsigned n_heavy = min(st->n_heavy, (unsigned)kHeavyCap);
    for (unsigned k = blockIdx.x * (kT / 32) + (threadIdx.x >> 5); k < n_heavy; k += gridDim.x * (kT / 32)) {
        const unsigned c = heavy[k];
        const unsigned s = cell_start[c], cnt = cell_start[c + 1] - s;
        for (unsigned round = 0; round < cnt; ++round) {
            bool sw = false;
            for (unsigned a = (round & 1u) + 2u * lane; a + 1 < cnt; a += 64u) {
                if (use_z) {
                    const double4 p = rec[s + a], q = rec[s + a + 1];
                    if (bits_to_index(p.w) > bits_to_index(q.w)) { rec[s + a] = q; rec[s + a + 1] = p; sw = true; }
                } else {
                    const int p = orig[s + a], q = orig[s + a + 1];
                    if (p > q) {
                        orig[s + a] = q; orig[s + a + 1] = p;
                        const double2 t = xy[s + a]; xy[s + a] = xy[s + a + 1]; xy[s + a + 1] = t;
                        sw = true;
                    }
                }
            }
            __syncwarp();
            // two consecutive rounds without a swap = sorted
            const bool any = __any_sync(0xFFFFFFFFu, sw);
            if (!any) {
                bool sw2 = false;
                const unsigned r2 = round + 1;
                for (unsigned a = (r2 & 1u) + 2u * lane; a + 1 < cnt; a += 64u) {
                    const int p = use_z ? bits_to_index(rec[s + a].w) : orig[s + a];
                    const int q = use_z ? bits_to_index(rec[s + a + 1].w) : orig[s + a + 1];
                    sw2 |= (p > q);
                }
                if (!__any_sync(0xFFFFFFFFu, sw2)) break;
            }
        }
        __syncwarp();
    }
}

}  // namespace

void target_free(Target* t) {
    if (!t) return;
    t->used.wait();  // queries / batches enqueued on any stream have finished reading the index
    const cudaStream_t rs = release_stream(t->alloc_stream);
    dev_free(t->d_xy, rs);
    dev_free(t->d_rec, rs);
    dev_free(t->d_orig, rs);
    dev_free(t->d_cell_start, rs);
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
    t->alloc_stream = stream;
    t->m = m;
    t->has_z = use_z ? 1 : 0;
    t->pts_per_cell = pts_per_cell;
    cudaGetDevice(&t->device);
    struct Guard {
        Target* t; bool armed = true;
        double* raw = nullptr; unsigned* cellid = nullptr; unsigned* rank = nullptr; unsigned* counts = nullptr;
        unsigned long long* desc = nullptr; BBox* part = nullptr; BuildState* st = nullptr; unsigned* heavy = nullptr;
        cudaEvent_t ev0 = nullptr, ev1 = nullptr;
        cudaStream_t s = nullptr;
        ~Guard() {
            if (ev0) cudaEventDestroy(ev0);
            if (ev1) cudaEventDestroy(ev1);
            // scratch is released in stream order: kernels enqueued before an early return may still use it
            dev_free(cellid, s); dev_free(rank, s); dev_free(counts, s); dev_free(desc, s); dev_free(part, s);
            dev_free(st, s); dev_free(heavy, s); dev_free(raw, s);
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

    // every size that depends on the geometry is bounded up front: at most max(m, 1024) cells
    const long long nc_max = std::max<long long>(m, 1024);
    const int nb_geo = (int)std::min<long long>((m + kT - 1) / kT, 148 * 4);
    const int nb_scan_max = (int)((nc_max + kScanChunk - 1) / kScanChunk);
    const unsigned nb_pts = (unsigned)((m + kT - 1) / kT);
    const unsigned nb_cells_max = (unsigned)((nc_max + kT - 1) / kT);
    FICP_CUDA(dev_alloc(reinterpret_cast<void**>(&g.part), sizeof(BBox) * nb_geo, stream));
    FICP_CUDA(dev_alloc(reinterpret_cast<void**>(&g.st), sizeof(BuildState), stream));
    FICP_CUDA(dev_alloc(reinterpret_cast<void**>(&g.cellid), sizeof(unsigned) * (size_t)m, stream));
    FICP_CUDA(dev_alloc(reinterpret_cast<void**>(&g.rank), sizeof(unsigned) * (size_t)m, stream));
    FICP_CUDA(dev_alloc(reinterpret_cast<void**>(&g.counts), sizeof(unsigned) * (size_t)nc_max, stream));
    FICP_CUDA(dev_alloc(reinterpret_cast<void**>(&g.desc), sizeof(unsigned long long) * (size_t)nb_scan_max, stream));
    FICP_CUDA(dev_alloc(reinterpret_cast<void**>(&g.heavy), sizeof(unsigned) * kHeavyCap, stream));
    FICP_CUDA(dev_alloc(reinterpret_cast<void**>(&t->d_cell_start), sizeof(unsigned) * (size_t)(nc_max + 1), stream));
    if (use_z) {
        FICP_CUDA(dev_alloc(reinterpret_cast<void**>(&t->d_rec), sizeof(double4) * (size_t)m, stream));
    } else {
        FICP_CUDA(dev_alloc(reinterpret_cast<void**>(&t->d_xy), sizeof(double2) * (size_t)m, stream));
        FICP_CUDA(dev_alloc(reinterpret_cast<void**>(&t->d_orig), sizeof(int) * (size_t)m, stream));
    }
    FICP_CUDA(cudaEventRecord(g.ev0, stream));
    FICP_CUDA(cudaMemsetAsync(g.st, 0, sizeof(BuildState), stream));
    FICP_CUDA(cudaMemsetAsync(g.counts, 0, sizeof(unsigned) * (size_t)nc_max, stream));
    FICP_CUDA(cudaMemsetAsync(g.desc, 0, sizeof(unsigned long long) * (size_t)nb_scan_max, stream));
    geometry_kernel<<<nb_geo, kT, 0, stream>>>(d_pts, m, ld, use_z, pts_per_cell, g.part, g.st);
    bin_kernel<<<nb_pts, kT, 0, stream>>>(d_pts, m, ld, g.st, g.cellid, g.rank, g.counts);
    scan_kernel<<<nb_scan_max, kT, 0, stream>>>(g.counts, g.st, g.desc, t->d_cell_start);
    scatter_kernel<<<nb_pts, kT, 0, stream>>>(d_pts, m, ld, use_z, g.cellid, g.rank, t->d_cell_start, t->d_xy, t->d_rec, t->d_orig);
    cell_order_kernel<<<nb_cells_max, kT, 0, stream>>>(g.st, t->d_cell_start, t->d_xy, t->d_rec, t->d_orig, use_z, g.heavy);
    heavy_cell_kernel<<<64, kT, 0, stream>>>(g.st, t->d_cell_start, t->d_xy, t->d_rec, t->d_orig, use_z, g.heavy);
    FICP_CUDA(cudaGetLastError());
    FICP_CUDA(cudaEventRecord(g.ev1, stream));
    // ---- the one read-back: geometry, bounding box, finiteness
    BuildState hs;
    FICP_CUDA(cudaMemcpyAsync(&hs, g.st, sizeof(BuildState), cudaMemcpyDeviceToHost, stream));
    FICP_CUDA(cudaStreamSynchronize(stream));
    cudaEventElapsedTime(&t->build_ms, g.ev0, g.ev1);
    if (hs.bb.nonfinite) {
        set_error("target contains non-finite coordinates ('x' must be finite)");
        return kErrNonFinite;
    }
    t->bbox[0] = hs.bb.xmin; t->bbox[1] = hs.bb.xmax; t->bbox[2] = hs.bb.ymin; t->bbox[3] = hs.bb.ymax;
    t->max_cell_pts = (long long)hs.max_cell;
    t->view.g = hs.g;
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
