// Kernel 1b, bulk form: nearest-neighbour query for LARGE query batches (north_star stage 1: cells staged through
// shared memory with coalesced bulk loads, lowest-index ties, deterministic indices).
//
// Replaces `tree.query(self._xyz_or_xy(source), k=1)` of /root/reference/ficp.py:70 for callers that hand
// find_correspondences a big source array (the per-pass queries of the ICP loop live in icp_persistent.cu / icp_team.cu).
//
// The thread-per-query kernel (nn_query.cu) is bound by L1 requests: every lane of a warp walks its own cells, so every
// candidate is a separate 32 B sector request (ncu: L1/TEX 94 %, 4.4x the algorithmic bytes).  Here the queries are first
// brought into CELL ORDER with the grid build's own machinery, so that the queries of one CTA share a few dozen cells:
//   qbin      one L2 atomic per query: cell id and arrival rank                        (reads q once)
//   qscan     chained scan of the per-cell query counts (chained_scan.cuh)
//   qscatter  perm[cell_start[cell] + rank] = query index                              (4 B per query)
//   search    one CTA per 256 consecutive queries of the cell order.  Their cells are a run inside one grid row (two or
//             three rows when the chunk wraps): the three rows of target points around that run are three CONTIGUOUS
//             ranges of the cell-sorted target and are brought into shared memory with one `cp.async.bulk` each
//             (completion on an mbarrier), the cell table of the window with coalesced loads.  Every thread then resolves
//             its query against shared memory: own cell first (a bound), then the pruned 3x3 block as one flat candidate
//             stream.  Lanes of a warp sit in the same two or three cells, so their shared-memory reads are broadcasts.
//             The few queries the 3x3 block does not settle are compacted and finished on the global grid (ring loop).
// Results are the same bits as the thread-per-query kernel: canonical squared distance, strict minimum, exact ties to
// the lowest original index (a tie flag in the stream, settled in a second look - rare).
// Sparse or scattered query batches (a chunk spanning more than three rows, or a window above the shared-memory
// budget) fall back to the global-grid search per chunk - still in cell order, i.e. with L1 locality.
#include <algorithm>
#include "chained_scan.cuh"
#include "ficp_internal.h"
#include "nn_search.cuh"

namespace ficp {

namespace {

#ifndef FICP_BULK_MINCTAS
#define FICP_BULK_MINCTAS 3   // resident CTAs per SM the search kernel is compiled for (80 registers; 4 CTAs = 64 registers spills: measured slower)
#endif
constexpr int kBT = 256;               // threads per CTA = queries per chunk
constexpr int kTileBytes = 20480;      // staged target records per window (typical window: 3 rows x ~25 cells x 3 points)
constexpr int kTileMaxW = 94;          // window width in cells, halo columns included
constexpr int kMaxSeg = 3;             // grid rows a chunk's queries may span before it is sent to the global grid

struct BulkState {
    unsigned ticket, n_deferred, pad1, pad2;
};

// One window: the queries of grid row `y` inside a chunk, and the three rows of target cells around their cell run.
struct SegPlan {
    int y, wx0, ww1, wy0;     // query row; first window column; window columns + 1; first window row
    int wh, ok, pad0, pad1;   // window rows (1..3); 0 = window above the shared-memory budget: global grid
    int gs[3], cnt[3];        // per window row: first position in the cell-sorted target, points
    int pad2, pad3;
};
struct ChunkPlan {
    int nseg, pad[3];         // 0 = scattered chunk (more than kMaxSeg rows): every query on the global grid
    SegPlan seg[kMaxSeg];
};

struct Deferred {
    double qx, qy, qz, best;
    int pos, qi;
};

__global__ void __launch_bounds__(kBT) qbin_kernel(GridGeom g, const double* __restrict__ q, long long n, int ld,
                                                   uint2* __restrict__ cr, unsigned* __restrict__ counts) {
    const long long i = blockIdx.x * (long long)kBT + threadIdx.x;
    if (i >= n) return;
    const double x = q[i * ld], y = q[i * ld + 1];
    const int cx = clamp_cell((x - g.x0) * g.inv_h, g.gw);
    const int cy = clamp_cell((y - g.y0) * g.inv_h, g.gh);
    const unsigned c = (unsigned)cy * (unsigned)g.gw + (unsigned)cx;
    cr[i] = make_uint2(c, atomicAdd(counts + c, 1u));
}

__global__ void __launch_bounds__(kScanT) qscan_kernel(const unsigned* __restrict__ counts, long long nc,
                                                       BulkState* __restrict__ st, unsigned long long* __restrict__ desc,
                                                       unsigned* __restrict__ qstart) {
    chained_scan_block(counts, nc, &st->ticket, nullptr, desc, qstart);
}

// cell-ordered query records {x, y, z, bits(query index)}: 32 B = one sector, written whole
template <bool Z3>
__global__ void __launch_bounds__(kBT) qscatter_kernel(const double* __restrict__ q, long long n, int ld,
                                                       const uint2* __restrict__ cr, const unsigned* __restrict__ qstart,
                                                       double4* __restrict__ qrec) {
    const long long i = blockIdx.x * (long long)kBT + threadIdx.x;
    if (i >= n) return;
    const uint2 v = cr[i];
    qrec[qstart[v.x] + v.y] = make_double4(q[i * ld], q[i * ld + 1], Z3 ? q[i * ld + 2] : 0.0, index_to_bits((int)i));
}

// One thread per chunk of kBT consecutive cell-ordered queries: from the cells of its first and last query, the windows
// of target cells the chunk needs (kept out of the search kernel: there it would be a chain of dependent global reads
// in front of every CTA's work).
__global__ void __launch_bounds__(128) qplan_kernel(GridView v, const double4* __restrict__ qrec, long long n, int rec_bytes,
                                                    ChunkPlan* __restrict__ plans, long long n_chunks) {
    const long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (c >= n_chunks) return;
    const GridGeom& g = v.g;
    const long long base = c * kBT;
    const long long last = (base + kBT <= n ? base + kBT : n) - 1;
    const double4 a = qrec[base], b = qrec[last];
    const int cxa = clamp_cell((a.x - g.x0) * g.inv_h, g.gw), cya = clamp_cell((a.y - g.y0) * g.inv_h, g.gh);
    const int cxb = clamp_cell((b.x - g.x0) * g.inv_h, g.gw), cyb = clamp_cell((b.y - g.y0) * g.inv_h, g.gh);
    ChunkPlan P;
    P.pad[0] = P.pad[1] = P.pad[2] = 0;
    P.nseg = (cyb - cya + 1 <= kMaxSeg) ? cyb - cya + 1 : 0;
    // first record of every further row inside the chunk (records are in cell order: rows ascend) - binary search
    long long lo_of[kMaxSeg + 1];
    lo_of[0] = base;
    for (int sgi = 1; sgi <= kMaxSeg; ++sgi) {
        lo_of[sgi] = last + 1;
        if (sgi >= P.nseg) continue;
        long long lo = lo_of[sgi - 1], hi = last + 1;   // first record with row >= cya + sgi
        while (lo < hi) {
            const long long mid = (lo + hi) >> 1;
            const int cym = clamp_cell((qrec[mid].y - g.y0) * g.inv_h, g.gh);
            if (cym >= cya + sgi) hi = mid; else lo = mid + 1;
        }
        lo_of[sgi] = lo;
    }
    for (int sgi = 0; sgi < kMaxSeg; ++sgi) {
        SegPlan& S = P.seg[sgi];
        S = SegPlan{};
        if (sgi >= P.nseg) continue;
        const int y = cya + sgi;
        S.y = y; S.ok = 1; S.ww1 = 1;
        if (lo_of[sgi + 1] <= lo_of[sgi]) continue;       // no query of the chunk in this row: empty window
        const int xa = clamp_cell((qrec[lo_of[sgi]].x - g.x0) * g.inv_h, g.gw);          // cell run of the chunk inside row y
        const int xb = clamp_cell((qrec[lo_of[sgi + 1] - 1].x - g.x0) * g.inv_h, g.gw);
        const int wx0 = (xa > 0) ? xa - 1 : 0, wx1 = (xb < g.gw - 1) ? xb + 1 : g.gw - 1;
        const int wy0 = (y > 0) ? y - 1 : 0, wy1 = (y < g.gh - 1) ? y + 1 : g.gh - 1;
        S.wx0 = wx0; S.ww1 = wx1 - wx0 + 2; S.wy0 = wy0; S.wh = wy1 - wy0 + 1;
        long long total = 0;
        for (int r = 0; r < S.wh; ++r) {
            const unsigned* row = v.cell_start + (size_t)(wy0 + r) * g.gw;
            const unsigned gs = __ldg(row + wx0), ge = __ldg(row + wx1 + 1);
            S.gs[r] = (int)gs;
            S.cnt[r] = (int)(ge - gs);
            total += ge - gs;
        }
        S.ok = (S.ww1 - 1 <= kTileMaxW && total * rec_bytes <= kTileBytes) ? 1 : 0;
    }
    plans[c] = P;
}

// ---- mbarrier / bulk-copy primitives (sm_90+; SASS: SYNCS.*, UBLKCP) ---------------------------------------------------
__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst)),
                 "l"(src), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    unsigned ok = 0;
    while (!ok) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_addr(bar)), "r"(parity)
            : "memory");
    }
}

// Window of target cells in shared memory: rows wy0..wy0+wh-1, columns wx0..wx0+ww-1.  `cell` holds, per row, ww + 1
// window-local start positions (the last = end of the row's run); rows are stored back to back.
template <bool REC>
struct TileAcc {
    const unsigned char* pts;   // shared: REC ? 32 B records {x, y, z, bits(orig)} : double2
    const int* cell;            // shared [wh][ww1]
    const int* gorig;           // global: original indices of the cell-sorted target (XY layout)
    int ro1, ro2;               // first local position of rows 1, 2
    int rd0, rd1, rd2;          // global position = local position + rd[row]
    int wx0, wy0, ww1;

    FICP_HD bool covers(int, int, int, int) const { return true; }
    FICP_HD bool admit(int) const { return true; }
    FICP_HD void seg(int y, int xa, int xb, int& s, int& e) const {
        const int* row = cell + (y - wy0) * ww1 - wx0;
        s = row[xa];
        e = row[xb + 1];
    }
    template <bool Z3>
    FICP_HD void load(int j, double& x, double& y, double& zz) const {
        if (REC) {
            const double2* p = reinterpret_cast<const double2*>(pts) + 2 * (size_t)j;
            const double2 a = p[0];
            x = a.x;
            y = a.y;
            if (Z3) zz = p[1].x;
        } else {
            const double2 a = reinterpret_cast<const double2*>(pts)[j];
            x = a.x;
            y = a.y;
        }
    }
    FICP_HD int global_pos(int j) const { return j + ((j >= ro2) ? rd2 : (j >= ro1) ? rd1 : rd0); }
    FICP_HD int orig(int j) const {
        if (REC) return bits_to_index(reinterpret_cast<const double*>(pts)[4 * (size_t)j + 3]);
        return FICP_LDG(gorig + global_pos(j));
    }
};

// cold paths of the search kernel, kept out of line so that they do not set its register count
template <bool Z3>
__device__ __noinline__ void bulk_global_search(const GridView& v, double qx, double qy, double qz, double* best, int* gpos) {
    const GlobalAcc ga = make_global_acc(v);
    double b;
    int p;
    nn_search_stream<Z3>(ga, v.g, qx, qy, qz, -1, b, p);
    *best = b;
    *gpos = p;
}
template <bool Z3>
__device__ __noinline__ void bulk_global_rings(const GridView& v, double qx, double qy, double qz, int cx, int cy, double* best, int* gpos) {
    const GlobalAcc ga = make_global_acc(v);
    double b = *best;
    int p = *gpos;
    nn_ring_loop_impl<Z3>(ga, v.g, qx, qy, qz, cx, cy, 2, b, p);
    *best = b;
    *gpos = p;
}

template <bool REC, bool Z3>
__global__ void __launch_bounds__(kBT, FICP_BULK_MINCTAS) nn_bulk_kernel(GridView v, const double4* __restrict__ qrec, long long n,
                                                      const ChunkPlan* __restrict__ plans, int* __restrict__ idx,
                                                      double* __restrict__ dist, double* __restrict__ d2out,
                                                      BulkState* __restrict__ st, Deferred* __restrict__ dlist, unsigned dcap,
                                                      unsigned long long* __restrict__ counters) {
    constexpr int kRecBytes = REC ? 32 : 16;
    __shared__ __align__(128) unsigned char s_pts[kTileBytes];
    __shared__ __align__(128) double4 s_q[kBT];
    __shared__ int s_cell[3 * (kTileMaxW + 1)];
    __shared__ __align__(8) unsigned long long s_bar;

    const GridGeom& g = v.g;
    const int tid = threadIdx.x;
    const long long base = blockIdx.x * (long long)kBT;
    const int n_here = (int)((n - base < kBT) ? (n - base) : kBT);
    const bool valid = tid < n_here;
    const ChunkPlan* plan = plans + blockIdx.x;
    const int nseg = __ldg(&plan->nseg);
    unsigned phase = 0;
    if (tid == 0) mbar_init(&s_bar, 1);
    __syncthreads();
    // the chunk's query records and (first window) the target rows: bulk copies on one mbarrier phase
    SegPlan S{};
    if (nseg > 0) S = plan->seg[0];
    const bool first_ok = nseg > 0 && S.ok;
    if (tid == 0) {
        unsigned bytes = (unsigned)n_here * 32u;
        if (first_ok) bytes += (unsigned)(S.cnt[0] + S.cnt[1] + S.cnt[2]) * kRecBytes;
        mbar_expect_tx(&s_bar, bytes);
        bulk_g2s(s_q, qrec + base, (unsigned)n_here * 32u, &s_bar);
    }
    double qx = 0.0, qy = 0.0, qz = 0.0, best = kInf;
    int cx = 0, cy = 0, gpos = -1, qi = -1;
    bool done = !valid;
    unsigned n_tile = 0, n_glob = 0, n_ring = 0;
    bool have_q = false;

    for (int sgi = 0; sgi < (nseg > 0 ? nseg : 1); ++sgi) {
        if (sgi > 0) S = plan->seg[sgi];
        const bool tile_ok = nseg > 0 && S.ok;
        const int ro1 = S.cnt[0], ro2 = S.cnt[0] + S.cnt[1], total = ro2 + S.cnt[2];
        if (tile_ok) {
            if (tid == 0) {
                const unsigned char* src = REC ? reinterpret_cast<const unsigned char*>(v.rec) : reinterpret_cast<const unsigned char*>(v.xy);
                const int ro[3] = {0, ro1, ro2};
                if (sgi > 0 && total > 0) mbar_expect_tx(&s_bar, (unsigned)total * kRecBytes);
#pragma unroll
                for (int r = 0; r < 3; ++r)
                    if (S.cnt[r] > 0)
                        bulk_g2s(s_pts + (size_t)ro[r] * kRecBytes, src + (size_t)S.gs[r] * kRecBytes, (unsigned)S.cnt[r] * kRecBytes, &s_bar);
            }
            for (int k = tid; k < S.wh * S.ww1; k += kBT) {
                const int r = k / S.ww1, c = k - r * S.ww1;
                const int off = (r == 0) ? 0 : (r == 1) ? ro1 : ro2;
                s_cell[k] = (int)(__ldg(v.cell_start + (size_t)(S.wy0 + r) * g.gw + S.wx0 + c) - (unsigned)S.gs[r]) + off;
            }
        }
        if (sgi == 0 || (tile_ok && total > 0)) {
            mbar_wait(&s_bar, phase);
            phase ^= 1u;
        }
        if (!have_q) {
            have_q = true;
            if (valid) {
                const double4 rq = s_q[tid];
                qx = rq.x; qy = rq.y; qz = rq.z; qi = bits_to_index(rq.w);
                cx = clamp_cell((qx - g.x0) * g.inv_h, g.gw);   // the expression of qbin_kernel: same cell, same order
                cy = clamp_cell((qy - g.y0) * g.inv_h, g.gh);
                if (!(isfinite(qx) && isfinite(qy) && isfinite(qz))) done = true;   // no nearest neighbour: index -1, NaN
            }
        }
        __syncthreads();   // cell table complete
        const bool active = !done && (nseg == 0 || cy == S.y);
        if (active) {
            done = true;
            if (!tile_ok) {
                // scattered chunk / window above the shared-memory budget: global grid, still in cell order
                bulk_global_search<Z3>(v, qx, qy, qz, &best, &gpos);
                ++n_glob;
            } else {
                const TileAcc<REC> acc{s_pts, s_cell, v.orig, ro1, ro2, S.gs[0], S.gs[1] - ro1, S.gs[2] - ro2, S.wx0, S.wy0, S.ww1};
                int lpos;
                nn_search_block3_unseeded<Z3>(acc, g, qx, qy, qz, cx, cy, best, lpos);
                gpos = (lpos >= 0) ? acc.global_pos(lpos) : -1;
                ++n_tile;
                if (!nn_block_settles(g, qx, qy, cx, cy, 1, best)) {
                    // rings 2, 3, ... follow on the global grid: handed to nn_ring_finish_kernel (all lanes busy there)
                    ++n_ring;
                    const unsigned slot = atomicAdd(&st->n_deferred, 1u);
                    if (slot < dcap) {
                        dlist[slot] = Deferred{qx, qy, qz, best, gpos, qi};
                        qi = -1;   // result written by the finishing kernel
                    } else {
                        bulk_global_rings<Z3>(v, qx, qy, qz, cx, cy, &best, &gpos);
                    }
                }
            }
        }
        if (sgi + 1 < nseg) __syncthreads();   // the window is reused by the next row
    }
    if (valid && qi >= 0) {
        const double nanv = __longlong_as_double(0x7FF8000000000000LL);
        idx[qi] = (gpos >= 0) ? grid_orig(v, gpos) : -1;
        if (dist) dist[qi] = (gpos >= 0) ? sqrt(best) : nanv;
        if (d2out) d2out[qi] = (gpos >= 0) ? best : nanv;
    }
    if (counters) {
        const unsigned a = __reduce_add_sync(0xFFFFFFFFu, n_tile), b = __reduce_add_sync(0xFFFFFFFFu, n_glob),
                       c = __reduce_add_sync(0xFFFFFFFFu, n_ring);
        if ((tid & 31) == 0) {
            if (a) atomicAdd(counters + 0, (unsigned long long)a);
            if (b) atomicAdd(counters + 1, (unsigned long long)b);
            if (c) atomicAdd(counters + 2, (unsigned long long)c);
        }
    }
}

// Queries the 3x3 block did not settle: rings 2, 3, ... on the global grid, seeded with the block's best candidate.
template <bool Z3>
__global__ void __launch_bounds__(128) nn_ring_finish_kernel(GridView v, const BulkState* __restrict__ st,
                                                             const Deferred* __restrict__ dlist, unsigned dcap,
                                                             int* __restrict__ idx, double* __restrict__ dist,
                                                             double* __restrict__ d2out) {
    const unsigned nd = min(st->n_deferred, dcap);
    const GridGeom& g = v.g;
    const GlobalAcc ga = make_global_acc(v);
    for (unsigned k = blockIdx.x * blockDim.x + threadIdx.x; k < nd; k += gridDim.x * blockDim.x) {
        Deferred d = dlist[k];
        const int cx = clamp_cell((d.qx - g.x0) * g.inv_h, g.gw), cy = clamp_cell((d.qy - g.y0) * g.inv_h, g.gh);
        nn_ring_loop_impl<Z3>(ga, g, d.qx, d.qy, d.qz, cx, cy, 2, d.best, d.pos);
        const double nanv = __longlong_as_double(0x7FF8000000000000LL);
        idx[d.qi] = (d.pos >= 0) ? grid_orig(v, d.pos) : -1;
        if (dist) dist[d.qi] = (d.pos >= 0) ? sqrt(d.best) : nanv;
        if (d2out) d2out[d.qi] = (d.pos >= 0) ? d.best : nanv;
    }
}

}  // namespace

// Planner (measured on B200, profiles/r02_nn_bulk_probe_*.jsonl): the bulk kernel pays where a query meets many
// candidates - indexes built with >= ~4.5 points per cell (the ICP density for XYZ) - and the batch is large and dense
// enough for its windows (at least half as many queries as cells).  On sparse grids (bulk-query density, 1.5 - 3 points per
// cell) the thread-per-query kernel is as fast or faster: both are bound by instruction issue / the FP64 pipe there.
bool nn_bulk_applies(const GridView& v, long long n) {
    const long long nc = (long long)v.g.gw * v.g.gh;
    return n >= kBulkMinQueries && n <= 0x7FFFFFFFLL && 2 * n >= nc && (double)v.m >= 4.5 * (double)nc;
}

int launch_nn_query_bulk(const GridView& v, bool z3, const double* d_q, long long n, int ld, int* d_idx, double* d_dist,
                         double* d_d2, unsigned long long* d_counters, cudaStream_t stream) {
    if (n <= 0) return kOk;
    if (v.m <= 0) {
        set_error("nn_query: empty target");
        return kErrInvalid;
    }
    if (n > 0x7FFFFFFFLL) {
        set_error("nn_query: the bulk kernel takes at most 2^31 - 1 queries per call");
        return kErrTooLarge;
    }
    const long long nc = (long long)v.g.gw * v.g.gh;
    const int nb_scan = (int)((nc + kScanChunk - 1) / kScanChunk);
    const long long n_chunks = (n + kBT - 1) / kBT;
    const unsigned dcap = (unsigned)std::max<long long>(n / 4, 4096);
    // one scratch block: [state | scan descriptors | per-cell counts || cell table of the queries | (cell, rank) |
    //                     cell-ordered query records | chunk plans | deferred list]
    auto up = [](size_t b) { return (b + 255) & ~size_t(255); };
    const size_t o_desc = up(sizeof(BulkState)), o_counts = o_desc + up(sizeof(unsigned long long) * (size_t)nb_scan);
    const size_t o_qstart = o_counts + up(sizeof(unsigned) * (size_t)nc);
    const size_t zero_bytes = o_qstart;   // state, descriptors and counts start at zero
    const size_t o_cr = o_qstart + up(sizeof(unsigned) * (size_t)(nc + 1));
    const size_t o_qrec = o_cr + up(sizeof(uint2) * (size_t)n);
    const size_t o_plan = o_qrec + up(sizeof(double4) * (size_t)n);
    const size_t o_def = o_plan + up(sizeof(ChunkPlan) * (size_t)n_chunks);
    const size_t bytes = o_def + up(sizeof(Deferred) * (size_t)dcap);
    unsigned char* scratch = nullptr;
    FICP_CUDA(dev_alloc(reinterpret_cast<void**>(&scratch), bytes, stream));
    struct Free { void* p; cudaStream_t s; ~Free() { dev_free(p, s); } } guard{scratch, stream};   // stream-ordered release
    BulkState* st = reinterpret_cast<BulkState*>(scratch);
    unsigned long long* desc = reinterpret_cast<unsigned long long*>(scratch + o_desc);
    unsigned* counts = reinterpret_cast<unsigned*>(scratch + o_counts);
    unsigned* qstart = reinterpret_cast<unsigned*>(scratch + o_qstart);
    uint2* cr = reinterpret_cast<uint2*>(scratch + o_cr);
    double4* qrec = reinterpret_cast<double4*>(scratch + o_qrec);
    ChunkPlan* plans = reinterpret_cast<ChunkPlan*>(scratch + o_plan);
    Deferred* dlist = reinterpret_cast<Deferred*>(scratch + o_def);
    FICP_CUDA(cudaMemsetAsync(scratch, 0, zero_bytes, stream));
    const unsigned nb = (unsigned)n_chunks;
    qbin_kernel<<<nb, kBT, 0, stream>>>(v.g, d_q, n, ld, cr, counts);
    qscan_kernel<<<nb_scan, kScanT, 0, stream>>>(counts, nc, st, desc, qstart);
    if (z3) qscatter_kernel<true><<<nb, kBT, 0, stream>>>(d_q, n, ld, cr, qstart, qrec);
    else qscatter_kernel<false><<<nb, kBT, 0, stream>>>(d_q, n, ld, cr, qstart, qrec);
    qplan_kernel<<<(unsigned)((n_chunks + 127) / 128), 128, 0, stream>>>(v, qrec, n, v.rec ? 32 : 16, plans, n_chunks);
    const int nb_fin = 148 * 8;
    if (v.rec) {
        if (z3) {
            nn_bulk_kernel<true, true><<<nb, kBT, 0, stream>>>(v, qrec, n, plans, d_idx, d_dist, d_d2, st, dlist, dcap, d_counters);
            nn_ring_finish_kernel<true><<<nb_fin, 128, 0, stream>>>(v, st, dlist, dcap, d_idx, d_dist, d_d2);
        } else {
            nn_bulk_kernel<true, false><<<nb, kBT, 0, stream>>>(v, qrec, n, plans, d_idx, d_dist, d_d2, st, dlist, dcap, d_counters);
            nn_ring_finish_kernel<false><<<nb_fin, 128, 0, stream>>>(v, st, dlist, dcap, d_idx, d_dist, d_d2);
        }
    } else {
        nn_bulk_kernel<false, false><<<nb, kBT, 0, stream>>>(v, qrec, n, plans, d_idx, d_dist, d_d2, st, dlist, dcap, d_counters);
        nn_ring_finish_kernel<false><<<nb_fin, 128, 0, stream>>>(v, st, dlist, dcap, d_idx, d_dist, d_d2);
    }
    FICP_CUDA(cudaGetLastError());
    return kOk;
}

}  // namespace ficp
