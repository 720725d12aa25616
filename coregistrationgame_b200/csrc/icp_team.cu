// Kernel 4b: the two-stage Fractional ICP loop with ONE CTA PER ICP - the latency shape of the persistent kernel.
//
//   run()/_iterate()             /root/reference/ficp.py:122-154   -> team_run_icp
//   find_correspondences          ficp.py:65-71                     -> skip test + search, one tree per thread
//   find_optimal_fraction, frmsd  ficp.py:54-60,73-86               -> block bitonic sort + canonical scan + block arg-min
//   get_n_first_elements          ficp.py:62-63                     -> (d2, index) threshold of the k-th element
//   compute_optimal_transform_2d  ficp.py:89-110                    -> fit_term / fit_reduce / fit_solve (icp_shared.cuh)
//
// Why it exists (DESIGN.md "strong scaling"): in icp_persistent.cu one warp owns an ICP, so a batch smaller than the
// machine - the literal BASELINE config 3, ONE stand x 4096 start poses, sharded over 8 GPUs = 512 ICPs per GPU for
// 2368 warp slots - is bounded by the serial chain of its longest hypothesis (115 passes x ~20 us per pass = 2.4 ms
// measured, 1.83x at 8 GPUs).  Here a whole CTA of T = 32 E threads (one tree per thread, 512 threads for the 500-tree
// stand) works on one ICP and EVERY phase of a pass is cooperative:
//   * skip test: one tree per thread; the trees that fail it are ballot-compacted into the search list;
//   * search: one list entry per thread (first pass: every tree), deferred queries re-compacted and finished in
//     chunks of 32 by all warps;
//   * trimming: block bitonic sort of 64-bit keys (d2 bits | tree index) - register shuffles for strides < 32,
//     double-buffered shared-memory exchanges above - verified against the exact (d2, index) order;
//   * prefix sums of d2 in the association of the one-warp kernel (serial inside chunks of E, Kogge-Stone across
//     the 32 chunks), FRMSD filter and exact arg-min as block reductions (min is order-free);
//   * fit: every thread prepares its tree's term, warp 0 adds them in the one-warp kernel's order.
// The results are BIT-IDENTICAL to icp_persistent.cu (shared arithmetic in icp_shared.cuh; the skip test and search
// bookkeeping only decide whether a query is searched, never its result).  All warps of the CTA are in the same phase,
// so the instruction-cache thrash of 16 independent warps (the top stall of the one-warp kernel) does not occur.
#include <algorithm>
#include <climits>
#include "icp_shared.cuh"

namespace ficp {

namespace {

// Phase clocks (diagnostic build only, -DFICP_PHASE_CLOCKS: make variant NAME=clk FLAGS=-DFICP_PHASE_CLOCKS): thread 0 of
// every CTA adds up the cycles between phase marks; tools/team_phase_clocks.py prints the shares.
#if defined(FICP_PHASE_CLOCKS)
__device__ unsigned long long g_phase_clk[16];
#define PHASE_DECL long long ph_last = clock64(); unsigned long long ph_acc[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}
#define PHASE(n) do { if (tid == 0) { const long long ph_now = clock64(); ph_acc[n] += (unsigned long long)(ph_now - ph_last); ph_last = ph_now; } } while (0)
#define PHASE_FLUSH do { if (tid == 0) { for (int q = 0; q < 16; ++q) atomicAdd(&g_phase_clk[q], ph_acc[q]); } } while (0)
#else
#define PHASE_DECL
#define PHASE(n)
#define PHASE_FLUSH
#endif

// Everything whose size follows from T (and Z3) alone comes first: with T and Z3 template parameters those offsets are
// compile-time constants (immediate offsets from the shared-memory base).  With the window arrays in front they were
// twenty run-time pointers - more than a kernel compiled for 64 registers can keep, so they were spilled and every phase
// of a pass began by fetching its pointers back from local memory (L2: ~400 cycles each; measured 28.0 k -> 17.7 k cycles
// per pass when the kernel was given 128 registers).  Only the window arrays (sized by the launch) stay dynamic.
struct TeamLayout {
    size_t s_u, s_z, s_g, sd2, snn, ssl, list, dlist, slist, sidx, kbuf, sdd, tmp, misc, rowoff, rowdelta, rowg, w_xy, w_z, w_cell, total;
};
__host__ __device__ constexpr size_t team_up16(size_t b) { return (b + 15) & ~size_t(15); }
__host__ __device__ inline TeamLayout team_layout(int t /* tree slots = 32 e */, bool z3, int wcap_pts, int wcap_cells) {
    TeamLayout L{};
    size_t o = 0;
    L.s_u = o; o += team_up16((size_t)t * 16);
    L.s_z = o; o += team_up16(z3 ? (size_t)t * 8 : 0);
    L.s_g = o; o += team_up16((size_t)kMaxStages * t * 8);
    L.sd2 = o; o += team_up16((size_t)t * 8);
    L.snn = o; o += team_up16((size_t)t * 4);
    L.ssl = o; o += team_up16((size_t)t * 2);
    L.list = o; o += team_up16((size_t)t * 2);
    L.dlist = o; o += team_up16((size_t)t * 2);
    L.slist = o; o += team_up16((size_t)t * 2);   // queries the group search hands to the one-lane search (possible exact ties)
    L.sidx = o; o += team_up16((size_t)t * 2);
    L.kbuf = o; o += team_up16((size_t)t * 16);   // two exchange buffers of the sort; later the fit terms ux, uy
    L.sdd = o; o += team_up16((size_t)(t + 32) * 8);   // d2 in trim order (padded: p + p / E); later the fit term vx
    L.tmp = o; o += team_up16((size_t)(t + 32) * 8);   // chunk-serial partial sums of the scan (padded); later the fit term vy
    L.misc = o; o += 1792;
    L.rowoff = o; o += team_up16((size_t)(kWindowRowsCap + 1) * 4);
    L.rowdelta = o; o += team_up16((size_t)kWindowRowsCap * 4);
    L.rowg = o; o += team_up16((size_t)kWindowRowsCap * 4);
    L.w_xy = o; o += team_up16((size_t)wcap_pts * 16);
    L.w_z = o; o += team_up16(z3 ? (size_t)wcap_pts * 8 : 0);
    L.w_cell = o; o += team_up16((size_t)wcap_cells * 4);
    L.total = o;
    return L;
}

// small per-CTA scratch (inside TeamLayout::misc)
struct TeamMisc {
    double stot[32], sexcl[32];     // chunk totals / exclusive prefixes of the canonical scan
    double red_a[32], red_b[32];    // block reductions
    int red_k[32];
    double pose[2][12];             // [buffer]: pose (6) + pose update of the last fit (6); the other buffer takes the next fit
    double sk;                      // S at the chosen k
    double fk[2];                   // FRMSD and RMSE of the pass (computed by the last warp while warp 0 fits)
    int nlist, ndef, nser, nglob, icp, win_ok;
    // CTA-uniform state kept here rather than in every thread's registers (the kernel is compiled for 64)
    int win[4];                     // window rectangle of the staged plot: wx0, wy0, wx1, wy1
    double ub[2];                   // mean of the plot's local coordinates (shift point of the fit)
    unsigned cnt[4];                // this ICP: fix-up rounds, searched, deferred, order rebuilds
    unsigned long long acc[6];      // this CTA: passes, fix-ups, queries, searched, deferred, order rebuilds
};
static_assert(sizeof(TeamMisc) <= 1792, "TeamMisc must fit its reservation");

// Block bitonic sort of NP = T * TPT 32-bit keys (TPT per thread), ascending in position order.  Strides below 32 exchange through
// shuffles; larger strides through two alternating shared-memory buffers (one barrier per step).  Fully unrolled:
// directions and buffers are compile-time, a step is shuffle + predicated min/max.
template <int T, int TPT>
__device__ __forceinline__ void block_sort32(unsigned (&key)[TPT], unsigned* buf, int tid) {
    // element p = tid + kk * T lives in key[kk] of thread tid; strides >= T pair two slots of the same thread
    constexpr int NP = T * TPT;
    int flip = 0;
#pragma unroll
    for (int k = 2; k <= NP; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j >= T) {
                const int dj = j / T;
#pragma unroll
                for (int kk = 0; kk < TPT; ++kk) {
                    if ((kk & dj) == 0) {
                        const bool asc = (((tid + kk * T) & k) == 0);   // (p & NP) == 0: the last phase ascends
                        const unsigned lo = min(key[kk], key[kk | dj]), hi = max(key[kk], key[kk | dj]);
                        key[kk] = asc ? lo : hi;
                        key[kk | dj] = asc ? hi : lo;
                    }
                }
            } else if (j >= 32) {
                unsigned* b = buf + flip * NP;
#pragma unroll
                for (int kk = 0; kk < TPT; ++kk) b[tid + kk * T] = key[kk];
                __syncthreads();
#pragma unroll
                for (int kk = 0; kk < TPT; ++kk) {
                    const int p = tid + kk * T;
                    const unsigned other = b[p ^ j];
                    const bool takemin = (((p & j) == 0) == ((p & k) == 0));
                    key[kk] = takemin ? min(key[kk], other) : max(key[kk], other);
                }
                flip ^= 1;
            } else {
#pragma unroll
                for (int kk = 0; kk < TPT; ++kk) {
                    const int p = tid + kk * T;
                    const unsigned other = __shfl_xor_sync(kFull, key[kk], j);
                    const bool takemin = (((p & j) == 0) == ((p & k) == 0));
                    key[kk] = takemin ? min(key[kk], other) : max(key[kk], other);
                }
            }
        }
    }
}

// Odd-even transposition rounds on the (d2, tree index) pairs in trim order until a round swaps nothing (-> true) or
// `max_rounds` rounds are spent (-> false).  `sdd` uses the padded layout of the scan (PAD).
#define PAD(p) ((p) + (p) / E)
#ifndef FICP_REPAIR_INV
#define FICP_REPAIR_INV 12
#endif
#ifndef FICP_REPAIR_ROUNDS
#define FICP_REPAIR_ROUNDS 5
#endif
constexpr int kRepairMaxInversions = FICP_REPAIR_INV;   // adjacent inversions above which the order is rebuilt by the block sort
constexpr int kRepairMaxRounds = FICP_REPAIR_ROUNDS;
template <int T, int TPT, int E>
__device__ __forceinline__ bool team_repair_order(double* sdd, unsigned short* sidx, int tid, int max_rounds) {
    constexpr int NP = T * TPT;
    for (int round = 0; round < max_rounds; ++round) {
        bool sw = false;
#pragma unroll 1
        for (int par = 0; par < 2; ++par) {
#pragma unroll
            for (int kk = 0; kk < TPT; ++kk) {
                const int p = tid + kk * T;
                if ((p & 1) == par && p + 1 < NP) {
                    const double a = sdd[PAD(p)], b = sdd[PAD(p + 1)];
                    const unsigned short ia = sidx[p], ib = sidx[p + 1];
                    if (key_greater(a, ia, b, ib)) {
                        sdd[PAD(p)] = b; sdd[PAD(p + 1)] = a; sidx[p] = ib; sidx[p + 1] = ia;
                        sw = true;
                    }
                }
            }
            __syncthreads();
        }
        if (!__syncthreads_or(sw)) return true;
    }
    return false;
}

// Group search (icp_shared.cuh: nn_search_group - same candidates, same folds in the same order, same result), arranged for
// the latency of ONE search, which is what a pass of this kernel waits for: the column runs of the three rows are worked out
// first and their row-table lookups issued together on safe operands (no branch around each), and the candidate loop scores
// two candidates per iteration (two independent load / distance chains in flight; the odd tail repeats its last candidate,
// which the top-3 bookkeeping ignores).
#ifndef FICP_TEAM_FAST_GROUP
#define FICP_TEAM_FAST_GROUP 1
#endif
template <bool Z3>
__device__ __forceinline__ int team_search_group(const WindowAcc& acc, const GridGeom& g, bool active, double qx, double qy,
                                                 double qz, int prev, int G, int sub, double& best, int& bestpos, int& cx,
                                                 int& cy, int& lb_hi, int& pos2) {
#if !FICP_TEAM_FAST_GROUP
    return nn_search_group<Z3>(acc, g, active, qx, qy, qz, prev, G, sub, best, bestpos, cx, cy, lb_hi, pos2);
#else
    int status = 0, lb = kHiInf;
    best = kInf; bestpos = -1; pos2 = -1; cx = 0; cy = 0;
    lb_hi = kHiInf;
    if (__ballot_sync(kFull, active) == 0u) return 0;   // a warp without a query has nothing to merge either
    Top3 top = top3_empty();
    if (active) {
        cx = clamp_cell((qx - g.x0) * g.inv_h, g.gw);
        cy = clamp_cell((qy - g.y0) * g.inv_h, g.gh);
        const int xl = (cx > 0) ? cx - 1 : 0, xh = (cx < g.gw - 1) ? cx + 1 : g.gw - 1;
        const int yl = (cy > 0) ? cy - 1 : 0, yh = (cy < g.gh - 1) ? cy + 1 : g.gh - 1;
        if (!acc.covers(xl, xh, yl, yh)) {
            status = 1;
        } else {
            const double seed_d2 = (prev >= 0) ? nn_dist2<Z3>(acc, prev, qx, qy, qz) : kInf;
            double gx[3], gy[3];
            nn_block3_gaps(g, qx, qy, cx, cy, gx, gy);
            const double bound = seed_d2 * FICP_PRUNE_PAD;
            int xa[3], xb[3];
            bool has[3];
#pragma unroll
            for (int ry = 0; ry < 3; ++ry) {
                const int y = cy - 1 + ry;
                const bool rowok = (y >= yl && y <= yh);
                xa[ry] = cx + 2;
                xb[ry] = cx - 2;
#pragma unroll
                for (int rx = 0; rx < 3; ++rx) {
                    const int x = cx - 1 + rx;
                    if (rowok && x >= xl && x <= xh) {
                        const double gap2 = gx[rx] + gy[ry];
                        if (gap2 <= bound) {
                            if (x < xa[ry]) xa[ry] = x;
                            xb[ry] = x;
                        } else {
                            const int c = d_hi(gap2);
                            lb = (c < lb) ? c : lb;
                        }
                    }
                }
                has[ry] = rowok && xa[ry] <= xb[ry];
            }
            int s[3], n[3];
            bool seed_in = false;
#pragma unroll
            for (int ry = 0; ry < 3; ++ry) {
                // a row without a run looks up the query's own cell (always inside the window) and counts nothing
                int s0, e0;
                acc.seg(has[ry] ? cy - 1 + ry : cy, has[ry] ? xa[ry] : cx, has[ry] ? xb[ry] : cx, s0, e0);
                s[ry] = has[ry] ? s0 : 0;
                n[ry] = has[ry] ? e0 - s0 : 0;
                seed_in = seed_in || (has[ry] && prev >= s0 && prev < e0);
            }
            if (prev >= 0 && !seed_in && sub == 0) nn_fold_track_notie(prev, seed_d2, best, bestpos, top);
            const int n01 = n[0] + n[1], total = n01 + n[2];
            const int o1 = s[1] - n[0], o2 = s[2] - n01;
            for (int t = sub; t < total; t += 2 * G) {
                const int t1 = (t + G < total) ? t + G : t;
                const int j0 = t + ((t < n[0]) ? s[0] : (t < n01) ? o1 : o2);
                const int j1 = t1 + ((t1 < n[0]) ? s[0] : (t1 < n01) ? o1 : o2);
                const double da = nn_dist2<Z3>(acc, j0, qx, qy, qz);
                const double db = nn_dist2<Z3>(acc, j1, qx, qy, qz);
                nn_fold_track_notie(j0, da, best, bestpos, top);
                nn_fold_track_notie(j1, db, best, bestpos, top);
            }
        }
    }
    __syncwarp();
    for (int o = 1; o < G; o <<= 1) {
        const double ob = __shfl_xor_sync(kFull, best, o);
        const int op = __shfl_xor_sync(kFull, bestpos, o);
        const int oc1 = __shfl_xor_sync(kFull, top.c1, o), oc2 = __shfl_xor_sync(kFull, top.c2, o);
        const int oc3 = __shfl_xor_sync(kFull, top.c3, o);
        const int op1 = __shfl_xor_sync(kFull, top.p1, o), op2 = __shfl_xor_sync(kFull, top.p2, o);
        const bool lt = ob < best;
        best = lt ? ob : best;
        bestpos = lt ? op : bestpos;
        top3_insert(top, oc1, op1);
        top3_insert(top, oc2, op2);
        top3_insert(top, oc3, -1);   // can only land in the third slot (oc3 >= oc2 >= what slot 2 now holds)
    }
    if (active && status == 0) {
        const int cb = d_hi(best);
        if (bestpos >= 0 && top.c1 == cb && (top.p1 != bestpos || top.c2 == cb)) status = 2;
        const int c = top3_finish(top, bestpos, pos2);
        lb = (c < lb) ? c : lb;
    }
    lb_hi = lb;
    return status;
#endif
}

// Skip test of one round (icp_shared.cuh: nn_test_round - the warp kernel's form, same arithmetic per tree), without a
// branch: every lane evaluates the test on a safe operand (a tree slot past the plot, or one whose code cannot be tested,
// reads the plot's first source row instead of a window point) and only the stores are predicated.  Straight-line code lets
// the compiler overlap the dependent chains of the TWO rounds a warp has (FICP_TEAM_TPT = 2) instead of walking them one
// after the other - each is ~150 instructions at one issue per ~5 cycles.
#ifndef FICP_TEAM_BRANCHFREE_TEST
#define FICP_TEAM_BRANCHFREE_TEST 1
#endif
template <bool Z3>
__device__ __forceinline__ int team_test_round(const WindowAcc& W, const PlotCtx& pc, const Pose& P, const Pose& D,
                                               double* __restrict__ sd2, int* __restrict__ snn, __half* __restrict__ ssl, int e,
                                               int lane) {
    const int i = e * 32 + lane;
    const bool live = i < pc.n;
    const int ii = live ? i : 0;
    const int code = snn[ii];
    const float s0 = __half2float(ssl[ii]);
    const bool test = live && code >= 0 && s0 > 0.f;
    const double2 u = pc.s_u[ii];
    const double ex = D.m00 * u.x + D.m01 * u.y + D.cx;
    const double ey = D.m10 * u.x + D.m11 * u.y + D.cy;
    const double pad = 1e-14 * ((fabs(P.cx) + fabs(P.cy)) + (fabs(u.x) + fabs(u.y)));
    const float move = __fadd_ru(__fsqrt_ru(__double2float_ru(ex * ex + ey * ey)), __double2float_ru(pad));
    const float s1 = __fmul_rd(__fsub_rd(s0, move), 0.99999904632568359375f);
    const __half sh = __float2half_rd(fmaxf(s1, 0.f));
    const float s = __half2float(sh);
    double qx, qy;
    pose_apply(P, u, qx, qy);
    const double qz = Z3 ? pc.s_z[ii] : 0.0;
    const int p1 = code & 0xFFFF, p2 = code >> 16;
    // operands of the two candidate distances: window points when the test applies, else the (always present) source row 0
    const double2* a1 = test ? (W.xy + p1) : pc.s_u;
    const double2* a2 = test ? (W.xy + p2) : pc.s_u;
    const double2 t1 = *a1, t2 = *a2;
    double z1 = 0.0, z2 = 0.0;
    if (Z3) {
        const double* b1 = test ? (W.z + p1) : pc.s_z;
        const double* b2 = test ? (W.z + p2) : pc.s_z;
        z1 = *b1; z2 = *b2;
    }
    double d1, dr;
    {
        const double dx = dsub(qx, t1.x), dy = dsub(qy, t1.y);
        d1 = dadd(dmul(dx, dx), dmul(dy, dy));
        if (Z3) { const double dz = dsub(qz, z1); d1 = dadd(d1, dmul(dz, dz)); }
    }
    {
        const double dx = dsub(qx, t2.x), dy = dsub(qy, t2.y);
        dr = dadd(dmul(dx, dx), dmul(dy, dy));
        if (Z3) { const double dz = dsub(qz, z2); dr = dadd(dr, dmul(dz, dz)); }
    }
    const bool swap = dr < d1;
    const double dmin = swap ? dr : d1;
    const bool settled = test && dmin < (double)s * (double)s && (p2 == p1 || d1 != dr);
    if (test) ssl[i] = sh;
    if (settled) {
        sd2[i] = dmin;
        if (swap) snn[i] = p2 | (p1 << 16);
    }
    return (live && !settled) ? i : -1;
}

// The pose is CTA-uniform state: it lives in shared memory (TeamMisc::pose) and is read where a phase needs it - held in
// registers across the whole pass it cost every thread 24 registers, and the kernel is compiled for 64.
__device__ __forceinline__ Pose ld_pose(const double* p) { return Pose{p[0], p[1], p[2], p[3], p[4], p[5]}; }

struct TeamPtrs {
    const double2* w_xy; const double* w_z; const unsigned* w_cell; const int* rowoff; const int* rowdelta;
};
__device__ __forceinline__ WindowAcc team_window(const TeamPtrs& tp, const int* win, const GridView& G) {
    const int wx0 = win[0], wy0 = win[1], wx1 = win[2], wy1 = win[3];
    return WindowAcc{tp.w_xy, tp.w_z, tp.w_cell, tp.rowoff, tp.rowdelta, G.orig, G.rec, wx0, wy0, wx1, wy1, wx1 - wx0, wy1 - wy0};
}

// Minimum over the warp of a value that is >= +0 and not NaN (+inf allowed): such doubles order like their bit patterns,
// so two integer warp reductions (REDUX: high words, then low words among the lanes holding the smallest high word) replace
// five shuffle + fmin steps.  The minimum is order-free: same bits as the butterfly.
#ifndef FICP_TEAM_REDUX_MIN
#define FICP_TEAM_REDUX_MIN 1
#endif
__device__ __forceinline__ double warp_min_nonneg(double v) {
#if FICP_TEAM_REDUX_MIN
    const unsigned hi = (unsigned)__double2hiint(v), lo = (unsigned)__double2loint(v);
    const unsigned mh = __reduce_min_sync(kFull, hi);
    const unsigned ml = __reduce_min_sync(kFull, (hi == mh) ? lo : 0xFFFFFFFFu);
    return __hiloint2double((int)mh, (int)ml);
#else
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(kFull, v, o));
    return v;
#endif
}
// min over the NW per-warp partials in shared memory, by every warp for itself
__device__ __forceinline__ double warp_min_of(const double* red, int nw, int lane) {
    return warp_min_nonneg((lane < nw) ? red[lane] : kInf);
}

// T threads work on one ICP of NP = T * TPT tree slots (TPT trees per thread: slot tid + kk * T).  TPT = 1 is the
// original shape (one tree per thread, 64 registers); TPT = 2 halves the threads of the two big size classes, so that
// the kernel can be compiled for 128 registers (no spills: 17.7 k instead of 28 k cycles per pass) and
// still keep two CTAs of the 512-slot class on an SM.
#ifndef FICP_TEAM_TPT
#define FICP_TEAM_TPT 2        // trees per thread for the size classes of 512 and 1024 slots
#endif
template <bool Z3, int T, int TPT>
__global__ void __launch_bounds__(T, (65536 / (TPT == 1 ? 64 : 128) / T) > 0 ? (65536 / (TPT == 1 ? 64 : 128) / T) : 1)
icp_team_kernel(const __grid_constant__ IcpParams P) {
    constexpr int NP = T * TPT;        // tree slots (NPAD of the launch)
    constexpr int E = NP / 32;         // elements per lane in the canonical (one-warp) association
    constexpr int NW = T / 32;
    constexpr int IB = (NP == 64) ? 6 : (NP == 128) ? 7 : (NP == 256) ? 8 : (NP == 512) ? 9 : 10;   // bits of a tree index
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const TeamLayout L = team_layout(NP, Z3, P.wcap_pts, P.wcap_cells);
    double2* s_u = reinterpret_cast<double2*>(smem + L.s_u);
    double* s_z = reinterpret_cast<double*>(smem + L.s_z);
    double* s_g = reinterpret_cast<double*>(smem + L.s_g);
    double2* w_xy = reinterpret_cast<double2*>(smem + L.w_xy);
    double* w_z = reinterpret_cast<double*>(smem + L.w_z);
    unsigned* w_cell = reinterpret_cast<unsigned*>(smem + L.w_cell);
    int* rowoff = reinterpret_cast<int*>(smem + L.rowoff);
    int* rowdelta = reinterpret_cast<int*>(smem + L.rowdelta);
    int* rowg = reinterpret_cast<int*>(smem + L.rowg);
    double* sd2 = reinterpret_cast<double*>(smem + L.sd2);
    int* snn = reinterpret_cast<int*>(smem + L.snn);
    __half* ssl = reinterpret_cast<__half*>(smem + L.ssl);
    unsigned short* list = reinterpret_cast<unsigned short*>(smem + L.list);
    unsigned short* dlist = reinterpret_cast<unsigned short*>(smem + L.dlist);
    unsigned short* slist = reinterpret_cast<unsigned short*>(smem + L.slist);
    unsigned short* sidx = reinterpret_cast<unsigned short*>(smem + L.sidx);
    unsigned* kbuf = reinterpret_cast<unsigned*>(smem + L.kbuf);
    double* sdd = reinterpret_cast<double*>(smem + L.sdd);
    double* f_ux = reinterpret_cast<double*>(smem + L.kbuf);          // fit terms alias the (dead) sort buffers
    double* f_uy = reinterpret_cast<double*>(smem + L.kbuf) + NP;
    double* f_vx = reinterpret_cast<double*>(smem + L.sdd);
    double* f_vy = reinterpret_cast<double*>(smem + L.tmp);
    double* spart = reinterpret_cast<double*>(smem + L.tmp);
    unsigned short* finl = list;                                       // inlier flags of the fit (the search list is dead then)
    TeamMisc* M = reinterpret_cast<TeamMisc*>(smem + L.misc);

    const GridView& G_ = P.grid;
    const unsigned lt_mask = (1u << lane) - 1u;
    const long long n_icps = (long long)P.n_plots * P.n_hyp_local;
    int staged_plot = -1;
    if (tid < 6) M->acc[tid] = 0ull;
    const TeamPtrs tp{w_xy, w_z, w_cell, rowoff, rowdelta};
    unsigned n_global = 0;   // per thread
    PHASE_DECL;

    for (;;) {
        __syncthreads();   // everyone is done with the previous ICP (and with M->icp)
        if (tid == 0) M->icp = atomicAdd(P.slice_counter, 1);
        __syncthreads();
        const int c = M->icp;
        if (c >= n_icps) break;
        const int plot = c / P.n_hyp_local, j = c - plot * P.n_hyp_local;
        const PlotMeta pm = P.plots[plot];

        if (plot != staged_plot) {
            // ---- stage the plot: source rows, weight tables, window of grid cells (as in icp_persistent.cu) ----
            for (int i = tid; i < NP; i += T) {
                s_u[i] = (i < pm.n) ? P.src_u[pm.off + i] : make_double2(0.0, 0.0);
                if (Z3) s_z[i] = (i < pm.n) ? P.src_z[pm.off + i] : 0.0;
            }
            const double* tab = P.tabs + (size_t)pm.tab * P.n_stages * 2 * NP;
            for (int i = tid; i < P.n_stages * NP; i += T) s_g[i] = tab[(size_t)(i / NP) * 2 * NP + (i % NP)];
            const int ww = pm.wx1 - pm.wx0, wh = pm.wy1 - pm.wy0;
            bool ok = (ww > 0 && wh > 0 && (long long)ww * wh <= P.wcap_cells && wh <= P.wcap_rows && G_.m > 0);
            if (ok) {
                for (int r = tid; r < wh; r += T) {
                    const size_t rowbase = (size_t)(pm.wy0 + r) * G_.g.gw;
                    const unsigned gs = G_.cell_start[rowbase + pm.wx0], ge = G_.cell_start[rowbase + pm.wx1];
                    rowg[r] = (int)gs;
                    rowdelta[r] = (int)(ge - gs);  // temporarily: the row's point count
                }
            }
            __syncthreads();
            if (tid == 0) {
                if (ok) {
                    int o = 0;
                    for (int r = 0; r < wh; ++r) {
                        const int cnt = rowdelta[r];
                        rowoff[r] = o;
                        rowdelta[r] = rowg[r] - o;
                        o += cnt;
                    }
                    rowoff[wh] = o;
                    if (o > P.wcap_pts || o > 32767) ok = false;  // window positions are packed in 15 bits
                }
                M->win_ok = ok ? 1 : 0;
                if (!ok) atomicAdd(P.stats + 2, 1ull);
            }
            __syncthreads();
            ok = (M->win_ok != 0);
            if (ok) {
                for (int cc = tid; cc < ww * wh; cc += T) {
                    const int r = cc / ww, col = cc - r * ww;
                    const size_t g = (size_t)(pm.wy0 + r) * G_.g.gw + pm.wx0 + col;
                    const unsigned a = G_.cell_start[g], b = G_.cell_start[g + 1];
                    w_cell[cc] = (unsigned)(rowoff[r] + (int)(a - (unsigned)rowg[r])) | ((b - a) << 16);
                }
                for (int r = warp; r < wh; r += NW) {
                    const int cnt = rowoff[r + 1] - rowoff[r], gs = rowg[r], lo = rowoff[r];
                    for (int q = lane; q < cnt; q += 32) {
                        w_xy[lo + q] = grid_xy(G_, gs + q);
                        if (Z3) w_z[lo + q] = grid_z(G_, gs + q);
                    }
                }
            }
            staged_plot = plot;
            __syncthreads();
        }
        const bool win_ok = (M->win_ok != 0);
        const int fixed_k = pm.fixed_k;
        const double* g_ctab = P.tabs + (size_t)pm.tab * P.n_stages * 2 * NP;  // [stage][0]=g [stage][1]=c
        const int n = pm.n;

        // ---- one ICP: start pose of hypothesis h (the expression of icp_persistent.cu / oracle.pre_transform)
        const int h = P.hyp_begin + j * P.hyp_stride;
        const double* hr = P.hyp + (size_t)h * 6;
        int pb = 0;   // current pose buffer
        if (tid == 0) {
            double* mp = M->pose[0];
            mp[0] = hr[0]; mp[1] = hr[1]; mp[2] = hr[2]; mp[3] = hr[3];
            mp[4] = dadd(pm.cinx, hr[4]); mp[5] = dadd(pm.ciny, hr[5]);
            mp[6] = mp[7] = mp[8] = mp[9] = mp[10] = mp[11] = 0.0;
        }   // visible after the first barrier of the pass loop
        if (tid == 0) {
            M->win[0] = pm.wx0; M->win[1] = pm.wy0; M->win[2] = pm.wx1; M->win[3] = pm.wy1;
            M->ub[0] = pm.ubx; M->ub[1] = pm.uby;
            M->cnt[0] = M->cnt[1] = M->cnt[2] = M->cnt[3] = 0u;
        }
        int passes = 0;
#pragma unroll
        for (int kk = 0; kk < TPT; ++kk) {
            const int i = tid + kk * T;
            if (i >= n) { sd2[i] = kInf; snn[i] = -1; }   // padding never changes
            sidx[i] = (unsigned short)i;                  // no previous trim order yet
        }
        PassOut po{0, kInf, 0.0, -1.0, -1};
        PHASE(0);

        for (int st = 0; st < P.n_stages; ++st) {
            const double* sg = s_g + (size_t)st * NP;
            const double* gc = g_ctab + ((size_t)st * 2 + 1) * NP;
            double cur = 0.0;
            int it = 0;
            bool first = true;
            for (;;) {
                // ================= nearest neighbours (ficp.py:65-71)
                const bool have_prev = passes > 0;
                if (tid == 0) { M->nlist = have_prev ? 0 : n; M->ndef = 0; M->nser = 0; }
                __syncthreads();
                if (have_prev) {
                    const WindowAcc Wt = team_window(tp, M->win, G_);
                    const PlotCtx pct{s_u, s_z, n, fixed_k, 0.0, 0.0};
                    const Pose pose_t = ld_pose(M->pose[pb]), dpose_t = ld_pose(M->pose[pb] + 6);
#if FICP_TEAM_BRANCHFREE_TEST
                    if (E == 2 * NW) {
                        // both rounds of the warp side by side, one reservation in the search list for the two
                        const int need0 = team_test_round<Z3>(Wt, pct, pose_t, dpose_t, sd2, snn, ssl, warp, lane);
                        const int need1 = team_test_round<Z3>(Wt, pct, pose_t, dpose_t, sd2, snn, ssl, warp + NW, lane);
                        const unsigned m0 = __ballot_sync(kFull, need0 >= 0), m1 = __ballot_sync(kFull, need1 >= 0);
                        const int c0 = __popc(m0);
                        int base = 0;
                        if (lane == 0 && (m0 | m1)) base = atomicAdd(&M->nlist, c0 + __popc(m1));
                        base = __shfl_sync(kFull, base, 0);
                        if (need0 >= 0) list[base + __popc(m0 & lt_mask)] = (unsigned short)need0;
                        if (need1 >= 0) list[base + c0 + __popc(m1 & lt_mask)] = (unsigned short)need1;
                    } else
#endif
#pragma unroll 1
                    for (int e = warp; e < E; e += NW) {
                        const int need = nn_test_round<Z3>(Wt, pct, pose_t, dpose_t, sd2, snn, ssl, e, lane);
                        const unsigned m = __ballot_sync(kFull, need >= 0);
                        int base = 0;
                        if (lane == 0 && m) base = atomicAdd(&M->nlist, __popc(m));
                        base = __shfl_sync(kFull, base, 0);
                        if (need >= 0) list[base + __popc(m & lt_mask)] = (unsigned short)need;
                    }
                    PHASE(14);
                } else {
#pragma unroll
                    for (int kk = 0; kk < TPT; ++kk) list[tid + kk * T] = (unsigned short)(tid + kk * T);
                }
                __syncthreads();
                PHASE(1);
                const int n_list = M->nlist;
                const Pose pose = ld_pose(M->pose[pb]);
                const WindowAcc W = team_window(tp, M->win, G_);
                const PlotCtx pc{s_u, s_z, n, fixed_k, 0.0, 0.0};
                // lanes per query: a pass that searches few queries spreads each candidate stream over G lanes
                int G = 1;
                if (win_ok) while (G < 16 && n_list * G * 2 <= T) G <<= 1;
                if (G == 1) {
#pragma unroll 1
                    for (int e = warp; e * 32 < n_list; e += NW) {
                        const int defer = nn_round<Z3, false>(G_, W, win_ok, pc, pose, sd2, snn, ssl, list, e, n_list, lane, have_prev);
                        const unsigned m = __ballot_sync(kFull, defer >= 0);
                        int base = 0;
                        if (lane == 0 && m) base = atomicAdd(&M->ndef, __popc(m));
                        base = __shfl_sync(kFull, base, 0);
                        if (defer >= 0) dlist[base + __popc(m & lt_mask)] = (unsigned short)defer;
                    }
                } else {
                    const int slot = tid / G, sub = tid - slot * G;     // G divides 32: a group never straddles warps
                    const bool active = slot < n_list;
                    int i = 0, prev = -1;
                    double qx = 0.0, qy = 0.0, qz = 0.0;
                    if (active) {
                        i = list[slot];
                        pose_apply(pose, s_u[i], qx, qy);
                        if (Z3) qz = s_z[i];
                        prev = have_prev ? code_win(snn[i]) : -1;
                    }
                    double best;
                    int pos, cx, cy, lb_hi, pos2;
                    const int status = team_search_group<Z3>(W, G_.g, active, qx, qy, qz, prev, G, sub, best, pos, cx, cy, lb_hi, pos2);
                    PHASE(13);
                    int defer = -1, ser = -1;
                    if (active && sub == 0) {
                        if (status == 0) {
                            defer = nn_finish_window<false>(G_.g, i, qx, qy, cx, cy, best, pos, pos2, lb_hi, sd2, snn, ssl);
                        } else if (status == 1) {
                            if (!have_prev) snn[i] = -1;
                            ssl[i] = __float2half_rd(0.f);
                            defer = i | 0x8000;
                        } else {
                            ser = i;
                        }
                    }
                    unsigned m = __ballot_sync(kFull, defer >= 0);
                    int base = 0;
                    if (lane == 0 && m) base = atomicAdd(&M->ndef, __popc(m));
                    base = __shfl_sync(kFull, base, 0);
                    if (defer >= 0) dlist[base + __popc(m & lt_mask)] = (unsigned short)defer;
                    m = __ballot_sync(kFull, ser >= 0);
                    base = 0;
                    if (lane == 0 && m) base = atomicAdd(&M->nser, __popc(m));
                    base = __shfl_sync(kFull, base, 0);
                    if (ser >= 0) slist[base + __popc(m & lt_mask)] = (unsigned short)ser;
                    __syncthreads();
                    const int n_ser = M->nser;
                    // possible exact ties: the one-lane search applies the lowest-original-index rule
#pragma unroll 1
                    for (int e = warp; e * 32 < n_ser; e += NW) {
                        const int d2nd = nn_round<Z3, false>(G_, W, win_ok, pc, pose, sd2, snn, ssl, slist, e, n_ser, lane, have_prev);
                        const unsigned m2 = __ballot_sync(kFull, d2nd >= 0);
                        int b2 = 0;
                        if (lane == 0 && m2) b2 = atomicAdd(&M->ndef, __popc(m2));
                        b2 = __shfl_sync(kFull, b2, 0);
                        if (d2nd >= 0) dlist[b2 + __popc(m2 & lt_mask)] = (unsigned short)d2nd;
                    }
                }
                __syncthreads();
                PHASE(2);
                const int n_def = M->ndef;
                for (int base = warp * 32; base < n_def; base += NW * 32)
                    nn_deferred_chunk<Z3, false>(G_, W, pc, pose, sd2, snn, dlist, base, n_def, lane, n_global);
                if (tid == 0) { M->cnt[1] += (unsigned)n_list; M->cnt[2] += (unsigned)n_def; }
                if (tid < 6) M->pose[pb][6 + tid] = 0.0;  // a stage may end without a fit: same pose again, no move
                __syncthreads();
                PHASE(3);

                // ================= trimming (ficp.py:62-63,73-86)
                // Trim order = the exact order of (d2, tree index).  Passes of a converging ICP barely change it, so
                // the PREVIOUS pass's order is tried first: verified against the new distances, repaired by a few
                // odd-even transposition rounds when only a handful of neighbours swapped, and only otherwise rebuilt by
                // the block sort.  The order is a total order, so every route ends in the same permutation.
                // Thread tid owns tree slots AND trim-order positions tid + kk * T.
                double my_d2[TPT], dd[TPT];
                int sidx_t[TPT];
#pragma unroll
                for (int kk = 0; kk < TPT; ++kk) {
                    const int p = tid + kk * T;
                    my_d2[kk] = sd2[p];
                    sidx_t[kk] = sidx[p];
                    dd[kk] = sd2[sidx_t[kk]];
                    sdd[PAD(p)] = dd[kk];
                }
                __syncthreads();
                bool sorted;
                {
                    int ninv = 0;
#pragma unroll
                    for (int kk = 0; kk < TPT; ++kk) {
                        const int p = tid + kk * T;
                        bool inv = false;
                        if (p + 1 < NP) inv = key_greater(dd[kk], (unsigned)sidx_t[kk], sdd[PAD(p + 1)], (unsigned)sidx[p + 1]);
                        ninv += __syncthreads_count(inv);
                    }
                    sorted = (ninv == 0);
                    if (!sorted && ninv <= kRepairMaxInversions) sorted = team_repair_order<T, TPT, E>(sdd, sidx, tid, kRepairMaxRounds);
                }
                PHASE(4);
                if (!sorted) {
                    if (tid == 0) ++M->cnt[3];
                    // packed key: monotone (32 - IB)-bit code of d2 (float bits, rounded down) | tree index; the exact
                    // order is verified afterwards and repaired where quantised codes collide
                    unsigned key[TPT];
#pragma unroll
                    for (int kk = 0; kk < TPT; ++kk) {
                        const unsigned fb = __float_as_uint(__double2float_rd(my_d2[kk]));
                        key[kk] = ((fb >> (IB - 1)) << IB) | (unsigned)(tid + kk * T);
                    }
                    block_sort32<T, TPT>(key, kbuf, tid);
#pragma unroll
                    for (int kk = 0; kk < TPT; ++kk) {
                        sidx_t[kk] = (int)(key[kk] & ((1u << IB) - 1u));
                        dd[kk] = sd2[sidx_t[kk]];
                    }
                    __syncthreads();    // the verification above may still be reading the old order
#pragma unroll
                    for (int kk = 0; kk < TPT; ++kk) {
                        const int p = tid + kk * T;
                        sdd[PAD(p)] = dd[kk];
                        sidx[p] = (unsigned short)sidx_t[kk];
                    }
                    __syncthreads();
                    bool inv = false;
#pragma unroll
                    for (int kk = 0; kk < TPT; ++kk) {
                        const int p = tid + kk * T;
                        if (p + 1 < NP) inv = inv || key_greater(dd[kk], (unsigned)sidx_t[kk], sdd[PAD(p + 1)], (unsigned)sidx[p + 1]);
                    }
                    if (__syncthreads_or(inv)) {
                        if (tid == 0) ++M->cnt[0];
                        (void)team_repair_order<T, TPT, E>(sdd, sidx, tid, 1 << 30);
                    }
                }
                PHASE(5);
                // inclusive prefix sums S_k of d2 in trim order, in the association of the one-warp kernel: serial
                // inside the chunk of E consecutive positions (lane l of warp 0 walks chunk l; the padded layout keeps the
                // 32 lanes on different banks), Kogge-Stone over the 32 chunk totals, prefix + partial
                if (warp == 0) {
                    // all E loads first (independent of the chain), then the E dependent additions
                    double vq[E];
#pragma unroll
                    for (int q = 0; q < E; ++q) vq[q] = sdd[PAD(lane * E + q)];
                    double run = 0.0;
#pragma unroll
                    for (int q = 0; q < E; ++q) {
                        run = __dadd_rn(run, vq[q]);
                        spart[PAD(lane * E + q)] = run;
                    }
                    double inc = run;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const double t = __shfl_up_sync(kFull, inc, o);
                        if (lane >= o) inc = __dadd_rn(inc, t);
                    }
                    double excl = __shfl_up_sync(kFull, inc, 1);
                    if (lane == 0) excl = 0.0;
                    M->sexcl[lane] = excl;
                }
                __syncthreads();
                PHASE(6);
                double S[TPT];
#pragma unroll
                for (int kk = 0; kk < TPT; ++kk) {
                    const int p = tid + kk * T;
                    S[kk] = __dadd_rn(M->sexcl[p / E], spart[PAD(p)]);
                }

                // subset size: first strict minimum of FRMSD(k) = c_k sqrt(S_k / k)  (ficp.py:80-85).
                // The filter G(k) = S_k (c_k^2 / k) leaves the k within rounding distance of the minimum - almost always
                // ONE: then k is known after a block count, and the FRMSD value itself (an fp64 division and a square root,
                // ~600 cycles) is computed by the last warp while the others already work on the fit.
                int kstar;
                bool f_pending = false;        // FRMSD / RMSE of the pass arrive through M->fk
                double fstar = kInf, rstar = 0.0;
                if (fixed_k > 0) {
                    kstar = fixed_k;
#pragma unroll
                    for (int kk = 0; kk < TPT; ++kk)
                        if (tid + kk * T == kstar - 1) M->sk = S[kk];
                    f_pending = true;
                    __syncthreads();   // every S has been read: the fit terms below reuse the scan's buffers
                } else {
                    double g[TPT];
                    double gb = kInf;
#pragma unroll
                    for (int kk = 0; kk < TPT; ++kk) {
                        const int p = tid + kk * T, lp = p / E, r = p - lp * E;
                        g[kk] = (p < n) ? __dmul_rn(S[kk], sg[r * 32 + lp]) : kInf;
                        gb = fmin(gb, g[kk]);
                    }
                    gb = warp_min_nonneg(gb);   // S_k >= 0 and the weights are positive: never negative, never NaN
                    if (lane == 0) M->red_a[warp] = gb;
                    __syncthreads();
                    const double gbest = warp_min_of(M->red_a, NW, lane);
                    const double gthr = gbest * (1.0 + 1e-12);
                    int ncand = 0;
                    bool cand[TPT];
#pragma unroll
                    for (int kk = 0; kk < TPT; ++kk) {
                        const int p = tid + kk * T;
                        cand[kk] = (p < n && g[kk] <= gthr);
                        if (cand[kk]) { M->red_k[0] = p + 1; M->sk = S[kk]; }     // meaningful when there is exactly one
                        ncand += __syncthreads_count(cand[kk]);
                    }
                    if (ncand == 1) {
                        kstar = M->red_k[0];
                        f_pending = true;
                    } else {
                        int kb = INT_MAX;
#pragma unroll
                        for (int kk = 0; kk < TPT; ++kk) {
                            if (cand[kk]) {
                                const int p = tid + kk * T, lp = p / E, r = p - lp * E;
                                const int k = p + 1;
                                const double rm = sqrt(S[kk] / (double)k);
                                const double f = __dmul_rn(__ldg(gc + r * 32 + lp), rm);
                                if (f < fstar || (f == fstar && k < kb)) { fstar = f; kb = k; rstar = rm; }
                            }
                        }
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) {
                            const double of = __shfl_xor_sync(kFull, fstar, o);
                            const int ok = __shfl_xor_sync(kFull, kb, o);
                            const double orr = __shfl_xor_sync(kFull, rstar, o);
                            if (of < fstar || (of == fstar && ok < kb)) { fstar = of; kb = ok; rstar = orr; }
                        }
                        __syncthreads();   // red_k[0] has been read by everybody
                        if (lane == 0) { M->red_b[warp] = fstar; M->red_k[warp] = kb; M->stot[warp] = rstar; }
                        __syncthreads();
                        fstar = (lane < NW) ? M->red_b[lane] : kInf;
                        kb = (lane < NW) ? M->red_k[lane] : INT_MAX;
                        rstar = (lane < NW) ? M->stot[lane] : 0.0;
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) {
                            const double of = __shfl_xor_sync(kFull, fstar, o);
                            const int ok = __shfl_xor_sync(kFull, kb, o);
                            const double orr = __shfl_xor_sync(kFull, rstar, o);
                            if (of < fstar || (of == fstar && ok < kb)) { fstar = of; kb = ok; rstar = orr; }
                        }
                        kstar = (kb == INT_MAX) ? 0 : kb;
                    }
                }
                po.k = kstar;
                if (kstar == 0) {
                    po.f = kInf; po.rmse = 0.0; po.thr = -1.0; po.thr_idx = -1;
                    f_pending = false;
                } else {
                    po.thr_idx = sidx[kstar - 1];
                    po.thr = sd2[po.thr_idx];
                    po.f = fstar;
                    po.rmse = rstar;
                }
                PHASE(7);

                // ================= rigid fit (ficp.py:89-110), started BEFORE the stage logic below has the FRMSD value: the
                // fit needs only the trimmed subset.  It is skipped when no outcome of the convergence test could use it.
                // Every thread prepares its trees' terms; warp 0 adds them in the order of the one-warp kernel - lane l adds
                // the trees l, l + 32, ... in turn, then an xor butterfly over the 32 lanes - and solves; the last warp
                // computes FRMSD / RMSE of the pass meanwhile.
                const bool fit_useful = kstar > 0 && (first ? (it < P.max_iter) : (it + 1 < P.max_iter));
                double ax = 0.0, ay = 0.0;
                if (fit_useful) {
                    const Pose pose_f = ld_pose(M->pose[pb]);
                    fit_shift(pose_f, M->ub[0], M->ub[1], ax, ay);
                    const WindowAcc Wf = team_window(tp, M->win, G_);
#pragma unroll
                    for (int kk = 0; kk < TPT; ++kk) {
                        const int i = tid + kk * T;
                        const bool inl = i < n && (my_d2[kk] < po.thr || (my_d2[kk] == po.thr && i <= po.thr_idx));
                        finl[i] = inl ? 1 : 0;
                        if (inl) {
                            double qx, qy, ux, uy, vx, vy;
                            pose_apply(pose_f, s_u[i], qx, qy);
                            const double2 t = corr_xy(G_, Wf, snn[i]);
                            fit_uv(qx, qy, t.x, t.y, ax, ay, ux, uy, vx, vy);
                            f_ux[i] = ux; f_uy[i] = uy; f_vx[i] = vx; f_vy[i] = vy;
                        }
                    }
                }
                __syncthreads();     // fit terms, M->sk
                PHASE(8);
                if (f_pending && warp == NW - 1) {
                    const int pstar = kstar - 1, lstar = pstar / E, rsel = pstar - lstar * E;
                    const double rm = sqrt(M->sk / (double)kstar);
                    const double fv = __dmul_rn(__ldg(gc + rsel * 32 + lstar), rm);
                    if (lane == 0) { M->fk[0] = fv; M->fk[1] = rm; }
                }
                if (fit_useful && warp == 0) {
                    FitSums fs = fit_zero();
#ifndef FICP_TEAM_BRANCHFREE_FIT
#define FICP_TEAM_BRANCHFREE_FIT 1
#endif
#pragma unroll
                    for (int e = 0; e < E; ++e) {
                        const int i = e * 32 + lane;
#if FICP_TEAM_BRANCHFREE_FIT
                        // no branch around the nine additions (the compiler can then keep the next trees' loads in flight): a
                        // tree outside the subset contributes exact zeros - s + 0.0 and fma(0.0, v, s) return s bit for bit
                        // (the sums start at +0.0 and x + (-x) rounds to +0.0, so no sum is ever -0.0; v is finite) - and
                        // its slots, which may hold anything, are never used in arithmetic
                        const bool in = finl[i] != 0;
                        const double ux = in ? f_ux[i] : 0.0, uy = in ? f_uy[i] : 0.0, vx = in ? f_vx[i] : 0.0, vy = in ? f_vy[i] : 0.0;
                        fit_acc(fs, ux, uy, vx, vy);
#else
                        const double ux = f_ux[i], uy = f_uy[i], vx = f_vx[i], vy = f_vy[i];
                        if (finl[i]) fit_acc(fs, ux, uy, vx, vy);
#endif
                    }
                    fit_reduce(fs);
                    PHASE(9);
                    Pose np = ld_pose(M->pose[pb]), nd;
                    fit_solve(fs, po.k, P.allow_reflection, ax, ay, np, nd);
                    if (lane == 0) {
                        double* mp = M->pose[pb ^ 1];
                        mp[0] = np.m00; mp[1] = np.m01; mp[2] = np.m10; mp[3] = np.m11; mp[4] = np.cx; mp[5] = np.cy;
                        mp[6] = nd.m00; mp[7] = nd.m01; mp[8] = nd.m10; mp[9] = nd.m11; mp[10] = nd.cx; mp[11] = nd.cy;
                    }
                }
                __syncthreads();
                if (f_pending) { po.f = M->fk[0]; po.rmse = M->fk[1]; }
                if (P.trace_cap > 0 && passes < P.trace_cap) {
                    // per-pass trace (tests only): original target row, squared distance, membership in the trimmed subset
                    const size_t rec = (size_t)c * P.trace_cap + passes, base = rec * P.trace_stride;
                    const WindowAcc Wr = team_window(tp, M->win, G_);
#pragma unroll
                    for (int kk = 0; kk < TPT; ++kk) {
                        const int i = tid + kk * T;
                        if (i < n) {
                            const int code = snn[i];
                            int orig = -1;
                            if (code != -1) orig = grid_orig(G_, (code < 0) ? (code & 0x7FFFFFFF) : Wr.global_pos(code & 0xFFFF));
                            P.tr_idx[base + i] = orig;
                            P.tr_d2[base + i] = my_d2[kk];
                            P.tr_in[base + i] = (po.k > 0 && (my_d2[kk] < po.thr || (my_d2[kk] == po.thr && i <= po.thr_idx))) ? 1 : 0;
                        }
                    }
                    if (tid == 0) { P.tr_k[rec] = po.k; P.tr_f[rec] = po.f; }
                }
                ++passes;

                // ================= stage logic (ficp.py:122-147)
                if (first) {
                    if (po.k == 0) break;  // ficp.py:125-126
                    cur = po.f;
                    first = false;
                } else {
                    if (cur - po.f <= P.threshold) break;  // also stops on a regression, keeping the pose (ficp.py:142)
                    cur = po.f;
                    ++it;
                }
                if (it >= P.max_iter) break;
                pb ^= 1;   // the fit is taken
                PHASE(10);
            }
        }
        PHASE(11);

        // ---- results (same record as icp_persistent.cu)
        if (tid == 0) M->nglob = 0;
        __syncthreads();
        {
            const unsigned ng = __reduce_add_sync(kFull, n_global);
            if (lane == 0 && ng) atomicAdd(&M->nglob, (int)ng);
            n_global = 0;
        }
        __syncthreads();
        if (tid == 0) {
            const int ng = M->nglob;
            HypResult res;
            const Pose pose = ld_pose(M->pose[pb]);
            res.m00 = pose.m00; res.m01 = pose.m01; res.m10 = pose.m10; res.m11 = pose.m11;
            res.cx = pose.cx; res.cy = pose.cy;
            res.frmsd = po.f; res.rmse = po.rmse; res.k = po.k; res.passes = passes;
            res.flags = (ng ? 1 : 0) | (win_ok ? 0 : 2);
            res.pad = 0;
            P.results[(size_t)plot * P.n_hyp_local + j] = res;
            const float score = (po.k >= P.min_k && po.k > 0) ? (float)po.f : __int_as_float(0x7F800000);
            const unsigned long long bk = ((unsigned long long)__float_as_uint(score) << 32) | (unsigned)h;
            atomicMin(P.best_key + plot, bk);
            M->acc[0] += passes; M->acc[1] += M->cnt[0]; M->acc[2] += (unsigned long long)passes * n;
            M->acc[3] += M->cnt[1]; M->acc[4] += M->cnt[2]; M->acc[5] += M->cnt[3];
            if (ng) atomicAdd(P.stats + 1, (unsigned long long)ng);
        }
        if (P.final_xy && P.n_hyp_local == 1) {
            const Pose pose_r = ld_pose(M->pose[pb]);
#pragma unroll
            for (int kk = 0; kk < TPT; ++kk) {
                const int i = tid + kk * T;
                if (i < n) {
                    double qx, qy;
                    pose_apply(pose_r, s_u[i], qx, qy);
                    P.final_xy[(pm.off + i) * 2] = qx;
                    P.final_xy[(pm.off + i) * 2 + 1] = qy;
                }
            }
        }
    }
    PHASE(12);
    PHASE_FLUSH;
    if (tid == 0 && M->acc[0]) {
        atomicAdd(P.stats + 0, M->acc[0]);
        atomicAdd(P.stats + 3, M->acc[1]);
        atomicAdd(P.stats + 4, M->acc[2]);
        atomicAdd(P.stats + 5, M->acc[3]);
        atomicAdd(P.stats + 6, M->acc[4]);
        atomicAdd(P.stats + 7, M->acc[5]);
    }
}

template <bool Z3, int T, int TPT>
int team_launch_one(const IcpParams& p, int ctas, size_t smem, cudaStream_t stream) {
    auto kern = icp_team_kernel<Z3, T, TPT>;
    FICP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<ctas, T, smem, stream>>>(p);
    FICP_CUDA(cudaGetLastError());
    return kOk;
}
template <bool Z3, int T, int TPT>
int team_occupancy_one(size_t smem, int* out) {
    auto kern = icp_team_kernel<Z3, T, TPT>;
    FICP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    FICP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(out, kern, T, smem));
    return kOk;
}

}  // namespace

#if defined(FICP_PHASE_CLOCKS)
extern "C" __attribute__((visibility("default"))) int ficp_debug_phase_clocks(unsigned long long* out16, int reset) {
    if (out16 && cudaMemcpyFromSymbol(out16, g_phase_clk, sizeof(unsigned long long) * 16) != cudaSuccess) return -3;
    if (reset) { unsigned long long z[16] = {0}; if (cudaMemcpyToSymbol(g_phase_clk, z, sizeof z) != cudaSuccess) return -3; }
    return 0;
}
#endif

size_t icp_team_smem_bytes(int e, bool z3, int wcap_pts, int wcap_cells, int wcap_rows) {
    (void)wcap_rows;
    return team_layout(32 * e, z3, wcap_pts, wcap_cells).total;
}

// threads of the CTA that works on a plot of 32 e tree slots
int icp_team_threads(int e) { return (e >= 16) ? 32 * e / FICP_TEAM_TPT : 32 * e; }

#define FICP_TEAM_DISPATCH(FN, ...)                                                                                         \
    switch (e) {                                                                                                            \
        case 2: return z3 ? FN<true, 64, 1>(__VA_ARGS__) : FN<false, 64, 1>(__VA_ARGS__);                                   \
        case 4: return z3 ? FN<true, 128, 1>(__VA_ARGS__) : FN<false, 128, 1>(__VA_ARGS__);                                 \
        case 8: return z3 ? FN<true, 256, 1>(__VA_ARGS__) : FN<false, 256, 1>(__VA_ARGS__);                                 \
        case 16: return z3 ? FN<true, 512 / FICP_TEAM_TPT, FICP_TEAM_TPT>(__VA_ARGS__) : FN<false, 512 / FICP_TEAM_TPT, FICP_TEAM_TPT>(__VA_ARGS__);   \
        case 32: return z3 ? FN<true, 1024 / FICP_TEAM_TPT, FICP_TEAM_TPT>(__VA_ARGS__) : FN<false, 1024 / FICP_TEAM_TPT, FICP_TEAM_TPT>(__VA_ARGS__); \
        default: set_error("icp_team: unsupported elements-per-lane"); return kErrInvalid;                                  \
    }

int icp_team_max_ctas_per_sm(int e, bool z3, size_t smem, int* out) { FICP_TEAM_DISPATCH(team_occupancy_one, smem, out) }

int launch_icp_team(const IcpParams& p, int e, bool z3, int ctas, size_t smem, cudaStream_t stream) {
    FICP_TEAM_DISPATCH(team_launch_one, p, ctas, smem, stream)
}

}  // namespace ficp
