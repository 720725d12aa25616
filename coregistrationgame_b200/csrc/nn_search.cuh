// Exact nearest-neighbour search over the uniform grid (replaces scipy.spatial.cKDTree build +
// query(k=1) at /root/reference/ficp.py:69-70).
//
// One thread resolves one query.  The search visits the query's own cell, then rings of cells of
// growing Chebyshev radius, and stops as soon as the best distance found is strictly smaller than
// the distance from the query to the border of the block of cells already visited - so the result
// is the exact NN over ALL target points, not an approximate one.  Cell segments whose box is
// farther than the current best are skipped.  Exact distance ties resolve to the LOWEST ORIGINAL
// INDEX (stricter than the reference, whose tie choice depends on kd-tree traversal order).
//
// The code is accessor-generic: `GlobalAcc` reads the cell-sorted target from global memory
// (L2-resident), `WindowAcc` reads a per-plot window of cells staged in shared memory.  It is also
// host-compilable (tests/hostcheck) so the ring / termination logic is unit-tested on the CPU.
#pragma once
#include "ficp_common.cuh"

namespace ficp {

#if defined(__CUDA_ARCH__)
#define FICP_LDG(p) __ldg(p)
#else
#define FICP_LDG(p) (*(p))
#endif

// Upper 32 bits of a non-negative double: a monotone integer code of the value, truncated TOWARDS ZERO (20 mantissa
// bits survive).  Free on the device (it is one register of the pair); used for the running lower bounds below.
constexpr int kHiInf = 0x7FF00000;  // code of +inf
FICP_HD int d_hi(double d) {
#if defined(__CUDA_ARCH__)
    return __double2hiint(d);
#else
    long long v;
    __builtin_memcpy(&v, &d, sizeof v);
    return (int)(v >> 32);
#endif
}
FICP_HD double hi_to_double(int hi) {
#if defined(__CUDA_ARCH__)
    return __hiloint2double(hi, 0);
#else
    const long long v = (long long)(unsigned)hi << 32;
    double d;
    __builtin_memcpy(&d, &v, sizeof d);
    return d;
#endif
}

// One 32 B record = one L2 sector = ONE load request: sm_100 has 256-bit global loads (SASS LDG.E.ENL2.256).
FICP_HD void grid_load_rec(const double4* rec, int j, double& x, double& y, double& z, double& w) {
#if defined(__CUDA_ARCH__)
    asm("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(x), "=d"(y), "=d"(z), "=d"(w) : "l"(rec + j));
#else
    const double4 r = rec[j];
    x = r.x; y = r.y; z = r.z; w = r.w;
#endif
}
FICP_HD double2 grid_xy(const GridView& v, long long j) {
    if (v.rec) return FICP_LDG(reinterpret_cast<const double2*>(v.rec + j));
    return FICP_LDG(v.xy + j);
}
FICP_HD double grid_z(const GridView& v, long long j) {   // XYZ layout only
    return FICP_LDG(reinterpret_cast<const double2*>(v.rec + j) + 1).x;
}
FICP_HD int grid_orig(const GridView& v, long long j) {
    if (v.rec) return bits_to_index(FICP_LDG(reinterpret_cast<const double2*>(v.rec + j) + 1).y);
    return FICP_LDG(v.orig + j);
}

struct GlobalAcc {
    const double2* xy;
    const double4* rec;
    const int* org;
    const unsigned* cell_start;
    int gw;

    FICP_HD bool covers(int, int, int, int) const { return true; }
    FICP_HD bool admit(int) const { return true; }
    FICP_HD void seg(int y, int xa, int xb, int& s, int& e) const {
        FICP_ASSERT(y >= 0 && xa >= 0 && xb < gw && xa <= xb);
        const unsigned* row = cell_start + (size_t)y * gw;
        s = (int)FICP_LDG(row + xa);
        e = (int)FICP_LDG(row + xb + 1);
    }
    template <bool Z3>
    FICP_HD void load(int j, double& x, double& y, double& zz) const {
        if (Z3) {
            double w;
            grid_load_rec(rec, j, x, y, zz, w);
        } else if (rec) {  // XY query against an XYZ-built index
            const double2 p = FICP_LDG(reinterpret_cast<const double2*>(rec + j));
            x = p.x;
            y = p.y;
        } else {
            const double2 p = FICP_LDG(xy + j);
            x = p.x;
            y = p.y;
        }
    }
    FICP_HD int orig(int j) const {
        if (rec) return bits_to_index(FICP_LDG(reinterpret_cast<const double2*>(rec + j) + 1).y);
        return FICP_LDG(org + j);
    }
};
FICP_HD GlobalAcc make_global_acc(const GridView& v) { return GlobalAcc{v.xy, v.rec, v.orig, v.cell_start, v.g.gw}; }

// Global grid minus a (small) set of removed points: used by the greedy match-and-remove pass that follows a
// confirmed registration (chm_plot.py:223-285).  `admit` is consulted only for candidates that would become the
// best, so the linear scan over the removed list is cheap.
struct MaskedGlobalAcc : GlobalAcc {
    const int* removed;  // sorted positions already taken
    int n_removed;
    FICP_HD bool admit(int j) const {
        for (int i = 0; i < n_removed; ++i)
            if (removed[i] == j) return false;
        return true;
    }
};

// Window of cells [wx0,wx1) x [wy0,wy1) copied to shared memory.  cell[] packs
// (first local position | count << 16) per window cell; rows are stored back to back, so a run of
// cells within one row is one contiguous range of points.
struct WindowAcc {
    const double2* xy;     // shared
    const double* z;       // shared
    const unsigned* cell;  // shared
    const int* rowoff;     // shared: first local position of each window row (+ total at [wh])
    const int* rowdelta;   // shared: global sorted position = local position + rowdelta[row]
    const int* gorg;       // global: original indices of the sorted target (XY layout)
    const double4* grec;   // global: packed records (XYZ layout) or nullptr
    int wx0, wy0, wx1, wy1, ww, wh;

    FICP_HD bool covers(int xl, int xh, int yl, int yh) const {
        return xl >= wx0 && xh < wx1 && yl >= wy0 && yh < wy1;
    }
    FICP_HD bool admit(int) const { return true; }
    FICP_HD void seg(int y, int xa, int xb, int& s, int& e) const {
        FICP_ASSERT(y >= wy0 && y < wy1 && xa >= wx0 && xb < wx1 && xa <= xb);
        const unsigned* row = cell + (y - wy0) * ww - wx0;
        const unsigned c0 = row[xa];
        s = (int)(c0 & 0xFFFFu);
        const unsigned c1 = (xb == xa) ? c0 : row[xb];
        e = (int)(c1 & 0xFFFFu) + (int)(c1 >> 16);
    }
    template <bool Z3>
    FICP_HD void load(int j, double& x, double& y, double& zz) const {
        FICP_ASSERT(j >= 0 && j < rowoff[wh]);
        const double2 p = xy[j];
        x = p.x;
        y = p.y;
        if (Z3) zz = z[j];
    }
    FICP_HD int global_pos(int j) const {  // rare path (exact ties, final index lookup)
        int lo = 0, hi = wh - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (rowoff[mid] <= j) lo = mid; else hi = mid - 1;
        }
        return j + rowdelta[lo];
    }
    FICP_HD int orig(int j) const {
        const int gp = global_pos(j);
        if (grec) return bits_to_index(FICP_LDG(reinterpret_cast<const double2*>(grec + gp) + 1).y);
        return FICP_LDG(gorg + gp);
    }
};

// Rare paths can be kept out of line (smaller hot loop) or inlined (no call ABI inside divergent code);
// which is faster was measured, see profiles/.
#if defined(__CUDA_ARCH__) && defined(FICP_NOINLINE_TIE)
#define FICP_COLD_TIE __device__ __noinline__
#else
#define FICP_COLD_TIE FICP_HD
#endif
#if defined(__CUDA_ARCH__) && defined(FICP_NOINLINE_RING)
#define FICP_COLD_RING __device__ __noinline__
#else
#define FICP_COLD_RING FICP_HD
#endif

// Exact distance tie between candidate j and the current best (rare): the lower ORIGINAL index wins.
// Out of line: it sits behind every distance comparison and would otherwise be inlined a dozen times.
template <class Acc>
FICP_COLD_TIE int nn_tie_winner(const Acc& acc, int j, int bestpos) {
    if (bestpos < 0 || j == bestpos) return j;
    return (acc.orig(j) < acc.orig(bestpos)) ? j : bestpos;
}

// Fold one scored candidate into the running best.  The common events are "worse" and "better": both are handled
// without a branch (predicated moves), because in a lock-step warp SOME lane improves in almost every iteration
// and a divergent branch would be paid by all 32.  Only an exact tie with a DIFFERENT point (rare; re-meeting the
// seed has j == bestpos) takes a branch, to apply the lowest-original-index rule.
template <class Acc>
FICP_HD void nn_fold(const Acc& acc, int j, double d2, double& best, int& bestpos) {
    const bool lt = d2 < best;
    if (d2 == best && j != bestpos) {
        if (acc.admit(j)) bestpos = nn_tie_winner(acc, j, bestpos);
    }
    if (lt && acc.admit(j)) {
        best = d2;
        bestpos = j;
    }
}

// Running top-3 of the streamed candidates by CODE (d_hi: squared distance truncated towards zero, a monotone integer):
// the two smallest with their positions, the third as a code only.  Branch-free insertion; a candidate that is met
// again (the seed comes by in the stream, the odd tail of the stream repeats its last candidate) is not inserted twice
// while it sits in one of the two position slots - and if it has dropped to the third slot, inserting its code again
// changes nothing.  The EXACT winner (fp64 compare, ties by original index) is kept by nn_fold as before; because the
// code is monotone, the winner's code is the smallest, so after the stream it sits in slot 1 - or in slot 2 / beyond
// when other candidates share its code, in which case c3 (or c1) has already come down to that code.
struct Top3 {
    int c1, c2, c3, p1, p2;
};
FICP_HD Top3 top3_empty() { return Top3{kHiInf, kHiInf, kHiInf, -1, -1}; }
template <class Acc>
FICP_HD void nn_fold_track(const Acc& acc, int j, double d2, double& best, int& bestpos, Top3& t) {
    const bool again = (j == t.p1) || (j == t.p2);
    const int c = again ? kHiInf : d_hi(d2);
    const bool m1 = c < t.c1, m2 = c < t.c2, m3 = c < t.c3;
    t.c3 = m2 ? t.c2 : (m3 ? c : t.c3);
    t.c2 = m1 ? t.c1 : (m2 ? c : t.c2);
    t.p2 = m1 ? t.p1 : (m2 ? j : t.p2);
    t.c1 = m1 ? c : t.c1;
    t.p1 = m1 ? j : t.p1;
    nn_fold(acc, j, d2, best, bestpos);
}
#if !defined(FICP_TIETEST_STREAM)
#define FICP_TIEFREE_STREAM 1
#endif
// DEFAULT since round 2 (+6 % on C3, 113 GPU parity tests green on it; -DFICP_TIETEST_STREAM restores the per-candidate tie
// test): the same insertion, but the winner is
// kept with a plain `<` (first met wins) - no tie test, no branch in the candidate loop.  An exact tie with a different
// point shows up afterwards as a second candidate carrying the winner's CODE; only then (rare: codes agree to 2^-20) the
// stream is looked at again with the index rule (see nn_search_block3_impl).  Accessors whose admit() can refuse a
// candidate must not use it (TRACK is instantiated for WindowAcc / GlobalAcc only).
FICP_HD void nn_fold_track_notie(int j, double d2, double& best, int& bestpos, Top3& t) {
    const bool again = (j == t.p1) || (j == t.p2);
    const int c = again ? kHiInf : d_hi(d2);
    const bool m1 = c < t.c1, m2 = c < t.c2, m3 = c < t.c3;
    t.c3 = m2 ? t.c2 : (m3 ? c : t.c3);
    t.c2 = m1 ? t.c1 : (m2 ? c : t.c2);
    t.p2 = m1 ? t.p1 : (m2 ? j : t.p2);
    t.c1 = m1 ? c : t.c1;
    t.p1 = m1 ? j : t.p1;
    const bool lt = d2 < best;
    best = lt ? d2 : best;
    bestpos = lt ? j : bestpos;
}
// Runner-up and the code of a lower bound on every streamed candidate other than winner and runner-up.
FICP_HD int top3_finish(const Top3& t, int bestpos, int& pos2) {
    if (t.p1 == bestpos) { pos2 = t.p2; return t.c3; }
    if (t.p2 == bestpos) { pos2 = t.p1; return t.c3; }   // shares its code with p1
    pos2 = -1;                                            // >= 3 candidates share the winner's code, or it was not streamed
    return t.c1;
}

// ---- one candidate ---------------------------------------------------------------------------------
template <bool Z3, class Acc>
FICP_HD void nn_eval(const Acc& acc, int j, double qx, double qy, double qz, double& best, int& bestpos) {
    double tx, ty, tz = 0.0;
    acc.template load<Z3>(j, tx, ty, tz);
    const double dx = dsub(qx, tx);
    const double dy = dsub(qy, ty);
    double d2 = dadd(dmul(dx, dx), dmul(dy, dy));
    if (Z3) {
        const double dz = dsub(qz, tz);
        d2 = dadd(d2, dmul(dz, dz));
    }
    nn_fold(acc, j, d2, best, bestpos);
}

// squared distance only (used by the two-at-a-time stream loop)
template <bool Z3, class Acc>
FICP_HD double nn_dist2(const Acc& acc, int j, double qx, double qy, double qz) {
    double tx, ty, tz = 0.0;
    acc.template load<Z3>(j, tx, ty, tz);
    const double dx = dsub(qx, tx);
    const double dy = dsub(qy, ty);
    double d2 = dadd(dmul(dx, dx), dmul(dy, dy));
    if (Z3) {
        const double dz = dsub(qz, tz);
        d2 = dadd(d2, dmul(dz, dz));
    }
    return d2;
}

template <bool Z3, class Acc>
FICP_HD void nn_scan_segment(const Acc& acc, int y, int xa, int xb, double qx, double qy, double qz,
                             double& best, int& bestpos) {
    int s, e;
    acc.seg(y, xa, xb, s, e);
    for (int j = s; j < e; ++j) nn_eval<Z3>(acc, j, qx, qy, qz, best, bestpos);
}

template <bool Z3, class Acc>
FICP_HD void nn_try_segment(const Acc& acc, const GridGeom& g, int y, int xa, int xb, double qx, double qy,
                            double qz, double& best, int& bestpos) {
    // distance from the query to the (slightly inflated) box of the segment; skip when it cannot
    // hold a point at distance <= best (ties must still be seen for the lowest-index rule)
    double bx0 = g.x0 + xa * g.h - g.eps, bx1 = g.x0 + (xb + 1) * g.h + g.eps;
    double by0 = g.y0 + y * g.h - g.eps, by1 = g.y0 + (y + 1) * g.h + g.eps;
    if (g.clamped) {   // border cells hold the points clamped in from outside the grid: no bound outward
        if (xa == 0) bx0 = -kInf;
        if (xb == g.gw - 1) bx1 = kInf;
        if (y == 0) by0 = -kInf;
        if (y == g.gh - 1) by1 = kInf;
    }
    const double dx = fmax(fmax(bx0 - qx, qx - bx1), 0.0);
    const double dy = fmax(fmax(by0 - qy, qy - by1), 0.0);
    if (dx * dx + dy * dy > best) return;
    nn_scan_segment<Z3>(acc, y, xa, xb, qx, qy, qz, best, bestpos);
}

// Squared lower bound on the distance from the query to ANY point outside the block of cells of Chebyshev
// radius `rad` around (cx, cy); +inf when the block already spans the whole grid.  A point beyond the block's
// left/right side is at least that side's distance away in x AND at least the query's gap to the points'
// y-extent away in y (every point lies inside the true bounding box), and vice versa - this keeps searches of queries far
// off the map (a start pose thrown off the stand) from walking the whole grid.
FICP_HD double nn_block_bound2(const GridGeom& g, double qx, double qy, int cx, int cy, int rad) {
    const int xl = cx - rad, xh = cx + rad, yl = cy - rad, yh = cy + rad;
    const double ox = fmax(fmax(g.tx0 - qx, qx - g.tx1) - g.eps, 0.0);  // gap to the x-extent of the points (true bbox)
    const double oy = fmax(fmax(g.ty0 - qy, qy - g.ty1) - g.eps, 0.0);
    double b2 = kInf;
    if (xl > 0) { const double b = fmax(qx - (g.x0 + xl * g.h) - g.eps, 0.0); b2 = fmin(b2, b * b + oy * oy); }
    if (xh < g.gw - 1) { const double b = fmax((g.x0 + (xh + 1) * g.h) - qx - g.eps, 0.0); b2 = fmin(b2, b * b + oy * oy); }
    if (yl > 0) { const double b = fmax(qy - (g.y0 + yl * g.h) - g.eps, 0.0); b2 = fmin(b2, b * b + ox * ox); }
    if (yh < g.gh - 1) { const double b = fmax((g.y0 + (yh + 1) * g.h) - qy - g.eps, 0.0); b2 = fmin(b2, b * b + ox * ox); }
    return b2;
}

// Rings r_start, r_start+1, ... around cell (cx, cy); every cell of Chebyshev radius < r_start has been
// visited already.  Stops when the best distance is strictly below the distance to the border of the
// visited block (so no unvisited point can be closer or tie).  Returns false on a window miss.
template <bool Z3, class Acc>
FICP_HD bool nn_ring_loop_impl(const Acc& acc, const GridGeom& g, double qx, double qy, double qz, int cx, int cy,
                               int r_start, double& best, int& bestpos) {
    for (int r = r_start;; ++r) {
        const double b2 = nn_block_bound2(g, qx, qy, cx, cy, r - 1);
        if (b2 == kInf) break;        // the block already spans the whole grid
        if (best < b2) break;         // strict: an unvisited point can neither beat nor tie the best
        const int nxl = (cx - r > 0) ? cx - r : 0;
        const int nxh = (cx + r < g.gw - 1) ? cx + r : g.gw - 1;
        const int nyl = (cy - r > 0) ? cy - r : 0;
        const int nyh = (cy + r < g.gh - 1) ? cy + r : g.gh - 1;
        if (!acc.covers(nxl, nxh, nyl, nyh)) return false;
        for (int y = nyl; y <= nyh; ++y) {
            if (y == cy - r || y == cy + r) {
                nn_try_segment<Z3>(acc, g, y, nxl, nxh, qx, qy, qz, best, bestpos);
            } else {
                if (cx - r >= 0) nn_try_segment<Z3>(acc, g, y, cx - r, cx - r, qx, qy, qz, best, bestpos);
                if (cx + r <= g.gw - 1) nn_try_segment<Z3>(acc, g, y, cx + r, cx + r, qx, qy, qz, best, bestpos);
            }
        }
    }
    return true;
}

// Quick exit when the visited block of radius r_start-1 already bounds the search (the common case), else the
// out-of-line ring loop.
struct NNState { double best; int pos; int ok; };
template <bool Z3, class Acc>
FICP_COLD_RING NNState nn_ring_loop_cold(const Acc& acc, const GridGeom& g, double qx, double qy, double qz, int cx, int cy,
                                    int r_start, double best, int bestpos) {
    NNState st;
    st.ok = nn_ring_loop_impl<Z3>(acc, g, qx, qy, qz, cx, cy, r_start, best, bestpos) ? 1 : 0;
    st.best = best;
    st.pos = bestpos;
    return st;
}
template <bool Z3, class Acc>
FICP_HD bool nn_ring_loop(const Acc& acc, const GridGeom& g, double qx, double qy, double qz, int cx, int cy,
                          int r_start, double& best, int& bestpos) {
    // cheap form of the bound first (side distances only; it never exceeds the exact block bound, so exiting on
    // it is safe) - this is the per-query common case
    {
        const int rad = r_start - 1;
        const int xl = cx - rad, xh = cx + rad, yl = cy - rad, yh = cy + rad;
        double b = kInf;
        if (xl > 0) b = fmin(b, qx - (g.x0 + xl * g.h));
        if (xh < g.gw - 1) b = fmin(b, (g.x0 + (xh + 1) * g.h) - qx);
        if (yl > 0) b = fmin(b, qy - (g.y0 + yl * g.h));
        if (yh < g.gh - 1) b = fmin(b, (g.y0 + (yh + 1) * g.h) - qy);
        if (b == kInf) return true;
        b -= g.eps;
        if (b > 0.0 && best < b * b) return true;
    }
    const NNState st = nn_ring_loop_cold<Z3>(acc, g, qx, qy, qz, cx, cy, r_start, best, bestpos);
    best = st.best;
    bestpos = st.pos;
    return st.ok != 0;
}

// Reference form of the search (ring by ring from the query's own cell).  Returns false when the search
// needs cells the accessor does not cover (window miss): the caller then repeats the query with
// GlobalAcc.  On success best = squared distance (canonical arithmetic) and bestpos = accessor-local
// position of the winner.
template <bool Z3, class Acc>
FICP_HD bool nn_search(const Acc& acc, const GridGeom& g, double qx, double qy, double qz, double& best,
                       int& bestpos) {
    const int cx = clamp_cell((qx - g.x0) * g.inv_h, g.gw);
    const int cy = clamp_cell((qy - g.y0) * g.inv_h, g.gh);
    best = kInf;
    bestpos = -1;
    if (!acc.covers(cx, cx, cy, cy)) return false;
    nn_scan_segment<Z3>(acc, cy, cx, cx, qx, qy, qz, best, bestpos);
    return nn_ring_loop<Z3>(acc, g, qx, qy, qz, cx, cy, 1, best, bestpos);
}

// Production form used by the kernels: the 3x3 block around the query's cell is visited as ONE flat
// stream of candidates, so that the 32 lanes of a warp (one query each, in lock-step) pay the largest
// TOTAL candidate count among them instead of the largest count of every cell separately.
//   1. `prev` (the neighbour found by the previous ICP pass, or -1) is scored first: it bounds the
//      search from the start (temporal coherence; the result is still the exact NN);
//   2. per row of the block, the run of cells whose box lies within that bound is one contiguous range
//      of the cell-sorted array: three (start, end) pairs;
//   3. one loop over the concatenation of the three ranges;
//   4. if the border of the block is not provably farther than the best distance, the ring loop goes on
//      from radius 2 (rare).
// TRACK: also return the runner-up `pos2` (accessor-local position, -1: none) and `lb_hi` = code (d_hi, rounded down)
// of a lower bound on the squared distance from the query to every target point of the 3x3 block OTHER than the winner
// and the runner-up: the third-best candidate streamed, and the (inflated) boxes of the block's cells that were
// pruned.  Together with the distance to the block's border (see nn_block_border2) it bounds every other point of the
// target - what the ICP kernel needs to prove, on later passes, that a query that moved by less than the slack still
// has one of these two as its nearest neighbour.
// Squared gaps between the query and the three cell columns / rows of the 3x3 block around its (clamped) cell, boxes
// inflated by eps.  u = offset of the query from the lower-left corner of its cell; it lies in [0, h] unless the query is
// off the grid.  With clamped targets a border cell has no outer side: a query beyond it has no gap to it.
FICP_HD void nn_block3_gaps(const GridGeom& g, double qx, double qy, int cx, int cy, double (&gx)[3], double (&gy)[3]) {
    const double h = g.h, eps = g.eps;
    const double ux = qx - (g.x0 + cx * h), uy = qy - (g.y0 + cy * h);
    gx[0] = fmax(ux - eps, 0.0);
    gx[1] = fmax(fmax(-ux, ux - h) - eps, 0.0);
    gx[2] = fmax(h - ux - eps, 0.0);
    gy[0] = fmax(uy - eps, 0.0);
    gy[1] = fmax(fmax(-uy, uy - h) - eps, 0.0);
    gy[2] = fmax(h - uy - eps, 0.0);
    if (g.clamped) {
        if ((cx == 0 && ux < 0.0) || (cx == g.gw - 1 && ux > h)) gx[1] = 0.0;
        if ((cy == 0 && uy < 0.0) || (cy == g.gh - 1 && uy > h)) gy[1] = 0.0;
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) { gx[i] *= gx[i]; gy[i] *= gy[i]; }
}

template <bool Z3, bool TRACK, class Acc>
FICP_HD bool nn_search_block3_impl(const Acc& acc, const GridGeom& g, double qx, double qy, double qz, int prev,
                                   double& best, int& bestpos, int& cx, int& cy, int& lb_hi, int& pos2) {
    cx = clamp_cell((qx - g.x0) * g.inv_h, g.gw);
    cy = clamp_cell((qy - g.y0) * g.inv_h, g.gh);
    const int xl = (cx > 0) ? cx - 1 : 0, xh = (cx < g.gw - 1) ? cx + 1 : g.gw - 1;
    const int yl = (cy > 0) ? cy - 1 : 0, yh = (cy < g.gh - 1) ? cy + 1 : g.gh - 1;
    if (!acc.covers(xl, xh, yl, yh)) return false;
    best = kInf;
    bestpos = -1;
    int lb = kHiInf;
    Top3 top = top3_empty();
    pos2 = -1;
    if (prev >= 0) nn_eval<Z3>(acc, prev, qx, qy, qz, best, bestpos);
#if defined(FICP_PRESCAN_OWN_CELL)
    // EXPERIMENT, off by default (not measured on the GPU yet): a search without a seed (first pass
    // of a hypothesis) looks at the query's own cell first, so that the pruning below has a bound to work with instead
    // of streaming all nine cells; the own cell is streamed again with the others (same candidates, same result).
    else if (TRACK) nn_scan_segment<Z3>(acc, cy, cx, cx, qx, qy, qz, best, bestpos);
#endif
    double gx[3], gy[3];
    nn_block3_gaps(g, qx, qy, cx, cy, gx, gy);
    // TRACK: cells a little beyond the seed's distance are streamed as well - every cell pruned caps the lower bound
    // (and with it how long the query can skip its searches) at the cell's gap, which would be barely above `best`
#ifndef FICP_PRUNE_PAD
#define FICP_PRUNE_PAD 1.5
#endif
    const double bound = TRACK ? best * FICP_PRUNE_PAD : best;
    int s[3], n[3];
#pragma unroll
    for (int ry = 0; ry < 3; ++ry) {
        const int y = cy - 1 + ry;
        s[ry] = 0;
        n[ry] = 0;
        if (y < yl || y > yh) continue;
        // columns of this row whose box can hold a point at distance <= best (a contiguous run)
        int xa = cx + 2, xb = cx - 2;
#pragma unroll
        for (int rx = 0; rx < 3; ++rx) {
            const int x = cx - 1 + rx;
            if (x >= xl && x <= xh) {
                const double gap2 = gx[rx] + gy[ry];
                if (gap2 <= bound) {
                    if (x < xa) xa = x;
                    xb = x;
                } else if (TRACK) {
                    const int c = d_hi(gap2);
                    lb = (c < lb) ? c : lb;
                }
            }
        }
        if (xa <= xb) {
            int e;
            acc.seg(y, xa, xb, s[ry], e);
            n[ry] = e - s[ry];
        }
    }
    // flat index t -> position: t + (offset of the run t falls in); two candidates per iteration so that two
    // independent load/arithmetic chains are in flight (the fold into `best` stays in stream order)
    const int n01 = n[0] + n[1], total = n01 + n[2];
    const int o1 = s[1] - n[0], o2 = s[2] - n01;
    for (int t = 0; t < total; t += 2) {
        const int t1 = (t + 1 < total) ? t + 1 : t;
        const int j0 = t + ((t < n[0]) ? s[0] : (t < n01) ? o1 : o2);
        const int j1 = t1 + ((t1 < n[0]) ? s[0] : (t1 < n01) ? o1 : o2);
        const double da = nn_dist2<Z3>(acc, j0, qx, qy, qz);
        const double db = nn_dist2<Z3>(acc, j1, qx, qy, qz);
        if (TRACK) {
#if defined(FICP_TIEFREE_STREAM)
            nn_fold_track_notie(j0, da, best, bestpos, top);
            nn_fold_track_notie(j1, db, best, bestpos, top);
#else
            nn_fold_track(acc, j0, da, best, bestpos, top);
            nn_fold_track(acc, j1, db, best, bestpos, top);
#endif
        } else {
            nn_fold(acc, j0, da, best, bestpos);
            nn_fold(acc, j1, db, best, bestpos);
        }
    }
#if defined(FICP_TIEFREE_STREAM)
    if (TRACK) {
        // a candidate other than the winner carries the winner's code: possibly an exact tie - settle it by original
        // index in a second look at the stream (the winner's distance is already the minimum, only its index can change)
        const int cb = d_hi(best);
        if (top.c1 == cb && (top.p1 != bestpos || top.c2 == cb)) {
            for (int t = 0; t < total; ++t) {
                const int j = t + ((t < n[0]) ? s[0] : (t < n01) ? o1 : o2);
                nn_eval<Z3>(acc, j, qx, qy, qz, best, bestpos);
            }
        }
    }
#endif
    if (TRACK) {
        const int c = top3_finish(top, bestpos, pos2);
        lb = (c < lb) ? c : lb;
    }
    lb_hi = lb;
    return true;
}
template <bool Z3, class Acc>
FICP_HD bool nn_search_block3(const Acc& acc, const GridGeom& g, double qx, double qy, double qz, int prev,
                              double& best, int& bestpos, int& cx, int& cy) {
    int lb_hi, pos2;
    return nn_search_block3_impl<Z3, false>(acc, g, qx, qy, qz, prev, best, bestpos, cx, cy, lb_hi, pos2);
}

// Unseeded form for bulk queries (nn_bulk.cu): the query's OWN cell is scored first, so that the pruning of the other
// eight cells has a bound to work with; then the surviving cells of the 3x3 block follow as at most four runs of
// consecutive cells (every point is visited once).  The winner is kept with a plain `<`; an exact tie between two points
// raises a flag and only then (rare) the candidates are looked at again with the lowest-original-index rule.  Same result as nn_search_block3: exact minimum over the block, ties to the lowest index.
template <bool Z3, class Acc>
FICP_HD void nn_scan_run_flag(const Acc& acc, int s, int e, double qx, double qy, double qz, double& best, int& bestpos, bool& tie) {
#pragma unroll 2
    for (int j = s; j < e; ++j) {
        const double d = nn_dist2<Z3>(acc, j, qx, qy, qz);
        tie = tie || (d == best);        // an exact tie between two different points (no point is visited twice)
        const bool lt = d < best;
        best = lt ? d : best;
        bestpos = lt ? j : bestpos;
    }
}
template <bool Z3, class Acc>
FICP_HD void nn_search_block3_unseeded(const Acc& acc, const GridGeom& g, double qx, double qy, double qz, int cx, int cy,
                                       double& best, int& bestpos) {
    const int xl = (cx > 0) ? cx - 1 : 0, xh = (cx < g.gw - 1) ? cx + 1 : g.gw - 1;
    const int yl = (cy > 0) ? cy - 1 : 0, yh = (cy < g.gh - 1) ? cy + 1 : g.gh - 1;
    best = kInf;
    bestpos = -1;
    bool tie = false;
    int s0, e0;
    acc.seg(cy, cx, cx, s0, e0);
    nn_scan_run_flag<Z3>(acc, s0, e0, qx, qy, qz, best, bestpos, tie);
    double gx[3], gy[3];
    nn_block3_gaps(g, qx, qy, cx, cy, gx, gy);
    const double bound = best;
    // the other eight cells as (at most) four runs of consecutive cells: row below, left and right neighbour, row above
    int rs[4], re[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { rs[k] = 0; re[k] = 0; }
#pragma unroll
    for (int ry = 0; ry < 3; ry += 2) {
        const int y = cy - 1 + ry;
        if (y < yl || y > yh) continue;
        int xa = cx + 2, xb = cx - 2;
#pragma unroll
        for (int rx = 0; rx < 3; ++rx) {
            const int x = cx - 1 + rx;
            if (x >= xl && x <= xh && gx[rx] + gy[ry] <= bound) {
                if (x < xa) xa = x;
                xb = x;
            }
        }
        if (xa <= xb) acc.seg(y, xa, xb, rs[ry ? 3 : 0], re[ry ? 3 : 0]);
    }
    if (cx - 1 >= xl && gx[0] + gy[1] <= bound) acc.seg(cy, cx - 1, cx - 1, rs[1], re[1]);
    if (cx + 1 <= xh && gx[2] + gy[1] <= bound) acc.seg(cy, cx + 1, cx + 1, rs[2], re[2]);
#pragma unroll
    for (int k = 0; k < 4; ++k) nn_scan_run_flag<Z3>(acc, rs[k], re[k], qx, qy, qz, best, bestpos, tie);
    if (tie) {   // rare: second look with the lowest-original-index rule
        for (int j = s0; j < e0; ++j) nn_eval<Z3>(acc, j, qx, qy, qz, best, bestpos);
#pragma unroll
        for (int k = 0; k < 4; ++k)
            for (int j = rs[k]; j < re[k]; ++j) nn_eval<Z3>(acc, j, qx, qy, qz, best, bestpos);
    }
}

// True when the visited block of Chebyshev radius `rad` around (cx, cy) provably bounds the search (cheap form of
// the bound: side distances only; it never exceeds nn_block_bound2, so stopping on it is safe).
FICP_HD bool nn_block_settles(const GridGeom& g, double qx, double qy, int cx, int cy, int rad, double best) {
    const int xl = cx - rad, xh = cx + rad, yl = cy - rad, yh = cy + rad;
    double b = kInf;
    if (xl > 0) b = fmin(b, qx - (g.x0 + xl * g.h));
    if (xh < g.gw - 1) b = fmin(b, (g.x0 + (xh + 1) * g.h) - qx);
    if (yl > 0) b = fmin(b, qy - (g.y0 + yl * g.h));
    if (yh < g.gh - 1) b = fmin(b, (g.y0 + (yh + 1) * g.h) - qy);
    if (b == kInf) return true;
    b -= g.eps;
    return b > 0.0 && best < b * b;
}

// The same bound as a number: squared distance below which the visited block provably holds the nearest neighbour
// (+inf: the block spans the whole grid; 0: nothing can be said).  nn_block_settles(best) == (best < border2).
FICP_HD double nn_block_border2(const GridGeom& g, double qx, double qy, int cx, int cy, int rad) {
    const int xl = cx - rad, xh = cx + rad, yl = cy - rad, yh = cy + rad;
    double b = kInf;
    if (xl > 0) b = fmin(b, qx - (g.x0 + xl * g.h));
    if (xh < g.gw - 1) b = fmin(b, (g.x0 + (xh + 1) * g.h) - qx);
    if (yl > 0) b = fmin(b, qy - (g.y0 + yl * g.h));
    if (yh < g.gh - 1) b = fmin(b, (g.y0 + (yh + 1) * g.h) - qy);
    if (b == kInf) return kInf;
    b -= g.eps;
    return (b > 0.0) ? b * b : 0.0;
}

template <bool Z3, class Acc>
FICP_HD bool nn_search_stream(const Acc& acc, const GridGeom& g, double qx, double qy, double qz, int prev,
                              double& best, int& bestpos) {
    int cx, cy;
    if (!nn_search_block3<Z3>(acc, g, qx, qy, qz, prev, best, bestpos, cx, cy)) return false;
    return nn_ring_loop<Z3>(acc, g, qx, qy, qz, cx, cy, 2, best, bestpos);
}

}  // namespace ficp
