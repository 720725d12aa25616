// Device code shared by the two persistent Fractional-ICP kernels:
//   icp_persistent.cu  one WARP per (plot, hypothesis) ICP - throughput shape, batches that fill the machine
//   icp_team.cu        one CTA  per (plot, hypothesis) ICP - latency shape, batches smaller than the machine
// Both must return the SAME BITS for the same ICP (tests/test_gpu_icp.py::test_cta_per_icp_is_bit_identical), so every
// piece of arithmetic whose rounding reaches a result lives here, written with explicit rounding intrinsics (no
// compiler-chosen FMA contraction): the start pose / pose application, the canonical squared distance (nn_search.cuh),
// and the rigid fit (fit_shift / fit_term / fit_reduce / fit_solve, ficp.py:89-110).  The skip test and the search
// bookkeeping (runner-up, slack) decide only WHETHER a query is searched again, never its result, so they may differ.
#pragma once
#include <cuda_fp16.h>
#include "ficp_internal.h"
#include "nn_search.cuh"

namespace ficp {
namespace {

constexpr unsigned kFull = 0xFFFFFFFFu;

struct Pose {  // q = M u + c ; warp-uniform
    double m00, m01, m10, m11, cx, cy;
};


__device__ __forceinline__ bool key_greater(double da, unsigned ka, double db, unsigned kb) {
    return (da > db) || (da == db && ka > kb);
}


struct PassOut {
    int k;          // trimmed subset size (0: none, like ficp.py:125)
    double f;       // FRMSD at k
    double rmse;    // sqrt(S_k / k)
    double thr;     // d2 of the k-th point in trim order
    int thr_idx;    // its source index (ties in d2 are ordered by index)
};

struct PlotCtx {
    const double2* s_u;
    const double* s_z;
    int n;
    int fixed_k;
    double ubx, uby;
};

__device__ __forceinline__ void pose_apply(const Pose& P, const double2 u, double& qx, double& qy) {
    // same expression, same order, no FMA, as oracle.pre_transform
    qx = dadd(dadd(dmul(P.m00, u.x), dmul(P.m01, u.y)), P.cx);
    qy = dadd(dadd(dmul(P.m10, u.x), dmul(P.m11, u.y)), P.cy);
}

// Global-grid form of the query (window miss): rare, so kept out of line to keep the hot loop small.
// measured (profiles/r01_variants.md): inlining the rare global-grid / tie / ring paths beats calling them
#if defined(FICP_NOINLINE_GLOBAL)
#define FICP_GLOBAL_ATTR __device__ __noinline__
#else
#define FICP_GLOBAL_ATTR __device__ __forceinline__
#endif
template <bool Z3>
FICP_GLOBAL_ATTR int nn_query_global(const GridView& G, double qx, double qy, double qz, int prev, double* best_out) {
    const GlobalAcc ga = make_global_acc(G);
    double best;
    int pos;
    nn_search_stream<Z3>(ga, G.g, qx, qy, qz, prev, best, pos);
    *best_out = best;
    return pos;
}

// Neighbour code of a query (`snn`): -1 = none; bit 31 set = position in the GLOBAL cell-sorted target (the query ran
// on the global grid); otherwise two window-local positions (< 32768): bits 0-15 the nearest neighbour, bits 16-30 the
// runner-up of its last search (== the neighbour: none).
__device__ __forceinline__ int code_pack(int pos, int pos2) {
    return (pos < 0) ? -1 : (pos | (((pos2 < 0) ? pos : pos2) << 16));
}
__device__ __forceinline__ int code_win(int code) { return (code < 0) ? -1 : (code & 0xFFFF); }

__device__ __forceinline__ int ld_volatile(const int* p) { return *reinterpret_cast<const volatile int*>(p); }

// A warp-uniform view of a shared word that other warps change: ONE lane reads, everybody gets that value.  (Every lane
// reading for itself can split the warp on a loop or branch condition when the lanes are not converged at the load.)
__device__ __forceinline__ int ld_volatile_uniform(const int* p, int lane) {
    __syncwarp();
    int v = 0;
    if (lane == 0) v = ld_volatile(p);
    return __shfl_sync(kFull, v, 0);
}

// 32 entries of the deferred list: rings >= 2 inside the window, or the whole query on the global grid.
template <bool Z3, bool ELASTIC>
__device__ __forceinline__ void nn_deferred_chunk(const GridView& G, const WindowAcc& W, const PlotCtx& pc, const Pose& P,
                                                  double* __restrict__ sd2, int* __restrict__ snn,
                                                  const unsigned short* __restrict__ sord, int base, int n_def,
                                                  int lane, unsigned& n_global) {
    if (base + lane < n_def) {
        const int d = sord[base + lane];
        const int i = d & 0x7FFF;
        FICP_ASSERT(i < pc.n);
        double qx, qy;
        pose_apply(P, pc.s_u[i], qx, qy);
        const double qz = Z3 ? pc.s_z[i] : 0.0;
        double best = kInf;
        int pos = -1;
        bool ok = false;
        if (!(d & 0x8000)) {
            best = ELASTIC ? fabs(sd2[i]) : sd2[i];
            pos = code_win(snn[i]);
            const int cx = clamp_cell((qx - G.g.x0) * G.g.inv_h, G.g.gw);
            const int cy = clamp_cell((qy - G.g.y0) * G.g.inv_h, G.g.gh);
            ok = nn_ring_loop_impl<Z3>(W, G.g, qx, qy, qz, cx, cy, 2, best, pos);
        }
        int code = code_pack(pos, -1);
        if (!ok) {
            // window miss: whole query on the global grid, seeded with the best candidate known so far
            const int seed = snn[i];
            const int gprev = (seed == -1) ? -1 : (seed >= 0 ? W.global_pos(seed & 0xFFFF) : (seed & 0x7FFFFFFF));
            pos = nn_query_global<Z3>(G, qx, qy, qz, gprev, &best);
            code = (int)((unsigned)pos | 0x80000000u);
            ++n_global;
        }
        sd2[i] = best;
        snn[i] = code;
    }
}

// A query whose 3x3 block was searched inside the window: store neighbour code, squared distance and the slack of the
// skip test; returns -1, or the tree index when the block does not settle the search (rings >= 2 follow, deferred).
template <bool MARK>
__device__ __forceinline__ int nn_finish_window(const GridGeom& g, int i, double qx, double qy, int cx, int cy, double best,
                                                int pos, int pos2, int lb_hi, double* __restrict__ sd2,
                                                int* __restrict__ snn, __half* __restrict__ ssl) {
    int defer = -1;
    float slack = 0.f;  // deferred queries carry no bound: they are searched again next pass
    snn[i] = code_pack(pos, pos2);
    const double border2 = nn_block_border2(g, qx, qy, cx, cy, 1);
    if ((border2 == kInf) || best < border2) {   // == nn_block_settles
        // every target point other than the winner and the runner-up is at least sqrt(lb2) away (third-best
        // streamed, pruned cells of the block, the block's border): rounded DOWN at every step
        const double lb2 = fmin(hi_to_double(lb_hi), border2);
        slack = fminf(__fsqrt_rd(__double2float_rd(lb2)), 60000.f);
    } else {
        defer = i;
        if (MARK) best = -best;
    }
    sd2[i] = best;
    ssl[i] = __float2half_rd(slack);
    return defer;
}

// One round of nearest-neighbour queries: the 32 queries at positions 32e..32e+31 of the PREVIOUS pass's trim order
// (identity on the first pass): neighbours in that order have similar residuals, hence similar search radii and
// candidate counts, so the lanes of a warp finish together and the rare wide searches (ring >= 2) fall into the same
// rounds.  Queries whose 3x3 block does not settle the search (wide search radius, or the block is not inside the
// shared-memory window) are NOT finished inline - a handful of lanes would drag the whole warp through the ring loop
// in almost every round - they are returned as `defer` (point index | 0x8000 if it must run on the global grid) and
// finished afterwards with all lanes busy.  MARK: also flag them in the sign bit of sd2 (-best: in-window candidate
// known, -inf: nothing known) for rounds that complete out of order.
template <bool Z3, bool MARK>
__device__ __forceinline__ int nn_round(const GridView& G, const WindowAcc& W, bool win_ok, const PlotCtx& pc,
                                        const Pose& P, double* __restrict__ sd2, int* __restrict__ snn,
                                        __half* __restrict__ ssl, const unsigned short* __restrict__ sord, int e,
                                        int count, int lane, bool have_prev) {
    const int p = e * 32 + lane;
    int defer = -1;
    if (p < count) {
        const int i = sord[p];
        FICP_ASSERT(i >= 0 && i < pc.n);
        double qx, qy;
        pose_apply(P, pc.s_u[i], qx, qy);
        const double qz = Z3 ? pc.s_z[i] : 0.0;
        // seed with the neighbour found by the previous pass of this hypothesis (same index space only)
        const int pc_prev = have_prev ? code_win(snn[i]) : -1;
        double best = kInf;
        int pos = -1, cx, cy;
        bool ok = false;
        int lb_hi = kHiInf, pos2 = -1;
        if (win_ok) ok = nn_search_block3_impl<Z3, true>(W, G.g, qx, qy, qz, pc_prev, best, pos, cx, cy, lb_hi, pos2);
        if (ok) {
            defer = nn_finish_window<MARK>(G.g, i, qx, qy, cx, cy, best, pos, pos2, lb_hi, sd2, snn, ssl);
        } else {
            if (!have_prev) snn[i] = -1;  // keep the previous pass's code as the seed of the deferred query
            if (MARK) sd2[i] = -kInf;
            defer = i | 0x8000;
            ssl[i] = __float2half_rd(0.f);  // deferred / global-grid queries carry no bound: they are searched again next pass
        }
    }
    return defer;
}

// ---- group search (CTA-per-ICP kernel): G = 2..16 lanes share ONE query and split its candidate stream ---------------
// The one-lane search above walks ~15 (up to ~50) candidates serially; when a pass searches only a few dozen queries a
// CTA has lanes to spare.  Lanes sub = 0..G-1 of a group compute the same set-up (cell, gaps, row runs), take the
// candidates t = sub, sub + G, ... of the flat stream and merge their (winner, runner-up, third-best code) with xor
// shuffles.  The winner is exact (canonical d2, strict <); a POSSIBLE exact tie - another candidate carrying the
// winner's truncated code - is not resolved here: status 2 sends the query to the one-lane search, which applies the
// lowest-original-index rule.  The seed (previous neighbour) only bounds the pruning; it is folded in by lane 0 only
// when the stream will not meet it.  Returns 0 ok, 1 window miss (global grid), 2 possible tie (one-lane search).
__device__ __forceinline__ void top3_insert(Top3& t, int c, int j) {
    const bool m1 = c < t.c1, m2 = c < t.c2, m3 = c < t.c3;
    t.c3 = m2 ? t.c2 : (m3 ? c : t.c3);
    t.c2 = m1 ? t.c1 : (m2 ? c : t.c2);
    t.p2 = m1 ? t.p1 : (m2 ? j : t.p2);
    t.c1 = m1 ? c : t.c1;
    t.p1 = m1 ? j : t.p1;
}
template <bool Z3>
__device__ __forceinline__ int nn_search_group(const WindowAcc& acc, const GridGeom& g, bool active, double qx, double qy,
                                               double qz, int prev, int G, int sub, double& best, int& bestpos, int& cx,
                                               int& cy, int& lb_hi, int& pos2) {
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
            int s[3], n[3];
            bool seed_in = false;
#pragma unroll
            for (int ry = 0; ry < 3; ++ry) {
                const int y = cy - 1 + ry;
                s[ry] = 0;
                n[ry] = 0;
                if (y < yl || y > yh) continue;
                int xa = cx + 2, xb = cx - 2;
#pragma unroll
                for (int rx = 0; rx < 3; ++rx) {
                    const int x = cx - 1 + rx;
                    if (x >= xl && x <= xh) {
                        const double gap2 = gx[rx] + gy[ry];
                        if (gap2 <= bound) {
                            if (x < xa) xa = x;
                            xb = x;
                        } else {
                            const int c = d_hi(gap2);
                            lb = (c < lb) ? c : lb;
                        }
                    }
                }
                if (xa <= xb) {
                    int e;
                    acc.seg(y, xa, xb, s[ry], e);
                    n[ry] = e - s[ry];
                    seed_in = seed_in || (prev >= s[ry] && prev < e);
                }
            }
            if (prev >= 0 && !seed_in && sub == 0) nn_fold_track_notie(prev, seed_d2, best, bestpos, top);
            const int n01 = n[0] + n[1], total = n01 + n[2];
            const int o1 = s[1] - n[0], o2 = s[2] - n01;
            for (int t = sub; t < total; t += G) {
                const int j = t + ((t < n[0]) ? s[0] : (t < n01) ? o1 : o2);
                const double d = nn_dist2<Z3>(acc, j, qx, qy, qz);
                nn_fold_track_notie(j, d, best, bestpos, top);
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
}

// Skip test of one round of 32 queries (i = 32e + lane), passes after the first.  `D` = pose of this pass minus the
// pose of the previous pass.  The query moved by |D.M u + D.c| (+ rounding of the two positions, `1e-14 |q|` is
// 50x what they can differ by); every point other than the winner and the runner-up of the query's last search is
// therefore still at least slack - move away (triangle inequality; in 3-D too, Z does not move).  If the smaller of
// their two distances - evaluated in the canonical arithmetic, it is the value the full search would return - is
// below that, that point is the unique nearest neighbour: no tie, nothing to search.  All roundings are directed against passing.  Returns the point
// index if the query must be searched, -1 if it is settled.
template <bool Z3>
__device__ __forceinline__ int nn_test_round(const WindowAcc& W, const PlotCtx& pc, const Pose& P, const Pose& D,
                                             double* __restrict__ sd2, int* __restrict__ snn,
                                             __half* __restrict__ ssl, int e, int lane) {
    const int i = e * 32 + lane;
    int need = -1;
    if (i < pc.n) {
        need = i;
        const int code = snn[i];
        const float s0 = __half2float(ssl[i]);
        if (code >= 0 && s0 > 0.f) {
            const double2 u = pc.s_u[i];
            const double ex = D.m00 * u.x + D.m01 * u.y + D.cx;
            const double ey = D.m10 * u.x + D.m11 * u.y + D.cy;
            const double pad = 1e-14 * ((fabs(P.cx) + fabs(P.cy)) + (fabs(u.x) + fabs(u.y)));
            const float move = __fadd_ru(__fsqrt_ru(__double2float_ru(ex * ex + ey * ey)), __double2float_ru(pad));
            const float s1 = __fmul_rd(__fsub_rd(s0, move), 0.99999904632568359375f);  // (1 - 2^-20): rounding of d2
            const __half sh = __float2half_rd(fmaxf(s1, 0.f));
            ssl[i] = sh;
            const float s = __half2float(sh);
            double qx, qy;
            pose_apply(P, u, qx, qy);
            const double qz = Z3 ? pc.s_z[i] : 0.0;
            const int p1 = code & 0xFFFF, p2 = code >> 16;
            const double d1 = nn_dist2<Z3>(W, p1, qx, qy, qz);
            const double dr = nn_dist2<Z3>(W, p2, qx, qy, qz);   // p2 == p1 when there is no runner-up
            const bool swap = dr < d1;
            const double dmin = swap ? dr : d1;
            // an exact tie between the two is left to the search (lowest original index wins there)
            if (dmin < (double)s * (double)s && (p2 == p1 || d1 != dr)) {
                sd2[i] = dmin;
                if (swap) snn[i] = p2 | (p1 << 16);
                need = -1;
            }
        }
    }
    return need;
}


// ---- rigid 2-D fit (ficp.py:89-110), arithmetic shared by both kernels --------------------------------------------
// Running sums about the shift point a = plot centroid under the current pose: sum u, sum v, sum u v^T, and the
// magnitude sum used to recognise an exactly-zero cross-covariance.  Per LANE the terms are added in tree order
// i = lane, lane + 32, ... (fit_term), then the 32 lane sums are combined by an xor butterfly (fit_reduce), which is
// bit-identical in every lane.  The CTA-per-ICP kernel reproduces exactly this order with one warp.
struct FitSums {
    double su0, su1, sv0, sv1, h00, h01, h10, h11, habs;
};
__device__ __forceinline__ FitSums fit_zero() { return FitSums{0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0}; }
__device__ __forceinline__ void fit_shift(const Pose& P, double ubx, double uby, double& ax, double& ay) {
    ax = __dadd_rn(__fma_rn(P.m01, uby, __dmul_rn(P.m00, ubx)), P.cx);
    ay = __dadd_rn(__fma_rn(P.m11, uby, __dmul_rn(P.m10, ubx)), P.cy);
}
// one inlier: source position q (under the current pose) and its correspondence t, both relative to the shift point
__device__ __forceinline__ void fit_uv(double qx, double qy, double tx, double ty, double ax, double ay, double& ux, double& uy,
                                       double& vx, double& vy) {
    ux = __dsub_rn(qx, ax); uy = __dsub_rn(qy, ay); vx = __dsub_rn(tx, ax); vy = __dsub_rn(ty, ay);
}
// the nine running sums of one inlier: the same operations, in the same order per sum, in both kernels
__device__ __forceinline__ void fit_acc(FitSums& s, double ux, double uy, double vx, double vy) {
    s.su0 = __dadd_rn(s.su0, ux); s.su1 = __dadd_rn(s.su1, uy);
    s.sv0 = __dadd_rn(s.sv0, vx); s.sv1 = __dadd_rn(s.sv1, vy);
    s.h00 = __fma_rn(ux, vx, s.h00); s.h01 = __fma_rn(ux, vy, s.h01);
    s.h10 = __fma_rn(uy, vx, s.h10); s.h11 = __fma_rn(uy, vy, s.h11);
    s.habs = __fma_rn(__dadd_rn(fabs(ux), fabs(uy)), __dadd_rn(fabs(vx), fabs(vy)), s.habs);  // noise scale of the terms
}
__device__ __forceinline__ void fit_term(FitSums& s, double qx, double qy, double tx, double ty, double ax, double ay) {
    double ux, uy, vx, vy;
    fit_uv(qx, qy, tx, ty, ax, ay, ux, uy, vx, vy);
    fit_acc(s, ux, uy, vx, vy);
}
__device__ __forceinline__ void fit_reduce(FitSums& s) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s.habs = __dadd_rn(s.habs, __shfl_xor_sync(kFull, s.habs, o));
        s.su0 = __dadd_rn(s.su0, __shfl_xor_sync(kFull, s.su0, o)); s.su1 = __dadd_rn(s.su1, __shfl_xor_sync(kFull, s.su1, o));
        s.sv0 = __dadd_rn(s.sv0, __shfl_xor_sync(kFull, s.sv0, o)); s.sv1 = __dadd_rn(s.sv1, __shfl_xor_sync(kFull, s.sv1, o));
        s.h00 = __dadd_rn(s.h00, __shfl_xor_sync(kFull, s.h00, o)); s.h01 = __dadd_rn(s.h01, __shfl_xor_sync(kFull, s.h01, o));
        s.h10 = __dadd_rn(s.h10, __shfl_xor_sync(kFull, s.h10, o)); s.h11 = __dadd_rn(s.h11, __shfl_xor_sync(kFull, s.h11, o));
    }
}
// Closed-form rotation (+ optional reflection) from the reduced sums and composition into the pose:
//   q' = R (q - a - mu) + a + mv  with q = M u + c   ->   M' = R M ;  c' = R (c - a - mu) + a + mv
// D = new pose - old pose (what the next pass's skip test moves the queries by).
__device__ __forceinline__ void fit_solve(const FitSums& s, int k, int allow_reflection, double ax, double ay, Pose& P, Pose& D) {
    const double inv_k = __drcp_rn((double)k);   // correctly rounded reciprocal == __ddiv_rn(1.0, k), a shorter sequence
    const double mu0 = __dmul_rn(s.su0, inv_k), mu1 = __dmul_rn(s.su1, inv_k);
    const double mv0 = __dmul_rn(s.sv0, inv_k), mv1 = __dmul_rn(s.sv1, inv_k);
    // centred cross-covariance H = sum (u - mu)(v - mv)^T, from the shifted sums.  When the exact H is zero
    // (k == 1, or all inlier trees coincide - the reference's centred sums are then exactly 0 and its SVD
    // returns R = I) the subtraction below leaves only rounding noise, of the order 1e-16 * sum |u||v|: detect
    // that (against the magnitude sum `habs`, not the signed sums, which may cancel too) and use H = 0.
    double h00 = __fma_rn(-s.su0, mv0, s.h00), h01 = __fma_rn(-s.su0, mv1, s.h01);
    double h10 = __fma_rn(-s.su1, mv0, s.h10), h11 = __fma_rn(-s.su1, mv1, s.h11);
    const double hsum = __dadd_rn(__dadd_rn(fabs(h00), fabs(h01)), __dadd_rn(fabs(h10), fabs(h11)));
    if (hsum <= __dmul_rn(1e-12, s.habs)) h00 = h01 = h10 = h11 = 0.0;
    double r00, r01, r10, r11;
    // reflection only when det(H) is negative beyond rounding noise (det == 0: SVD's choice is arbitrary)
    const double p1 = __dmul_rn(h00, h11), p2 = __dmul_rn(h01, h10);
    if (allow_reflection && __dsub_rn(p1, p2) < __dmul_rn(-1e-14, __dadd_rn(fabs(p1), fabs(p2)))) {
        const double a = __dsub_rn(h00, h11), b = __dadd_rn(h01, h10), nrm = __dsqrt_rn(__fma_rn(b, b, __dmul_rn(a, a)));
        const double c = (nrm == 0.0) ? 1.0 : __ddiv_rn(a, nrm), sn = (nrm == 0.0) ? 0.0 : __ddiv_rn(b, nrm);
        r00 = c; r01 = sn; r10 = sn; r11 = -c;
    } else {
        const double a = __dadd_rn(h00, h11), b = __dsub_rn(h01, h10), nrm = __dsqrt_rn(__fma_rn(b, b, __dmul_rn(a, a)));
        const double c = (nrm == 0.0) ? 1.0 : __ddiv_rn(a, nrm), sn = (nrm == 0.0) ? 0.0 : __ddiv_rn(b, nrm);
        r00 = c; r01 = -sn; r10 = sn; r11 = c;
    }
    const double ex = __dsub_rn(__dsub_rn(P.cx, ax), mu0), ey = __dsub_rn(__dsub_rn(P.cy, ay), mu1);
    Pose Q;
    Q.m00 = __fma_rn(r01, P.m10, __dmul_rn(r00, P.m00)); Q.m01 = __fma_rn(r01, P.m11, __dmul_rn(r00, P.m01));
    Q.m10 = __fma_rn(r11, P.m10, __dmul_rn(r10, P.m00)); Q.m11 = __fma_rn(r11, P.m11, __dmul_rn(r10, P.m01));
    Q.cx = __dadd_rn(__fma_rn(r01, ey, __dmul_rn(r00, ex)), __dadd_rn(ax, mv0));
    Q.cy = __dadd_rn(__fma_rn(r11, ey, __dmul_rn(r10, ex)), __dadd_rn(ay, mv1));
    D.m00 = __dsub_rn(Q.m00, P.m00); D.m01 = __dsub_rn(Q.m01, P.m01); D.m10 = __dsub_rn(Q.m10, P.m10); D.m11 = __dsub_rn(Q.m11, P.m11);
    D.cx = __dsub_rn(Q.cx, P.cx); D.cy = __dsub_rn(Q.cy, P.cy);
    P = Q;
}
// correspondence of tree i from its neighbour code (window-local or global position)
__device__ __forceinline__ double2 corr_xy(const GridView& G, const WindowAcc& W, int code) {
    return (code < 0) ? grid_xy(G, code & 0x7FFFFFFF) : W.xy[code & 0xFFFF];
}

}  // namespace
}  // namespace ficp
