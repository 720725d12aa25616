// Host-side geometry of a batch of plots, shared by ficp_batch_create, ficp_plot_centres and ficp_plot_geometry
// (capi.cu).  Pure C++ (no CUDA types): tests/test_host_cabi.py drives it through the C ABI on a machine without a GPU.
//
// Per plot p with rows r_0 .. r_{n-1} (row-major, `ld` doubles per row):
//   centre  = (r_0 + r_1 + ... + r_{n-1}) / n, added in row order - the value of `rows[:, :2].mean(axis=0)`, which is what
//             the reference-side callers and the oracle take as the point start poses rotate about (trees.py:201-222)
//   u_i     = r_i - centre          one subtraction per coordinate, the oracle's `pre_transform` arithmetic
//   ubar    = (u_0 + ... + u_{n-1}) / n in row order (shift point of the fit sums)
//   rho     >= max_i |u_i - ubar|   radius of the plot about ubar (at most 1e-6 (rho + 1) above it); only sizes the shared-memory
//                                   window, never a result
// Config 4 (1250 plots x 150 trees per GPU, one ICP each) spends 0.22 ms on the device per batch, so this pass is what an
// end-to-end step costs: rho from the largest SQUARED distance (one square root per plot instead of one hypot per tree:
// 6.6 -> 1.0 ms for 187 500 rows in the build container); plots are independent, so large batches use a few host threads.
#pragma once
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <thread>
#include <vector>

namespace ficp {

// Threads for a host pass over `rows` source rows.  Measured on the B200 box (profiles/r02_c4_e2e_probe.jsonl: 187 500 rows,
// ficp_batch_create as a whole): 1 thread 0.70 ms, 2 threads 0.68, 4 threads 0.90, 8 threads 0.97 - starting and joining a
// thread costs as much as 50 K rows of the pass, so a thread is added per 512 K rows only, at most 4 (ranks of a multi-GPU
// job share the host).  FICP_HOST_THREADS=n asks for up to n threads, one per 32 K rows (1 = always serial).
inline int host_threads_for(long long rows) {
    int cap = 4;
    long long rows_per_thread = 524288;
    if (const char* e = std::getenv("FICP_HOST_THREADS")) {
        const int v = std::atoi(e);
        if (v >= 1) { cap = std::min(v, 64); rows_per_thread = 32768; }
    }
    const unsigned hw = std::thread::hardware_concurrency();
    if (hw >= 1) cap = std::min<int>(cap, (int)hw);
    return (int)std::max<long long>(1, std::min<long long>(cap, rows / rows_per_thread));
}

// run fn(p_begin, p_end) over [0, n_plots) cut into contiguous ranges of roughly equal ROW count
template <class Fn>
inline void for_plot_ranges(const int64_t* offsets, int64_t n_plots, int threads, Fn fn) {
    if (threads <= 1 || n_plots < 2 * threads) { fn((int64_t)0, n_plots); return; }
    std::vector<int64_t> cut((size_t)threads + 1, n_plots);
    cut[0] = 0;
    const long long rows = offsets[n_plots] - offsets[0];
    for (int t = 1; t < threads; ++t) {
        const long long want = offsets[0] + rows * t / threads;
        cut[(size_t)t] = std::lower_bound(offsets, offsets + n_plots, want) - offsets;
    }
    std::vector<std::thread> pool;
    pool.reserve((size_t)threads - 1);
    for (int t = 1; t < threads; ++t)
        pool.emplace_back([&, t] { if (cut[(size_t)t] < cut[(size_t)t + 1]) fn(cut[(size_t)t], cut[(size_t)t + 1]); });
    if (cut[0] < cut[1]) fn(cut[0], cut[1]);
    for (auto& th : pool) th.join();
}

// centres[2p..2p+1] = mean of the first two columns of plot p, rows added in order
inline void plot_centres_host(const double* src, int ld, const int64_t* offsets, int64_t n_plots, double* centres, int threads) {
    for_plot_ranges(offsets, n_plots, threads, [&](int64_t p0, int64_t p1) {
        for (int64_t p = p0; p < p1; ++p) {
            const long long off = offsets[p], n = offsets[p + 1] - offsets[p];
            double sx = 0.0, sy = 0.0;
            for (long long i = 0; i < n; ++i) {
                const double* r = src + (size_t)(off + i) * ld;
                sx += r[0];
                sy += r[1];
            }
            centres[2 * p] = sx / (double)n;
            centres[2 * p + 1] = sy / (double)n;
        }
    });
}

// u (2 doubles per row), z (1 per row, when z3), ubar (2 per plot), rho (1 per plot).  Returns false when a matched
// coordinate is not finite (the reference raises in scipy: "'x' must be finite", ficp.py:70).  u == nullptr: the per-plot
// values only, nothing is written per row (ficp_batch_create when the device splits the rows itself) - a read-only pass.
inline bool plot_geometry_host(const double* src, int ld, bool z3, const int64_t* offsets, int64_t n_plots, const double* centres,
                               double* u, double* z, double* ubar, double* rho, int threads) {
    std::atomic<bool> finite{true};
    constexpr double kUp = 1.0 + 8.0 * 2.220446049250313e-16;   // covers the roundings of a*a + b*b and of the square root
    for_plot_ranges(offsets, n_plots, threads, [&](int64_t p0, int64_t p1) {
        bool ok = true;
        for (int64_t p = p0; p < p1; ++p) {
            const long long off = offsets[p], n = offsets[p + 1] - offsets[p];
            const double cx = centres[2 * p], cy = centres[2 * p + 1];
            // One walk over the rows, four independent chains: the two in-order sums, the largest squared distance from the
            // centre, and a finiteness witness for z.  A non-finite x or y needs no test of its own: it makes u = x - c
            // non-finite, and a sum that has met an inf or a NaN never becomes finite again.
            double sx = 0.0, sy = 0.0, zw = 0.0, m2c = 0.0;
            for (long long i = 0; i < n; ++i) {
                const double* r = src + (size_t)(off + i) * ld;
                const double ux = r[0] - cx, uy = r[1] - cy;   // same single subtraction as the oracle
                if (u) {
                    u[2 * (size_t)(off + i)] = ux;
                    u[2 * (size_t)(off + i) + 1] = uy;
                    if (z3) z[(size_t)(off + i)] = r[2];
                }
                if (z3) zw += r[2] - r[2];                     // 0 for a finite z, NaN for inf / NaN
                sx += ux;
                sy += uy;
                m2c = std::max(m2c, ux * ux + uy * uy);
            }
            if (!(((sx - sx) + (sy - sy)) + zw == 0.0)) {
                // suspicious (a non-finite coordinate - or finite ones whose sum overflowed): the rows decide
                for (long long i = 0; i < n; ++i) {
                    const double* r = src + (size_t)(off + i) * ld;
                    const double probe = (r[0] - r[0]) + (r[1] - r[1]) + (z3 ? (r[2] - r[2]) : 0.0);
                    ok &= (probe == 0.0);
                }
            }
            const double bx = sx / (double)n, by = sy / (double)n;
            ubar[2 * p] = bx;
            ubar[2 * p + 1] = by;
            // radius about ubar, rounded up.  When the plot turns about its own centroid ubar is rounding noise and
            // max |u| + |ubar| bounds it (triangle inequality) without a second walk; otherwise (a caller's centre far from
            // the plot, e.g. the origin for FractionalICP.run) the distances from ubar are taken one by one.
            const double rc = std::sqrt(m2c) * kUp, nb = std::sqrt(bx * bx + by * by) * kUp;
            if (nb <= 1e-6 * (rc + 1.0)) {
                rho[p] = (rc + nb) * kUp;
            } else {
                double m2 = 0.0;
                for (long long i = 0; i < n; ++i) {
                    const double* r = src + (size_t)(off + i) * ld;   // the plot's rows are in L1 from the walk above
                    const double a = (r[0] - cx) - bx, b = (r[1] - cy) - by;
                    m2 = std::max(m2, a * a + b * b);
                }
                rho[p] = std::sqrt(m2) * kUp;
            }
        }
        if (!ok) finite.store(false, std::memory_order_relaxed);
    });
    return finite.load();
}

}  // namespace ficp
