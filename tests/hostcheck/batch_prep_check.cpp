// Host check of csrc/batch_prep.h under AddressSanitizer / UBSan / ThreadSanitizer (tests/test_host_cabi.py compiles and runs it):
// ragged plots, every thread count, the row-writing and the read-only form, slices of a batch addressed the way
// ficp_batch_create addresses them (offset pointers, absolute row indices) - compared with a plain serial restatement.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <random>
#include <vector>
#include "../../coregistrationgame_b200/csrc/batch_prep.h"

int main() {
    std::mt19937_64 rng(7);
    int checked = 0;
    for (int trial = 0; trial < 24; ++trial) {
        const int ld = 2 + (int)(rng() % 4);
        const bool z3 = ld >= 3 && (rng() & 1);
        const int64_t n_plots = 1 + (int64_t)(rng() % 700);
        std::vector<int64_t> off((size_t)n_plots + 1, 0);
        for (int64_t p = 0; p < n_plots; ++p) off[(size_t)p + 1] = off[(size_t)p] + 1 + (int64_t)(rng() % (trial % 3 == 0 ? 1024 : 200));
        const long long rows = off[(size_t)n_plots];
        std::vector<double> src((size_t)rows * ld);
        std::uniform_real_distribution<double> U(-300.0, 300.0);
        for (auto& v : src) v = 6.48e6 + U(rng);
        // serial restatement
        std::vector<double> cen(2 * (size_t)n_plots), u_ref(2 * (size_t)rows), ubar_ref(2 * (size_t)n_plots), far(n_plots);
        for (int64_t p = 0; p < n_plots; ++p) {
            const long long o = off[(size_t)p], n = off[(size_t)p + 1] - o;
            double sx = 0, sy = 0;
            for (long long i = 0; i < n; ++i) { sx += src[(size_t)(o + i) * ld]; sy += src[(size_t)(o + i) * ld + 1]; }
            cen[2 * p] = sx / (double)n; cen[2 * p + 1] = sy / (double)n;
            sx = sy = 0;
            for (long long i = 0; i < n; ++i) {
                const double ux = src[(size_t)(o + i) * ld] - cen[2 * p], uy = src[(size_t)(o + i) * ld + 1] - cen[2 * p + 1];
                u_ref[2 * (size_t)(o + i)] = ux; u_ref[2 * (size_t)(o + i) + 1] = uy; sx += ux; sy += uy;
            }
            ubar_ref[2 * p] = sx / (double)n; ubar_ref[2 * p + 1] = sy / (double)n;
            double m = 0;
            for (long long i = 0; i < n; ++i)
                m = std::max(m, std::hypot(u_ref[2 * (size_t)(o + i)] - ubar_ref[2 * p], u_ref[2 * (size_t)(o + i) + 1] - ubar_ref[2 * p + 1]));
            far[(size_t)p] = m;
        }
        for (int threads : {1, 2, 3, 8}) {
            std::vector<double> c2(2 * (size_t)n_plots, NAN), u(2 * (size_t)rows, NAN), z((size_t)rows, NAN), ubar(2 * (size_t)n_plots, NAN),
                rho((size_t)n_plots, NAN), ubar0(2 * (size_t)n_plots, NAN), rho0((size_t)n_plots, NAN);
            // in up to four slices of plots, like ficp_batch_create's staging route
            const int n_sl = 1 + (int)(rng() % 4);
            int64_t p0 = 0;
            for (int sl = 1; sl <= n_sl; ++sl) {
                const int64_t p1 = sl == n_sl ? n_plots : std::max<int64_t>(p0, n_plots * sl / n_sl);
                if (p1 == p0) continue;
                ficp::plot_centres_host(src.data(), ld, off.data() + p0, p1 - p0, c2.data() + 2 * p0, threads);
                if (!ficp::plot_geometry_host(src.data(), ld, z3, off.data() + p0, p1 - p0, c2.data() + 2 * p0, u.data(), z3 ? z.data() : nullptr,
                                              ubar.data() + 2 * p0, rho.data() + p0, threads)) { printf("FAIL: finite input refused\n"); return 1; }
                p0 = p1;
            }
            if (!ficp::plot_geometry_host(src.data(), ld, z3, off.data(), n_plots, c2.data(), nullptr, nullptr, ubar0.data(), rho0.data(), threads)) return 1;
            for (size_t i = 0; i < cen.size(); ++i) if (c2[i] != cen[i] || ubar[i] != ubar_ref[i] || ubar0[i] != ubar_ref[i]) { printf("FAIL: centre / ubar bits\n"); return 1; }
            for (size_t i = 0; i < u.size(); ++i) if (u[i] != u_ref[i]) { printf("FAIL: u bits\n"); return 1; }
            if (z3) for (long long i = 0; i < rows; ++i) if (z[(size_t)i] != src[(size_t)i * ld + 2]) { printf("FAIL: z\n"); return 1; }
            for (int64_t p = 0; p < n_plots; ++p)
                if (!(rho[(size_t)p] >= far[(size_t)p]) || rho[(size_t)p] > far[(size_t)p] * (1 + 1e-14) + 2.1e-6 * (far[(size_t)p] + 1) || rho0[(size_t)p] != rho[(size_t)p]) {
                    printf("FAIL: radius of plot %lld: %.17g vs %.17g\n", (long long)p, rho[(size_t)p], far[(size_t)p]); return 1;
                }
            // one bad coordinate anywhere is found by every thread count
            const size_t bad = (size_t)(rng() % (uint64_t)rows) * ld + (size_t)(rng() % (z3 ? 3 : 2));
            const double keep = src[bad];
            src[bad] = (rng() & 1) ? INFINITY : NAN;
            if (ficp::plot_geometry_host(src.data(), ld, z3, off.data(), n_plots, c2.data(), nullptr, nullptr, ubar0.data(), rho0.data(), threads)) { printf("FAIL: non-finite row accepted\n"); return 1; }
            src[bad] = keep;
            ++checked;
        }
    }
    printf("batch_prep host check: %d cases ok\n", checked);
    return 0;
}
