// CPU unit test of the accessor-generic grid NN search (coregistrationgame_b200/csrc/nn_search.cuh).
// Builds the same cell-sorted layout the CUDA grid-build kernels produce, then checks
// nn_search<GlobalAcc> and nn_search<WindowAcc> against an O(N*M) brute force with the
// lowest-original-index tie rule.  Test infrastructure only.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>
#include "nn_search.cuh"

using namespace ficp;

struct HostGrid {
    GridGeom g;
    std::vector<double2> xy;
    std::vector<double> z;
    std::vector<int> org;
    std::vector<unsigned> cs;
};

// `shrink` > 0: the grid spans only the central part of the bounding box (robust extent of grid_build.cu for skewed
// targets); the points outside are clamped into the border cells.  `shuffle`: arbitrary order inside every cell (what the
// atomics of the scatter kernel may leave in cells too heavy to re-order) - results must not depend on it.
static HostGrid build(const std::vector<double>& px, const std::vector<double>& py, const std::vector<double>& pz,
                      double ppc, double shrink = 0.0, bool shuffle = false) {
    HostGrid G;
    const size_t m = px.size();
    double x0 = 1e300, x1 = -1e300, y0 = 1e300, y1 = -1e300;
    for (size_t i = 0; i < m; ++i) {
        x0 = std::min(x0, px[i]); x1 = std::max(x1, px[i]);
        y0 = std::min(y0, py[i]); y1 = std::max(y1, py[i]);
    }
    G.g.tx0 = x0; G.g.tx1 = x1; G.g.ty0 = y0; G.g.ty1 = y1; G.g.clamped = 0; G.g.pad = 0;
    if (shrink > 0.0) {
        const double wx = x1 - x0, wy = y1 - y0;
        x0 += shrink * wx; x1 -= shrink * wx; y0 += 0.5 * shrink * wy; y1 -= 1.5 * shrink * wy;
        G.g.clamped = 1;
    }
    double ex = x1 - x0, ey = y1 - y0;
    double big = std::max(ex, ey);
    double h;
    if (big <= 0) h = 1.0;
    else {
        double exx = std::max(ex, big * 1e-6), eyy = std::max(ey, big * 1e-6);
        h = std::sqrt(ppc * exx * eyy / (double)m);
    }
    G.g.x0 = x0; G.g.y0 = y0; G.g.h = h; G.g.inv_h = 1.0 / h;
    G.g.gw = (int)std::floor(ex / h) + 1; G.g.gh = (int)std::floor(ey / h) + 1;
    G.g.eps = h * 1e-9 + (std::fabs(x0) + std::fabs(y0) + big) * 8e-16;
    const int nc = G.g.gw * G.g.gh;
    std::vector<int> cell(m);
    G.cs.assign(nc + 1, 0);
    for (size_t i = 0; i < m; ++i) {
        int cx = clamp_cell((px[i] - x0) * G.g.inv_h, G.g.gw), cy = clamp_cell((py[i] - y0) * G.g.inv_h, G.g.gh);
        cell[i] = cy * G.g.gw + cx;
        G.cs[cell[i] + 1]++;
    }
    for (int c = 0; c < nc; ++c) G.cs[c + 1] += G.cs[c];
    std::vector<unsigned> fill(G.cs.begin(), G.cs.end() - 1);
    G.xy.resize(m); G.z.resize(m); G.org.resize(m);
    for (size_t i = 0; i < m; ++i) {  // stable: ascending original index within a cell
        unsigned p = fill[cell[i]]++;
        G.xy[p] = make_double2(px[i], py[i]); G.z[p] = pz[i]; G.org[p] = (int)i;
    }
    if (shuffle) {
        std::mt19937 r2(777);
        for (int c = 0; c < nc; ++c)
            for (unsigned a = G.cs[c]; a + 1 < G.cs[c + 1]; ++a) {
                const unsigned b = a + r2() % (G.cs[c + 1] - a);
                std::swap(G.xy[a], G.xy[b]); std::swap(G.z[a], G.z[b]); std::swap(G.org[a], G.org[b]);
            }
    }
    return G;
}

template <bool Z3>
static int check(const HostGrid& G, const std::vector<double>& px, const std::vector<double>& py,
                 const std::vector<double>& pz, const std::vector<double>& qx, const std::vector<double>& qy,
                 const std::vector<double>& qz, int wx0, int wy0, int wx1, int wy1, long* n_window_hits,
                 long* n_lb_checked, long* n_lb_tight, long* n_lb_runner) {
    // XYZ instantiation reads the packed 32 B records, XY the split arrays (exactly like the device layouts)
    std::vector<double4> rec(G.xy.size());
    for (size_t i = 0; i < rec.size(); ++i) rec[i] = make_double4(G.xy[i].x, G.xy[i].y, G.z[i], index_to_bits(G.org[i]));
    GlobalAcc ga{Z3 ? nullptr : G.xy.data(), Z3 ? rec.data() : nullptr, Z3 ? nullptr : G.org.data(), G.cs.data(), G.g.gw};
    // window staging exactly like the kernel does it
    const int ww = wx1 - wx0, wh = wy1 - wy0;
    std::vector<double2> wxy; std::vector<double> wz; std::vector<unsigned> wc(std::max(ww * wh, 1));
    std::vector<int> rowoff(wh + 1, 0), rowdelta(std::max(wh, 1), 0);
    for (int r = 0; r < wh; ++r) {
        unsigned gs = G.cs[(size_t)(wy0 + r) * G.g.gw + wx0], ge = G.cs[(size_t)(wy0 + r) * G.g.gw + wx1];
        rowoff[r + 1] = rowoff[r] + (int)(ge - gs);
        rowdelta[r] = (int)gs - rowoff[r];
        for (unsigned j = gs; j < ge; ++j) { wxy.push_back(G.xy[j]); wz.push_back(G.z[j]); }
        for (int c = 0; c < ww; ++c) {
            unsigned a = G.cs[(size_t)(wy0 + r) * G.g.gw + wx0 + c], b = G.cs[(size_t)(wy0 + r) * G.g.gw + wx0 + c + 1];
            wc[r * ww + c] = (unsigned)(rowoff[r] + (int)(a - gs)) | ((b - a) << 16);
        }
    }
    WindowAcc wa{wxy.data(), wz.data(), wc.data(), rowoff.data(), rowdelta.data(), Z3 ? nullptr : G.org.data(), Z3 ? rec.data() : nullptr,
                 wx0, wy0, wx1, wy1, ww, wh};
    int bad = 0;
    for (size_t i = 0; i < qx.size(); ++i) {
        // brute force
        double bb = kInf; int bi = -1;
        for (size_t j = 0; j < px.size(); ++j) {
            double dx = qx[i] - px[j], dy = qy[i] - py[j];
            double d2 = dx * dx + dy * dy;
            if (Z3) { double dz = qz[i] - pz[j]; d2 = d2 + dz * dz; }
            if (d2 < bb) { bb = d2; bi = (int)j; }
        }
        // second-nearest (any point other than the winner itself; a duplicate of the winner counts)
        double b2nd = kInf;
        for (size_t j = 0; j < px.size(); ++j) {
            if ((int)j == bi) continue;
            double dx = qx[i] - px[j], dy = qy[i] - py[j];
            double d2 = dx * dx + dy * dy;
            if (Z3) { double dz = qz[i] - pz[j]; d2 = d2 + dz * dz; }
            if (d2 < b2nd) b2nd = d2;
        }
        double best; int pos;
        nn_search<Z3>(ga, G.g, qx[i], qy[i], qz[i], best, pos);
        if (G.org[pos] != bi || best != bb) {
            if (bad < 5) printf("GLOBAL mismatch q%zu: got %d (%.17g) want %d (%.17g)\n", i, G.org[pos], best, bi, bb);
            ++bad;
        }
        // streamed 3x3 form, seeded with: nothing, the true NN, an arbitrary point, a near point
        for (int variant = 0; variant < 4; ++variant) {
            int prev = -1;
            if (variant == 1) prev = pos;
            if (variant == 2) prev = (int)((i * 7919u + 13u) % px.size());
            if (variant == 3) prev = std::min<int>((int)px.size() - 1, pos + 1);
            double b3; int p3;
            nn_search_stream<Z3>(ga, G.g, qx[i], qy[i], qz[i], prev, b3, p3);
            if (G.org[p3] != bi || b3 != bb) {
                if (bad < 5) printf("STREAM(global,v%d) mismatch q%zu: got %d (%.17g) want %d (%.17g)\n", variant, i, G.org[p3], b3, bi, bb);
                ++bad;
            }
        }
        // bulk-query form (nn_bulk.cu): own cell first, pruned 3x3 stream with a tie flag, then rings >= 2 if unsettled
        {
            const int cxq = clamp_cell((qx[i] - G.g.x0) * G.g.inv_h, G.g.gw), cyq = clamp_cell((qy[i] - G.g.y0) * G.g.inv_h, G.g.gh);
            double b6; int p6;
            nn_search_block3_unseeded<Z3>(ga, G.g, qx[i], qy[i], qz[i], cxq, cyq, b6, p6);
            if (!nn_block_settles(G.g, qx[i], qy[i], cxq, cyq, 1, b6)) nn_ring_loop_impl<Z3>(ga, G.g, qx[i], qy[i], qz[i], cxq, cyq, 2, b6, p6);
            if (p6 < 0 || G.org[p6] != bi || b6 != bb) {
                if (bad < 5) printf("BULK(unseeded) mismatch q%zu: got %d (%.17g) want %d (%.17g)\n", i, p6 < 0 ? -1 : G.org[p6], b6, bi, bb);
                ++bad;
            }
        }
        // lower bound on every OTHER point from the tracked 3x3 search (what the ICP kernel's skip test relies on)
        for (int variant = 0; variant < 4; ++variant) {
            int prev = -1;
            if (variant == 1) prev = pos;
            if (variant == 2) prev = (int)((i * 7919u + 13u) % px.size());
            if (variant == 3) prev = std::min<int>((int)px.size() - 1, pos + 1);
            double b3; int p3, cx, cy, lbh, p2nd;
            if (!nn_search_block3_impl<Z3, true>(ga, G.g, qx[i], qy[i], qz[i], prev, b3, p3, cx, cy, lbh, p2nd)) continue;
            const double border2 = nn_block_border2(G.g, qx[i], qy[i], cx, cy, 1);
            if (nn_block_settles(G.g, qx[i], qy[i], cx, cy, 1, b3) != (b3 < border2)) {
                if (bad < 5) printf("BORDER2 disagrees with nn_block_settles q%zu\n", i);
                ++bad;
            }
            if (!(b3 < border2)) continue;
            ++*n_lb_checked;
            const double lb2 = std::min(hi_to_double(lbh), border2);
            // the bound covers every point other than the winner and the runner-up
            double b_others = kInf;
            const int o2 = (p2nd >= 0) ? G.org[p2nd] : -1;
            if (p2nd == p3) { if (bad < 5) printf("runner-up equals winner q%zu\n", i); ++bad; }
            for (size_t j = 0; j < px.size(); ++j) {
                if ((int)j == bi || (int)j == o2) continue;
                double dx = qx[i] - px[j], dy = qy[i] - py[j];
                double d2 = dx * dx + dy * dy;
                if (Z3) { double dz = qz[i] - pz[j]; d2 = d2 + dz * dz; }
                if (d2 < b_others) b_others = d2;
            }
            if (o2 >= 0 && b2nd < b_others) ++*n_lb_runner;   // the runner-up is the true second-nearest
            b2nd = b_others;
            if (G.org[p3] != bi || b3 != bb || lb2 > b2nd) {
                if (bad < 5) printf("LOWER-BOUND(v%d) q%zu: winner %d (%.17g) want %d (%.17g); lb2 %.17g > second %.17g\n", variant, i, G.org[p3], b3, bi, bb, lb2, b2nd);
                ++bad;
            }
            if (lb2 >= b2nd * 0.999) ++*n_lb_tight;
        }
        if (ww > 0 && wh > 0) {
            double b2; int p2;
            if (nn_search<Z3>(wa, G.g, qx[i], qy[i], qz[i], b2, p2)) {
                ++*n_window_hits;
                int gp = wa.global_pos(p2);
                if (G.org[gp] != bi || b2 != bb) {
                    if (bad < 5) printf("WINDOW mismatch q%zu: got %d (%.17g) want %d (%.17g)\n", i, G.org[gp], b2, bi, bb);
                    ++bad;
                }
            }
            for (int variant = 0; variant < 3; ++variant) {
                int prev = -1;
                const int wtot = rowoff[wh];
                if (variant == 1 && wtot > 0) prev = (int)((i * 104729u + 7u) % (unsigned)wtot);
                if (variant == 2 && wtot > 0) prev = wtot - 1;
                // tracked form over the WINDOW accessor - the instantiation the ICP kernel uses
                {
                    double b5; int p5, cx5, cy5, lbh5, r5;
                    if (nn_search_block3_impl<Z3, true>(wa, G.g, qx[i], qy[i], qz[i], prev, b5, p5, cx5, cy5, lbh5, r5)) {
                        const double border2 = nn_block_border2(G.g, qx[i], qy[i], cx5, cy5, 1);
                        if (b5 < border2) {
                            ++*n_lb_checked;
                            const int won = G.org[wa.global_pos(p5)];
                            const int o2 = (r5 >= 0) ? G.org[wa.global_pos(r5)] : -1;
                            double b_others = kInf;
                            for (size_t j = 0; j < px.size(); ++j) {
                                if ((int)j == bi || (int)j == o2) continue;
                                double dx = qx[i] - px[j], dy = qy[i] - py[j];
                                double d2 = dx * dx + dy * dy;
                                if (Z3) { double dz = qz[i] - pz[j]; d2 = d2 + dz * dz; }
                                if (d2 < b_others) b_others = d2;
                            }
                            const double lb2 = std::min(hi_to_double(lbh5), border2);
                            if (won != bi || b5 != bb || r5 == p5 || lb2 > b_others) {
                                if (bad < 5) printf("LOWER-BOUND(window,v%d) q%zu: winner %d (%.17g) want %d (%.17g); lb2 %.17g vs others %.17g\n", variant, i, won, b5, bi, bb, lb2, b_others);
                                ++bad;
                            }
                        }
                    }
                }
                double b4; int p4;
                if (nn_search_stream<Z3>(wa, G.g, qx[i], qy[i], qz[i], prev, b4, p4)) {
                    ++*n_window_hits;
                    int gp = wa.global_pos(p4);
                    if (G.org[gp] != bi || b4 != bb) {
                        if (bad < 5) printf("STREAM(window,v%d) mismatch q%zu: got %d (%.17g) want %d (%.17g)\n", variant, i, G.org[gp], b4, bi, bb);
                        ++bad;
                    }
                }
            }
        }
    }
    return bad;
}

int main() {
    std::mt19937_64 rng(12345);
    std::uniform_real_distribution<double> U(0.0, 1.0);
    int bad = 0; long hits = 0; long total = 0; long lbc = 0, lbt = 0, lbr = 0;
    for (int trial = 0; trial < 24; ++trial) {
        const size_t m = (trial % 4 == 0) ? 37 : 3000 + 500 * trial;
        const double side = 300.0 * (1 + trial % 3);
        const double offx = (trial % 2) ? 420000.0 : 0.0, offy = (trial % 2) ? 6483000.0 : -50.0;
        std::vector<double> px(m), py(m), pz(m);
        for (size_t i = 0; i < m; ++i) { px[i] = offx + side * U(rng); py[i] = offy + side * (trial % 5 == 1 ? 0.02 : 1.0) * U(rng); pz[i] = 5 + 30 * U(rng); }
        if (trial % 3 == 1 && m > 200) {  // lattice patch + duplicates: exact ties
            int k = 0;
            for (int a = 0; a < 8; ++a) for (int b = 0; b < 8; ++b, ++k) { px[k] = offx + side / 2 + a; py[k] = offy + side * (trial % 5 == 1 ? 0.01 : 0.5) + b * 0.25; pz[k] = 20; }
            for (size_t i = m - m / 10; i < m; ++i) { px[i] = px[i - m / 2]; py[i] = py[i - m / 2]; pz[i] = pz[i - m / 2]; }
        }
        if (trial == 7) for (size_t i = 0; i < m; ++i) py[i] = offy;           // collinear
        if (trial == 11) for (size_t i = 0; i < m; ++i) { px[i] = offx; py[i] = offy; }  // all identical
        // every third trial: far outliers in the target (placeholder rows at the origin, a stray point)
        if (trial % 3 == 2 && m > 100) { px[1] = 0.0; py[1] = 0.0; px[2] = offx - 40 * side; py[2] = offy + 3 * side; px[3] = offx + 9 * side; py[3] = offy - 7 * side; }
        // grids: spanning the bounding box (even trials) / robust core extent with clamped border cells (odd trials >= 5);
        // arbitrary in-cell order on trials 2 mod 4
        HostGrid G = build(px, py, pz, trial % 2 ? 2.0 : 1.0, (trial >= 5 && trial % 2 == 1) ? 0.07 + 0.01 * (trial % 5) : (trial % 3 == 2 && m > 100 ? 0.02 : 0.0), trial % 4 == 2);
        const size_t nq = 400;
        std::vector<double> qx(nq), qy(nq), qz(nq);
        for (size_t i = 0; i < nq; ++i) {
            double r = U(rng);
            if (r < 0.6) { qx[i] = offx + side * U(rng); qy[i] = offy + side * U(rng); }
            else if (r < 0.8) { size_t j = rng() % m; qx[i] = px[j] + 0.3 * (U(rng) - 0.5); qy[i] = py[j] + 0.3 * (U(rng) - 0.5); }
            else if (r < 0.86) { qx[i] = offx + side * (3 * U(rng) - 1); qy[i] = offy + side * (3 * U(rng) - 1); }  // outside the grid
            else if (r < 0.88) { size_t j = 1 + rng() % 3; qx[i] = px[j] + 5 * (U(rng) - 0.5); qy[i] = py[j] + 5 * (U(rng) - 0.5); }  // next to a (possibly far-off) target point
            else if (r < 0.9) { qx[i] = offx + side * (40 * U(rng) - 20); qy[i] = offy + side * (40 * U(rng) - 20); }  // far off the map
            else { qx[i] = offx + side / 2 + (rng() % 8) + 0.5; qy[i] = offy + side * 0.5 + (rng() % 8) * 0.25 + 0.125; }  // lattice centres
            qz[i] = 5 + 30 * U(rng);
        }
        // window: central part of the grid
        int wx0 = G.g.gw / 4, wx1 = std::max(wx0 + 1, 3 * G.g.gw / 4), wy0 = G.g.gh / 4, wy1 = std::max(wy0 + 1, 3 * G.g.gh / 4);
        wx1 = std::min(wx1, G.g.gw); wy1 = std::min(wy1, G.g.gh);
        size_t wpts = 0;
        for (int r = wy0; r < wy1; ++r) wpts += G.cs[(size_t)r * G.g.gw + wx1] - G.cs[(size_t)r * G.g.gw + wx0];
        if (wpts > 65535) { wx1 = wx0; }
        bad += check<false>(G, px, py, pz, qx, qy, qz, wx0, wy0, wx1, wy1, &hits, &lbc, &lbt, &lbr);
        bad += check<true>(G, px, py, pz, qx, qy, qz, wx0, wy0, wx1, wy1, &hits, &lbc, &lbt, &lbr);
        total += 2 * nq;
    }
    printf("queries=%ld window_resolved=%ld lower_bounds_checked=%ld (tight: %ld, runner-up is the 2nd nearest: %ld) mismatches=%d\n", total, hits, lbc, lbt, lbr, bad);
    return bad ? 1 : 0;
}
