"""CPU model of the ICP kernel's skip test (coregistrationgame_b200/csrc/icp_persistent.cu::nn_test_round and the bound
kept by nn_search_block3_impl<TRACK>): a searched query remembers its nearest neighbour, the runner-up and a lower bound
L on the distance to every OTHER target point; on later passes L shrinks by the distance the pose update moved the tree,
and while the nearer of the two remembered points is strictly inside the remaining bound it IS the nearest neighbour.

This restates the rule in numpy - same sources of the bound (third-best of the streamed cells, pruned cells of the 3x3
block, block border), same directed roundings (bound stored as a round-down half, moves rounded up) - and checks on real
ICP trajectories of the oracle that no skipped query ever disagrees with an exact search.  Test infrastructure only;
the CUDA code itself is checked bit for bit in tests/test_gpu_icp.py (window path vs global-grid path) and, for the
bound, in tests/hostcheck/nn_search_check.cpp."""
import math

import numpy as np
import pytest
from scipy.spatial import cKDTree

from oracle import ficp_oracle as orc

PRUNE = math.sqrt(1.5)   # cells up to 1.5 x the seed's SQUARED distance are streamed (FICP_PRUNE_PAD)


def half_round_down(x):
    x = np.maximum(x, 0.0)
    h = x.astype(np.float16)
    h = np.where(h.astype(np.float64) > x, np.nextafter(h, np.float16(-np.inf)), h)
    return np.maximum(h.astype(np.float64), 0.0)


def trajectory(src, tgt, tree, md):
    """Positions of every pass of the two-stage loop (ficp.py:122-154), from the oracle's building blocks."""
    out = []
    for lam in (3.0, orc.STAGE2_LAMBDA[md]):
        first, cur = True, 0.0
        while True:
            q = src[:, :md].copy()
            d, idx = tree.query(q, k=3)
            out.append((q, idx, d))
            k, value, order = orc.select_fraction_cumsum(d[:, 0] ** 2, lam)
            if first:
                cur, first = value, False
            else:
                if cur - value <= 1e-6:
                    break
                cur = value
            inl = order[:k]
            src = orc.apply_xy(src, orc.fit_rigid2d_closed(src[inl, :2], tgt[idx[inl, 0], :2], False))
    return out


@pytest.mark.parametrize("dims,ppc", [(3, 6.0), (2, 3.0), (3, 1.0)])
def test_skipped_queries_keep_their_exact_neighbour(dims, ppc):
    n, m = 120, 30000
    tgt, plots, _ = orc.synthetic_scene(m, n, seed=5, dims=dims, hidden_pose=True, out_frac=0.1)
    src0 = plots[0]
    centre = src0[:, :2].mean(axis=0)
    hyp = orc.hypothesis_table(8, flips=(0, 1), translations=orc.translation_lattice(2, 2.5))
    tree = cKDTree(tgt[:, :dims])
    h = math.sqrt(ppc / 0.05)
    x0, y0 = tgt[:, 0].min(), tgt[:, 1].min()
    total = searched = 0
    for hi in range(0, hyp.shape[0], 5):
        slack = np.zeros(n)
        p1 = p2 = prev_q = None
        for q, idx, d in trajectory(orc.pre_transform(src0, hyp[hi], centre), tgt, tree, dims):
            nn = idx[:, 0]
            if prev_q is None:
                need = np.ones(n, bool)
                dseed = np.full(n, np.inf)
            else:
                move = np.sqrt(((q[:, :2] - prev_q[:, :2]) ** 2).sum(1)) * (1 + 1e-9) + 1e-9     # rounded up
                slack = half_round_down((slack - move) * (1 - 2.0 ** -20))
                d1 = np.sqrt(((q - tgt[p1, :dims]) ** 2).sum(1))
                d2 = np.sqrt(((q - tgt[p2, :dims]) ** 2).sum(1))
                dseed = d1
                ok = (np.minimum(d1, d2) < slack) & ((p1 == p2) | (d1 != d2))
                winner = np.where(d2 < d1, p2, p1)
                # THE property: a query that skips its search has exactly the neighbour an exact search finds
                assert np.array_equal(winner[ok], nn[ok])
                swap = ok & (d2 < d1)
                p1, p2 = np.where(swap, p2, p1), np.where(swap, p1, p2)
                need = ~ok
            # bound kept by a search of the 3x3 block around the query's cell
            cx, cy = np.floor((q[:, 0] - x0) / h), np.floor((q[:, 1] - y0) / h)
            ux, uy = q[:, 0] - (x0 + cx * h), q[:, 1] - (y0 + cy * h)
            gx = np.stack([ux, np.zeros(n), h - ux], 1)
            gy = np.stack([uy, np.zeros(n), h - uy], 1)
            gap = np.sqrt(gx[:, :, None] ** 2 + gy[:, None, :] ** 2)
            pruned = gap > dseed[:, None, None] * PRUNE
            min_pruned = np.where(pruned, gap, np.inf).reshape(n, 9).min(1)
            border = np.minimum(np.minimum(ux, h - ux), np.minimum(uy, h - uy)) + h
            settled = d[:, 0] < border
            bound = np.where(settled, np.minimum(np.minimum(d[:, 2], min_pruned), border), 0.0)   # deferred: no bound
            slack = np.where(need, half_round_down(np.minimum(bound, 60000.0)), slack)
            p1 = nn.copy() if p1 is None else np.where(need, nn, p1)
            p2 = idx[:, 1].copy() if p2 is None else np.where(need, idx[:, 1], p2)
            total += n
            searched += int(need.sum())
            prev_q = q
    assert searched < total            # the model does skip (otherwise the test checks nothing)
    if ppc >= 3.0:
        assert searched < 0.5 * total
