"""CPU model of the ICP kernel's search phase on oracle trajectories (analysis aid, not a test; test infrastructure only):
which queries fail the skip test, how many candidates each of their searches streams, and what a warp pays for a round of
32 of them (the slowest lane) under different orders of the search list.  Reproduces the counters measured on the B200
(profiles/r01_summary.md): ~7 % searched queries, 14 + 28 candidate-pair iterations per pass at 6 points per cell.

    python tests/sim_search_balance.py

Finding (round 1): ordering the list does not help - there are only ~1.2 seeded rounds per pass, so a round's cost is set
by its single longest query (max ~49 candidates vs 15 on average); only splitting the candidates of a round evenly over
the lanes (flattened query x candidate list, segmented min) would approach the ideal, 3x fewer iterations."""
import math
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scipy.spatial import cKDTree
from oracle import ficp_oracle as orc
from coregistrationgame_b200.synthetic import synthetic_scene
M = 200000; N = 500; PPC = 6.0
tgt, plots, poses = synthetic_scene(M, N, seed=3, dims=3, hidden_pose=True)
src0 = plots[0]; centre = src0[:, :2].mean(axis=0)
hyp = orc.hypothesis_table(128, flips=(0, 1), translations=orc.translation_lattice(4, 2.5))
tree = cKDTree(tgt)
h = math.sqrt(PPC / 0.05); x0 = tgt[:,0].min(); y0 = tgt[:,1].min()
gw = int((tgt[:,0].max()-x0)//h)+1; gh=int((tgt[:,1].max()-y0)//h)+1
cellcnt = np.zeros((gh+2, gw+2), int)
cxp = np.minimum(((tgt[:,0]-x0)/h).astype(int), gw-1); cyp = np.minimum(((tgt[:,1]-y0)/h).astype(int), gh-1)
np.add.at(cellcnt, (cyp+1, cxp+1), 1)
def half_rd(x):
    x = np.maximum(x, 0); hv = x.astype(np.float16)
    hv = np.where(hv.astype(np.float64) > x, np.nextafter(hv, np.float16(-np.inf)), hv)
    return np.maximum(hv.astype(np.float64), 0)
PR = math.sqrt(1.5)
tot = dict(identity=0, prev_d2=0, by_count=0, ideal=0, rounds=0, cands=0)
tot0 = dict(maxsum=0, cands=0, rounds=0)
npass = 0
for hi in range(0, 4096, 4096 // 24):
    src = orc.pre_transform(src0, hyp[hi], centre)
    slack = np.zeros(N); p1 = p2 = prev_q = None; prev_d = None
    for lam in (3.0, 0.95):
        first = True; cur = 0
        while True:
            q = src[:, :3].copy()
            d, idx = tree.query(q, k=3); nn = idx[:, 0]
            if prev_q is None:
                need = np.ones(N, bool); dseed = np.full(N, np.inf)
            else:
                move = np.sqrt(((q[:, :2] - prev_q[:, :2]) ** 2).sum(1)) * (1 + 1e-9) + 1e-9
                slack = half_rd((slack - move) * (1 - 2.0 ** -20))
                d1 = np.sqrt(((q - tgt[p1]) ** 2).sum(1)); d2 = np.sqrt(((q - tgt[p2]) ** 2).sum(1))
                dseed = d1
                ok = (np.minimum(d1, d2) < slack) & ((p1 == p2) | (d1 != d2))
                sw = ok & (d2 < d1); p1, p2 = np.where(sw, p2, p1), np.where(sw, p1, p2)
                need = ~ok
            cx = np.floor((q[:, 0] - x0) / h).astype(int); cy = np.floor((q[:, 1] - y0) / h).astype(int)
            ux = q[:, 0] - (x0 + cx * h); uy = q[:, 1] - (y0 + cy * h)
            gx = np.stack([ux, np.zeros(N), h - ux], 1); gy = np.stack([uy, np.zeros(N), h - uy], 1)
            gap = np.sqrt(gx[:, :, None] ** 2 + gy[:, None, :] ** 2)     # [N, rx, ry]
            pruned = gap > dseed[:, None, None] * PR
            cnt = np.zeros(N, int)
            for rx in range(3):
                for ry in range(3):
                    c = cellcnt[np.clip(cy + ry, 0, gh+1), np.clip(cx + rx, 0, gw+1)]
                    cnt += np.where(pruned[:, rx, ry], 0, c)
            minpr = np.where(pruned, gap, np.inf).reshape(N, 9).min(1)
            border = np.minimum(np.minimum(ux, h - ux), np.minimum(uy, h - uy)) + h
            settled = d[:, 0] < border
            L = np.where(settled, np.minimum(np.minimum(d[:, 2], minpr), border), 0.0)
            slack = np.where(need, half_rd(L), slack)
            p1 = nn.copy() if p1 is None else np.where(need, nn, p1)
            p2 = idx[:, 1].copy() if p2 is None else np.where(need, idx[:, 1], p2)
            lst = np.nonzero(need)[0]
            def cost(order):
                c = cnt[order]; pad = (-len(c)) % 32
                c = np.concatenate([c, np.zeros(pad, int)]).reshape(-1, 32)
                return int(np.ceil(c.max(1) / 2).sum())     # pair iterations
            if prev_q is None:
                tot0['maxsum'] += cost(lst); tot0['cands'] += cnt[lst].sum(); tot0['rounds'] += math.ceil(len(lst)/32)
            elif len(lst):
                tot['identity'] += cost(lst)
                tot['prev_d2'] += cost(lst[np.argsort(prev_d[lst], kind='stable')])
                tot['by_count'] += cost(lst[np.argsort(cnt[lst], kind='stable')])
                tot['ideal'] += int(np.ceil(cnt[lst].sum() / 64))
                tot['rounds'] += math.ceil(len(lst)/32); tot['cands'] += cnt[lst].sum()
            prev_q = q; prev_d = d[:, 0].copy(); npass += 1
            k, value, order = orc.select_fraction_cumsum(d[:, 0] ** 2, lam)
            if first: cur = value; first = False
            else:
                if cur - value <= 1e-6: break
                cur = value
            inl = order[:k]
            src = orc.apply_xy(src, orc.fit_rigid2d_closed(src[inl, :2], tgt[nn[inl], :2], False))
print("passes", npass)
print("first pass: pair-iters per pass-avg %.1f, rounds/pass %.2f, cands/query %.1f" % (tot0['maxsum']/npass, tot0['rounds']/npass, tot0['cands']/ (24*N)))
for k in ('identity','prev_d2','by_count','ideal'): print("seeded %-9s pair-iters per pass %.1f" % (k, tot[k]/npass))
print("seeded rounds/pass %.2f, cands per searched query %.1f" % (tot['rounds']/npass, tot['cands']/max(1,(tot['rounds']*32))))
