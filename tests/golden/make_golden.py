"""Generate golden vectors by running the UNMODIFIED reference ``/root/reference/ficp.py``.

Run once in the build container (the reference does not travel to the GPU box):

    python tests/golden/make_golden.py

Writes ``tests/golden/*.npz``.  Each file holds the inputs and, for every NN pass the reference
made (one ``find_correspondences`` call = one hypothesis-iteration): the NN index, the NN distance,
the trimmed subset size ``k`` and the FRMSD value, plus the final aligned array and the value of
``lambda_val`` after ``run()``.  The recording subclass only observes; it calls the reference's
own methods for every number it stores.
"""
import os
import sys

import numpy as np
import pandas as pd
from scipy.spatial import cKDTree

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, "/root/reference")
sys.path.insert(1, REPO)

from ficp import FractionalICP as RefICP  # noqa: E402  (the reference)
from oracle import ficp_oracle as orc  # noqa: E402  (only for scene generators)


class Recorder(RefICP):
    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self.rec_idx, self.rec_d, self.rec_k, self.rec_val = [], [], [], []
        self._last = None

    def find_correspondences(self, source, target):
        corr, d = super().find_correspondences(source, target)
        if len(d):
            md = self.match_dims
            _, idx = cKDTree(np.ascontiguousarray(target[:, :md])).query(
                np.ascontiguousarray(source[:, :md]), k=1)
            assert np.array_equal(target[idx], corr)
            self._last = (idx.astype(np.int32), d.copy())
        return corr, d

    def find_optimal_fraction(self, corr, d):
        frac, k = super().find_optimal_fraction(corr, d)
        if self._last is not None and len(d):
            idx, dd = self._last
            sel = np.argsort(d)[:k]
            val = self.frmsd(frac, k, self.source[sel], corr[sel])
            self.rec_idx.append(idx)
            self.rec_d.append(dd)
            self.rec_k.append(k)
            self.rec_val.append(val)
        return frac, k


def record(name, source, target, **kw):
    r = Recorder(source, target, **kw)
    aligned = r.run()
    n = r.source.shape[0]
    out = dict(
        source=np.asarray(source, dtype=float), target=np.asarray(target, dtype=float),
        aligned=aligned, lambda_after=np.float64(r.lambda_val), match_dims=np.int32(r.match_dims),
        idx=np.array(r.rec_idx, dtype=np.int32).reshape(-1, n),
        dist=np.array(r.rec_d, dtype=np.float64).reshape(-1, n),
        k=np.array(r.rec_k, dtype=np.int32), val=np.array(r.rec_val, dtype=np.float64),
        allow_reflection=np.bool_(kw.get("allow_reflection", False)),
        lambda_val=np.float64(kw.get("lambda_val", 3.0)),
    )
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(f"{name}: N={n} M={len(target)} md={r.match_dims} passes={len(r.rec_k)} k_final={r.rec_k[-1] if r.rec_k else 0}")


# ---- clouds shaped like the reference's own tests (tests/test_ficp.py:12-23) ----
def make_cloud(n, seed):
    rng = np.random.default_rng(seed)
    xy = rng.normal(size=(n, 2)) @ np.array([[1.0, 0.3], [0.0, 0.6]]).T
    z = np.linspace(0.0, 20.0, n).reshape(-1, 1) + rng.normal(scale=0.02, size=(n, 1))
    return np.hstack([xy, z])


def rigid(src, angle_deg, t):
    th = np.deg2rad(angle_deg)
    r = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
    return np.hstack([src[:, :2] @ r.T + np.asarray(t), src[:, 2:]])


def main():
    # A/B: reference-test shapes, 3D and 2D
    src = make_cloud(150, 1)
    tgt = rigid(src, 27.0, [1.6, -2.2])
    record("ref_basic_3d", src, tgt)
    record("ref_basic_2d", src[:, :2], tgt[:, :2])

    src = make_cloud(200, 2)
    full = rigid(src, 31.0, [2.5, -1.8])
    keep = np.random.default_rng(123).choice(200, 100, replace=False)
    record("ref_missing_3d", src, full[keep])
    record("ref_missing_2d", src[:, :2], full[keep][:, :2])

    src = make_cloud(200, 3)
    clean = rigid(src, -22.0, [-1.2, 2.0])
    rng = np.random.default_rng(7)
    keep = rng.choice(200, 100, replace=False)
    t = clean[keep]
    no = int(0.3 * len(t))
    t = np.vstack([t, np.hstack([rng.uniform(-20, 20, (no, 2)), rng.uniform(-5, 25, (no, 1))])])
    record("ref_outliers_3d", src, t)
    record("ref_outliers_refl_3d", src, t, allow_reflection=True)
    record("ref_outliers_lam1_2d", src[:, :2], t[:, :2], lambda_val=1.0)

    # C: real data, the way app.py mode 2 loads it (SavedStand / SavedPlot, dist=70): 2D
    d14 = pd.read_csv("/root/reference/Data/2014/Stand_10_trees.csv")
    d19 = pd.read_csv("/root/reference/Data/2019/Stand_10_trees.csv")
    plot_centres = d14.groupby("PlotID", sort=False)[["CurrentX", "CurrentY"]].mean().values
    centre = plot_centres.mean(axis=0)  # Stand._update_center: mean of plot centres
    xy19 = d19[["CurrentX", "CurrentY"]].values
    tgt19 = xy19[np.sqrt(((xy19 - centre) ** 2).sum(1)) <= 70.0]
    srcs, offs = [], [0]
    for pid, g in d14.groupby("PlotID", sort=False):
        s = g[["CurrentX", "CurrentY"]].values.astype(float)
        srcs.append(s)
        offs.append(offs[-1] + len(s))
    aligned, passes, kfin = [], [], []
    for s in srcs:
        r = Recorder(s, tgt19)
        aligned.append(r.run())
        passes.append(len(r.rec_k))
        kfin.append(r.rec_k[-1])
    np.savez_compressed(os.path.join(HERE, "c1_real_2d.npz"), source=np.vstack(srcs),
                        offsets=np.array(offs, dtype=np.int64), target=tgt19,
                        aligned=np.vstack(aligned), passes=np.array(passes, dtype=np.int32),
                        k_final=np.array(kfin, dtype=np.int32))
    print(f"c1_real_2d: plots={len(srcs)} M={len(tgt19)} passes={passes}")

    # D: synthetic scenes per SURVEY 8(d), small
    for dims in (2, 3):
        tgt, plots, _ = orc.synthetic_scene(2000, 60, seed=11, dims=dims, hidden_pose=True)
        hyp = orc.hypothesis_table(8, flips=(0, 1))
        c = plots[0][:, :2].mean(axis=0)
        for h in (0, 3, 9):
            record(f"syn_d{dims}_h{h}", orc.pre_transform(plots[0], hyp[h], c), tgt)

    # E: adversarial - outliers, omissions, duplicated targets, lattice ties
    tgt, plots, _ = orc.synthetic_scene(3000, 80, seed=5, dims=3, out_frac=0.3, omit_frac=0.3,
                                        dup_every=10, lattice_patch=6, hidden_pose=True)
    record("adv_d3", plots[0], tgt)
    tgt, plots, _ = orc.synthetic_scene(3000, 80, seed=6, dims=2, out_frac=0.3, omit_frac=0.3,
                                        dup_every=10, lattice_patch=6, hidden_pose=False)
    record("adv_d2", plots[0], tgt)


if __name__ == "__main__":
    main()
