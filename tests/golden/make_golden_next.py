"""Golden vectors for the steps right after the ICP, produced by the UNMODIFIED reference
classes: ``CHMPlot.remove_matches`` (chm_plot.py:223-285) and ``Plot.get_transform`` (trees.py:248-280).

    python tests/golden/make_golden_next.py        # build container only

matplotlib (imported by chm_plot.py for a preview plot, absent here) is stubbed the same way the reference's own
test stubs pynput (tests/test_transformation_serialization.py:10-15)."""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")
for name in ("matplotlib", "matplotlib.pyplot"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]

from chm_plot import CHMPlot  # noqa: E402
from trees import Plot, Tree  # noqa: E402


def mk_tree(i, x, y, h):
    t = Tree(i, float(x), float(y), stemdiam_cm=20.0, height_dm=(None if not np.isfinite(h) else float(h) * 10.0))
    if not np.isfinite(h):
        t.height = float("nan")
    return t


def run_remove(plot_xyh, chm_xyh, pct):
    chm = CHMPlot.__new__(CHMPlot)
    chm.trees = [mk_tree(i, *r) for i, r in enumerate(chm_xyh)]
    chm.removed_stems = []
    ids = {id(t): i for i, t in enumerate(chm.trees)}
    plot = Plot(plotid=1)
    for i, r in enumerate(plot_xyh):
        plot.append_tree(mk_tree(i, *r))
    # heights as given (append_tree keeps them); NaN heights stay NaN
    before = list(chm.trees)
    chm.remove_matches(plot, pct)
    removed = [ids[id(t)] for t in chm.removed_stems[-1]]
    remaining = [ids[id(t)] for t in chm.trees]
    assert sorted(removed + remaining) == list(range(len(before)))
    return np.array(removed, dtype=np.int64), np.array(remaining, dtype=np.int64)


def main():
    rng = np.random.default_rng(42)
    cases = {}
    # A: 3-D, plot trees near CHM trees + noise, some too far, two plot trees competing for one CHM tree
    chm = np.column_stack([rng.uniform(0, 60, 200), rng.uniform(0, 60, 200), rng.uniform(8, 30, 200)])
    pick = rng.choice(200, 25, replace=False)
    plot = chm[pick] + np.column_stack([rng.normal(0, 0.8, 25), rng.normal(0, 0.8, 25), rng.normal(0, 0.7, 25)])
    plot[3] = plot[2] + [0.05, -0.03, 0.1]          # competes with tree 2 for the same CHM tree
    plot[7, :2] += 30.0                              # far from everything -> no match
    cases["a3d"] = (plot, chm, 15)
    cases["a3d_pct40"] = (plot, chm, 40)
    # B: a missing plot height -> XY fallback with the 10 m default
    plot_b = plot.copy()
    plot_b[5, 2] = np.nan
    cases["b2d_nan_plot"] = (plot_b, chm, 15)
    # C: a missing CHM height -> XY fallback
    chm_c = chm.copy()
    chm_c[11, 2] = np.nan
    cases["c2d_nan_chm"] = (plot, chm_c, 15)
    # D: exact duplicates in the CHM layer and more plot trees than CHM trees
    chm_d = np.vstack([chm[:6], chm[:6]])
    plot_d = np.vstack([chm[:6], chm[:6], chm[:3]]) + [0.01, 0.0, 0.0]
    cases["d_dups_exhaust"] = (plot_d, chm_d, 15)
    for name, (p, c, pct) in cases.items():
        removed, remaining = run_remove(p, c, pct)
        np.savez_compressed(os.path.join(HERE, f"next_remove_{name}.npz"), plot=p, chm=c, pct=np.float64(pct),
                            removed=removed, remaining=remaining)
        print(name, "removed", len(removed), "of", len(c))

    # get_transform: translate / rotate / flip sequences on a plot (Plot.get_transform is pure numpy)
    recs = []
    for k in range(6):
        plot = Plot(plotid=k)
        n = int(rng.integers(3, 30))
        xy = rng.uniform(-20, 20, (n, 2)) + [420100.0, 6483100.0]
        for i, (x, y) in enumerate(xy):
            plot.append_tree(Tree(i, float(x), float(y)))
        plot.translate_plot((float(rng.uniform(-5, 5)), float(rng.uniform(-5, 5))))
        plot.rotate_plot(float(rng.uniform(-180, 180)))
        if k % 2:
            plot.coordinate_flip()
        plot.translate_plot((float(rng.uniform(-2, 2)), float(rng.uniform(-2, 2))))
        R, t, flipped = plot.get_transform()
        cur = plot.get_tree_current_array()[:, 1:3].astype(float)
        recs.append(dict(orig=xy, cur=cur, R=R, t=t, flipped=np.bool_(flipped)))
    np.savez_compressed(os.path.join(HERE, "next_get_transform.npz"),
                        **{f"{key}_{i}": r[key] for i, r in enumerate(recs) for key in r}, n=np.int64(len(recs)))
    print("get_transform cases", len(recs))


if __name__ == "__main__":
    main()
