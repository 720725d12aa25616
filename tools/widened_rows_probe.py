"""Timings of the paths around the hot loop that had no number yet:
  * a plot above the persistent kernels' 1024-tree limit (host-stepped stage kernels), one drop-in call;
  * greedy match-and-remove (CHMPlot.remove_matches) for one plot and for a batch of plots against a resident index;
  * radial crop of the CHM layer.
    python tools/widened_rows_probe.py"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ficp import FractionalICP
from coregistrationgame_b200 import TargetIndex, synthetic as syn
from coregistrationgame_b200.matching import radial_crop, remove_matches, remove_matches_batch

def timed(fn, reps):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        r = fn()
    return (time.perf_counter() - t0) / reps * 1e3, r

out = {}
tgt, plots, _ = syn.synthetic_scene(200_000, 2000, seed=3, dims=3, hidden_pose=False)
src = plots[0].copy(); src[:, :2] += [0.8, -0.5]
ms, icp = timed(lambda: (lambda i: (i.run(), i)[1])(FractionalICP(src, tgt)), 3)
out["stepwise_2000_trees_vs_2e5"] = {"ms_per_call": ms, "passes": icp.n_passes_, "ms_per_pass": ms / max(icp.n_passes_, 1)}
tgt7, plots7, _ = syn.synthetic_scene(10_000_000, 150, seed=4, dims=3, n_plots=1250, hidden_pose=False)
idx = TargetIndex(tgt7, purpose="query")
ms1, m1 = timed(lambda: remove_matches(plots7[0], tgt7, 15, index=idx), 20)
msb, mb = timed(lambda: remove_matches_batch(plots7, tgt7, 15, index=idx), 5)
out["remove_matches_one_plot_150_trees_vs_1e7_resident_index"] = {"ms": ms1, "matched": int((m1 >= 0).sum())}
out["remove_matches_batch_1250_plots_resident_index"] = {"ms": msb, "trees": 1250 * 150, "trees_per_s": 1250 * 150 / (msb * 1e-3), "matched": int(sum((m >= 0).sum() for m in mb))}
c = tgt7[:, :2].mean(0)
msc, rows = timed(lambda: radial_crop(idx, c[0], c[1], 70.0), 5)
out["radial_crop_70m_of_1e7"] = {"ms": msc, "rows": int(len(rows))}
idx.close()
print(json.dumps(out, indent=1))
