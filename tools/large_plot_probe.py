"""One drop-in call on a plot above the persistent kernels' 1024-tree limit: device-resident stepper (run()) vs the
host-stepped stage entry points (_iterate_stepwise), same kernels, same bits.
    python tools/large_plot_probe.py"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ficp import FractionalICP
from coregistrationgame_b200 import synthetic as syn
from coregistrationgame_b200.batch import STAGE2_LAMBDA

def host_stepped(src, tgt):
    i = FractionalICP(src, tgt)
    for lam in (3.0, STAGE2_LAMBDA[i.match_dims]):
        i.lambda_val = lam
        i._iterate_stepwise()
    return i

def resident(src, tgt):
    i = FractionalICP(src, tgt)
    i.run()
    return i

for n, m in ((2000, 200_000), (12_000, 1_000_000), (100_000, 1_000_000)):
    tgt, plots, _ = syn.synthetic_scene(m, n, seed=3, dims=3, hidden_pose=False)
    src = plots[0].copy(); src[:, :2] += [0.8, -0.5]
    row = {"trees": n, "target_points": m}
    for name, fn in (("resident", resident), ("host_stepped", host_stepped)):
        fn(src, tgt)
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            i = fn(src, tgt)
        ms = (time.perf_counter() - t0) / reps * 1e3
        row[name] = {"ms_per_call": ms, "passes": i.n_passes_, "ms_per_pass": ms / max(i.n_passes_, 1), "k": i.k_}
        row.setdefault("xy", []).append(i.source[:, :2].copy())
    row["bit_identical"] = bool(np.array_equal(*row.pop("xy")))
    print(json.dumps(row), flush=True)
