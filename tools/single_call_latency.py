"""Latency of ONE drop-in call `FractionalICP(source, target).run()` (what app.py:658-661 does on the J key),
CUDA path vs the reference algorithm restated on the host CPU (oracle port, as shipped: kd-tree rebuilt every pass)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ficp import FractionalICP  # noqa: E402
from coregistrationgame_b200.synthetic import synthetic_scene  # noqa: E402
from oracle import ficp_oracle as orc  # noqa: E402  (CPU comparison leg only)


def timed(fn, reps):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps * 1e3


def main():
    g = np.load(os.path.join(ROOT, "tests", "golden", "c1_real_2d.npz"))
    offs = g["offsets"]
    cases = [("C1 real plot (N=%d, M=%d, XY)" % (offs[1] - offs[0], len(g["target"])), g["source"][offs[0]:offs[1]], g["target"], 200, 20)]
    for m, n, dims, reps, creps in ((100_000, 200, 3, 50, 2), (1_000_000, 500, 3, 20, 1)):
        tgt, plots, _ = synthetic_scene(m, n, seed=3, dims=dims, hidden_pose=False)
        rng = np.random.default_rng(0)
        th = np.radians(4.0)
        src = plots[0].copy()
        c = src[:, :2].mean(0)
        src[:, :2] = (src[:, :2] - c) @ np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]]).T + c + [1.0, -0.7]
        cases.append((f"synthetic (N={n}, M={m}, {'XYZ' if dims == 3 else 'XY'})", src, tgt, reps, creps))
    for name, src, tgt, reps, creps in cases:
        gpu_ms = timed(lambda: FractionalICP(src, tgt).run(), reps)
        icp = FractionalICP(src, tgt)
        icp.run()
        cpu_ms = timed(lambda: orc.ficp_run(src, tgt, nn="reference", hoist_tree=False, pairwise=True), creps)
        print(f"{name}: drop-in {gpu_ms:.2f} ms/call ({icp.n_passes_} passes), reference algorithm on CPU {cpu_ms:.1f} ms/call "
              f"-> {cpu_ms / gpu_ms:.0f}x")


if __name__ == "__main__":
    main()
