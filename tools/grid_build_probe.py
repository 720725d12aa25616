"""Times the grid build (device time of the whole chain, CUDA events inside the library) for 1e5 / 1e6 / 1e7 points,
XY and XYZ, uniform and skewed.  python tools/grid_build_probe.py [reps]"""
import sys, os, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coregistrationgame_b200 import TargetIndex, synthetic as syn
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
for m in (100_000, 1_000_000, 10_000_000):
    tgt, _, _ = syn.synthetic_scene(m, 50, seed=4, dims=3, n_plots=1, hidden_pose=False)
    for dims, ppc in ((3, 6.0), (2, 3.0)):
        arr = np.ascontiguousarray(tgt[:, :dims])
        ms = []
        for _ in range(reps):
            ti = TargetIndex(arr, pts_per_cell=ppc)
            info = ti.info()
            ms.append(info["build_ms"])
            ti.close()
        print(json.dumps({"points": m, "dims": dims, "build_ms_min": min(ms), "build_ms_median": float(np.median(ms)),
                          "alg_GBps_at_48B": m * 48.0 / (min(ms) * 1e-3) / 1e9, "grid": [info["grid_w"], info["grid_h"]],
                          "max_cell_pts": info["max_cell_pts"], "clamped": info["clamped"]}), flush=True)
