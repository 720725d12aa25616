"""Where an end-to-end step of config 4 goes (1250 plots x 150 trees, one ICP per plot, resident target index): wall clock
around the parts of register_batch((rows, offsets), index=...), for the stacked rows in pageable memory (staged by the library),
in page-locked memory (uploaded as they are, split on the device) and in page-locked memory with FICP_HOST_STAGING=1.
One JSON line per case.  PROBE_THREADS=1,2,4,8 sweeps FICP_HOST_THREADS on pageable rows instead (profiles/r02_c4_e2e_probe.jsonl).
(GPU box only.)"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coregistrationgame_b200 import IcpBatch, TargetIndex, _lib           # noqa: E402
from coregistrationgame_b200 import synthetic as syn                      # noqa: E402
from coregistrationgame_b200.batch import world_transforms                # noqa: E402


def main():
    points = int(os.environ.get("PROBE_POINTS", "1000000"))
    n_plots, trees, reps = 1250, 150, 40
    tgt, plots, _ = syn.synthetic_scene(points, trees, seed=4, dims=3, n_plots=n_plots, hidden_pose=False)
    rows = np.ascontiguousarray(np.vstack(plots))
    offs = np.concatenate([[0], np.cumsum([len(p) for p in plots])]).astype(np.int64)
    hyp = np.array([[1.0, 0.0, 0.0, 1.0, 0.0, 0.0]])
    _lib.require_device()
    index = TargetIndex(tgt)
    lib = _lib.load()
    raw_create = lib.ficp_batch_create
    clock = {"create_c": 0.0}

    def timed_create(*a):
        t = time.perf_counter()
        rc = raw_create(*a)
        clock["create_c"] += time.perf_counter() - t
        return rc
    lib.ficp_batch_create = timed_create
    import torch
    pinned = torch.empty(rows.shape, dtype=torch.float64, pin_memory=True).numpy()
    pinned[...] = rows
    pageable = rows
    if os.environ.get("PROBE_THREADS"):
        cases = [("pageable", pageable, {"FICP_HOST_THREADS": t}) for t in os.environ["PROBE_THREADS"].split(",")]
    else:
        cases = [("pageable", pageable, {}), ("pinned", pinned, {}), ("pinned, FICP_HOST_STAGING=1", pinned, {"FICP_HOST_STAGING": "1"})]
    for what, rows, env in cases:
        for k in ("FICP_HOST_THREADS", "FICP_HOST_STAGING"):
            os.environ.pop(k, None)
        os.environ.update(env)
        threads = env.get("FICP_HOST_THREADS", "0")
        acc = {k: 0.0 for k in ("init", "create_c", "run", "best", "close", "transform", "total")}
        passes = 0
        for rep in range(reps + 3):
            if rep == 3:
                acc = {k: 0.0 for k in acc}
                passes = 0
            clock["create_c"] = 0.0
            t0 = time.perf_counter()
            b = IcpBatch(index, (rows, offs), hyp, min_k=0)
            t1 = time.perf_counter()
            b.run()
            t2 = time.perf_counter()
            out = b.best()
            t3 = time.perf_counter()
            b.close()
            t4 = time.perf_counter()
            world_transforms(out["best_row"], out["best_b"])
            t5 = time.perf_counter()
            for k, v in (("init", t1 - t0), ("create_c", clock["create_c"]), ("run", t2 - t1), ("best", t3 - t2), ("close", t4 - t3),
                         ("transform", t5 - t4), ("total", t5 - t0)):
                acc[k] += v
            passes += out["stats"]["passes"]
        line = {"rows_in": what, "rows_direct": int(b.info["rows_direct"]), "host_threads": int(threads), "points": points, "plots": n_plots, "trees": trees, "reps": reps,
                "ms_per_step": {k: round(v / reps * 1e3, 4) for k, v in acc.items()},
                "python_in_init_ms": round((acc["init"] - acc["create_c"]) / reps * 1e3, 4),
                "hyp_iter_per_s": passes / acc["total"], "launch": {k: b.info[k] for k in ("cta_per_icp", "warps_per_cta", "n_ctas") if k in b.info}}
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
