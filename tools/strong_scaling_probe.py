"""Per-GPU time of the literal BASELINE config 3 (ONE stand x 4096 start poses) when its hypotheses are sharded over
w = 1, 2, 4, 8 GPUs, measured on ONE GPU by running rank 0's shard (hyp_shard=(0, w)): kernel shape A/B
(warp-per-ICP elastic kernel vs CTA-per-ICP kernel).  The multi-GPU job adds one 8-byte NCCL all-reduce(MIN).

    python tools/strong_scaling_probe.py [--dims 3] [--trees 500] [--points 1000000] [--reps 5]
"""
import argparse, json, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coregistrationgame_b200 import IcpBatch, TargetIndex, synthetic as syn
from coregistrationgame_b200.batch import hypothesis_table, translation_lattice

ap = argparse.ArgumentParser()
ap.add_argument("--dims", type=int, default=3)
ap.add_argument("--trees", type=int, default=500)
ap.add_argument("--points", type=int, default=1_000_000)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--worlds", default="1,2,4,8")
ap.add_argument("--kernels", default="warp,cta,cta1,auto")
args = ap.parse_args()
tgt, plots, _ = syn.synthetic_scene(args.points, args.trees, seed=3, dims=args.dims, n_plots=1, hidden_pose=True)
hyp = hypothesis_table(128, flips=(0, 1), translations=translation_lattice(4, 2.5))
ti = TargetIndex(tgt)
stream = torch.cuda.current_stream()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
base = None
for w in [int(x) for x in args.worlds.split(",")]:
    for name, kw in (("warp", dict(cta_per_icp=False)), ("cta", dict(cta_per_icp=True)), ("cta1", dict(cta_per_icp=True, ctas_per_sm=1)),
                     ("auto", dict())):
        if name not in args.kernels.split(","):
            continue
        b = IcpBatch(ti, [plots[0]], hyp, hyp_shard=(0, w), **kw)
        for _ in range(2):
            b.run(stream)
        torch.cuda.synchronize()
        ms = []
        for _ in range(args.reps):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream); b.run(stream); e1.record(stream)
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        out = b.results(per_hypothesis=True)
        row = out["hyp"][0]
        if base is None and w == 1:
            base = row.copy()
        same = bool(all(np.array_equal(row[f], base[::w][f]) for f in row.dtype.names if f not in ("flags", "pad"))) if base is not None else None
        print(json.dumps({"world": w, "kernel": name, "ms_median": float(np.median(ms)), "ms_min": float(min(ms)),
                          "passes": out["stats"]["passes"], "icps": int(b.n_hyp_local), "longest": int(row["passes"].max()),
                          "hyp_iter_per_s": out["stats"]["passes"] / (np.median(ms) * 1e-3), "bit_identical_to_w1_warp": same,
                          "searched": out["stats"]["searched_queries"], "fixups": out["stats"]["fixup_rounds"], "order_rebuilds": out["stats"].get("order_rebuilds"),
                          "launch": {k: b.info[k] for k in ("cta_per_icp", "warps_per_cta", "ctas", "ctas_per_sm", "team_warps", "helpers", "smem_bytes")}}), flush=True)
        b.close()
ti.close()
