"""Pass counts per hypothesis of the literal C3 stand (one plot x 4096 hypotheses) - the serial chain that bounds
strong scaling (DESIGN.md "strong scaling").  Run on the GPU box: python tools/pass_hist.py [out.npy]"""
import sys, os, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coregistrationgame_b200 import IcpBatch, TargetIndex, synthetic as syn
from coregistrationgame_b200.batch import hypothesis_table, translation_lattice

dims = 3
tgt, plots, _ = syn.synthetic_scene(1_000_000, 500, seed=3, dims=dims, n_plots=1, hidden_pose=True)
hyp = hypothesis_table(128, flips=(0, 1), translations=translation_lattice(4, 2.5))
ti = TargetIndex(tgt)
b = IcpBatch(ti, [plots[0]], hyp)
out = b.run().results()
p = out["hyp"]["passes"][0]
np.save(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/r02_passes_per_hyp.npy", p)
q = np.percentile(p, [50, 90, 99, 99.9])
print(json.dumps({"n": int(p.size), "sum": int(p.sum()), "mean": float(p.mean()), "p50": q[0], "p90": q[1], "p99": q[2],
                  "p999": q[3], "max": int(p.max()), "top": np.sort(p)[-12:].tolist(), "launch": b.info}))
