"""Which part of an end-to-end register_batch step carries the occasional slow step?  Per-step wall clock of every part.
    python tools/e2e_outlier_probe.py [steps]"""
import json, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import coregistrationgame_b200.batch as B
from coregistrationgame_b200 import _lib, synthetic as syn
from coregistrationgame_b200.batch import hypothesis_table, translation_lattice

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 24
torch.cuda.set_device(0)
_lib.check(_lib.load().ficp_set_device(0))
tgt, plots, _ = syn.synthetic_scene(1_000_000, 500, seed=3, dims=3, n_plots=16, hidden_pose=True)
hyp = hypothesis_table(128, flips=(0, 1), translations=translation_lattice(4, 2.5))
def pinned(a):
    t = torch.empty(a.shape, dtype=torch.float64, pin_memory=True); t.numpy()[...] = a; return t.numpy()
h_tgt, h_plots, h_hyp = pinned(tgt), [pinned(p) for p in plots], pinned(hyp)
T = {}
def wrap(cls, name, key):
    orig = getattr(cls, name)
    def f(*a, **k):
        t0 = time.perf_counter(); r = orig(*a, **k); T[key] = T.get(key, 0.0) + (time.perf_counter() - t0) * 1e3; return r
    setattr(cls, name, f)
wrap(B.TargetIndex, "__init__", "index_create"); wrap(B.TargetIndex, "close", "index_close")
wrap(B.IcpBatch, "__init__", "batch_create"); wrap(B.IcpBatch, "close", "batch_close"); wrap(B.IcpBatch, "run", "run_enqueue")
wrap(B.IcpBatch, "best", "best(sync+readback)")
rows = []
for it in range(steps):
    T.clear()
    t0 = time.perf_counter()
    B.register_batch(h_plots, h_tgt, h_hyp, per_hypothesis=False)
    T["step"] = (time.perf_counter() - t0) * 1e3
    rows.append({k: round(v, 2) for k, v in T.items()})
for i, r in enumerate(rows):
    print(i, json.dumps(r))
