"""Per-phase cycle shares of the CTA-per-ICP kernel (diagnostic build):
    make -C coregistrationgame_b200/csrc variant NAME=clk FLAGS=-DFICP_PHASE_CLOCKS
    FICP_B200_LIB=$PWD/coregistrationgame_b200/libficp_clk.so python tools/team_phase_clocks.py [world] [ctas_per_sm]
Thread 0 of every CTA accumulates clock64() deltas between phase marks (icp_team.cu, PHASE(n))."""
import ctypes as C, json, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coregistrationgame_b200 import IcpBatch, TargetIndex, _lib, synthetic as syn
from coregistrationgame_b200.batch import hypothesis_table, translation_lattice

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
cps = int(sys.argv[2]) if len(sys.argv) > 2 else 0
names = ["start/stage", "skip test+list", "search", "deferred", "order check+repair", "block sort", "scan", "arg-min",
         "fit prep", "f ‖ fit sums (warp 0)", "solve+barrier+logic", "results", "tail",
         "search: up to the merged group result", "skip test: test itself", "-"]
lib = _lib.load()
fn = lib.ficp_debug_phase_clocks
fn.argtypes = [C.POINTER(C.c_uint64), C.c_int]
tgt, plots, _ = syn.synthetic_scene(1_000_000, 500, seed=3, dims=3, n_plots=1, hidden_pose=True)
hyp = hypothesis_table(128, flips=(0, 1), translations=translation_lattice(4, 2.5))
ti = TargetIndex(tgt)
stream = torch.cuda.current_stream()
b = IcpBatch(ti, [plots[0]], hyp, hyp_shard=(0, world), cta_per_icp=True, ctas_per_sm=cps)
for _ in range(2):
    b.run(stream)
torch.cuda.synchronize()
buf = (C.c_uint64 * 16)()
fn(buf, 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream); b.run(stream); e1.record(stream)
torch.cuda.synchronize()
fn(buf, 0)
st = b.results(per_hypothesis=False)["stats"]
clk = np.array(list(buf)[:16], dtype=np.float64)
tot = clk.sum()
out = {"world": world, "ms": e0.elapsed_time(e1), "passes": st["passes"], "ctas": b.info["ctas"], "ctas_per_sm": b.info["ctas_per_sm"],
       "cycles_per_pass_per_cta": tot / st["passes"],
       "phases": {n: {"share": round(float(c / tot), 4), "cycles_per_pass": round(float(c / st["passes"]), 1)} for n, c in zip(names, clk)}}
print(json.dumps(out, indent=1))
