import sys
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from coregistrationgame_b200 import TargetIndex, IcpBatch, register_batch
from coregistrationgame_b200.matching import remove_matches, radial_crop
from ficp import FractionalICP
from oracle import ficp_oracle as orc
# small but covering: window path, global path (disable_window), deferred queries, E=1/4/16, 2D/3D, stage kernels
for dims in (2, 3):
    tgt, plots, _ = orc.synthetic_scene(20000, 130, seed=5, dims=dims, n_plots=3, hidden_pose=True, out_frac=0.2, dup_every=9)
    hyp = orc.hypothesis_table(6, flips=(0, 1), translations=[(0.0, 0.0), (3000.0, -2000.0)])
    a = register_batch(plots + [plots[0][:20], plots[1][:5]], tgt, hyp)
    ti = TargetIndex(tgt)
    b = IcpBatch(ti, plots, hyp, disable_window=True); r = b.run().results(); b.close()
    b = IcpBatch(ti, [orc.synthetic_scene(20000, 500, seed=6, dims=dims)[1][0]], hyp[:4]); r2 = b.run().results(); b.close()
    b = IcpBatch(ti, plots, hyp[:4], fixed_frac=0.7, allow_reflection=True); r3 = b.run().results(); b.close()
    icp = FractionalICP(plots[0], tgt); icp.run(); icp2 = FractionalICP(plots[1], tgt); icp2._iterate_stepwise()
    idx, d = ti.query(plots[2]); ti.close()
    print(dims, a["best_hyp"], r["stats"], r2["stats"]["passes"], r3["stats"]["passes"], icp.n_passes_, icp2.n_passes_)
chm = np.column_stack([np.random.default_rng(0).uniform(0, 50, (300, 2)), np.random.default_rng(1).uniform(8, 30, 300)])
print(remove_matches(chm[:20] + 0.1, chm, 15)[:5], len(radial_crop(chm, 25.0, 25.0, 10.0)))
print("san_probe done")
