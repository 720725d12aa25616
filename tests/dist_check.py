"""Run under torchrun on >= 2 GPUs: the sharded search must pick exactly the winner the single-GPU search picks.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/dist_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coregistrationgame_b200 import _lib, register_batch  # noqa: E402
from coregistrationgame_b200.dist import register_batch_distributed  # noqa: E402
from oracle import ficp_oracle as orc  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    _lib.check(_lib.load().ficp_set_device(local))
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    tgt, plots, _ = orc.synthetic_scene(200000, 200, seed=7, dims=3, n_plots=5, hidden_pose=True)
    hyp = orc.hypothesis_table(32, flips=(0, 1), translations=orc.translation_lattice(2, 2.5))
    d = register_batch_distributed(plots, tgt, hyp)
    s = register_batch(plots, tgt, hyp)
    np.testing.assert_array_equal(d["best_key"], s["best_key"])
    np.testing.assert_array_equal(d["best_hyp"], s["best_hyp"])
    np.testing.assert_array_equal(d["best_transform"], s["best_transform"])
    np.testing.assert_array_equal(d["k"], s["best_row"]["k"])
    assert d["passes_global"] == s["stats"]["passes"], (d["passes_global"], s["stats"]["passes"])
    assert d["passes_local"] < s["stats"]["passes"]
    # one start pose per plot (config 4 shape): sharded over PLOTS, rows gathered
    tgt2, plots2, _ = orc.synthetic_scene(200000, 150, seed=9, dims=3, n_plots=21, hidden_pose=False)
    rng = np.random.default_rng(5)
    starts = [orc.pre_transform(p, np.r_[orc.hypothesis_matrix(rng.uniform(-5, 5), 0).ravel(), rng.uniform(-1, 1, 2)], p[:, :2].mean(0)) for p in plots2]
    ident = np.array([[1.0, 0.0, 0.0, 1.0, 0.0, 0.0]])
    d2 = register_batch_distributed(starts, tgt2, ident, min_k=0)
    s2 = register_batch(starts, tgt2, ident, min_k=0)
    np.testing.assert_array_equal(d2["best_key"], s2["best_key"])
    np.testing.assert_array_equal(d2["best_transform"], s2["best_transform"])
    np.testing.assert_array_equal(d2["k"], s2["best_row"]["k"])
    assert d2["passes_global"] == s2["stats"]["passes"]
    # a batch of stands (one per rank and more): sharded over PLOTS with every hypothesis on the owner (dist.shard_plan)
    from coregistrationgame_b200.dist import shard_plan
    tgt3, plots3, _ = orc.synthetic_scene(200000, 120, seed=11, dims=3, n_plots=2 * world, hidden_pose=True)
    assert shard_plan([len(p) for p in plots3], hyp.shape[0], world) == "plots"
    d3 = register_batch_distributed(plots3, tgt3, hyp)
    s3 = register_batch(plots3, tgt3, hyp)
    np.testing.assert_array_equal(d3["best_key"], s3["best_key"])
    np.testing.assert_array_equal(d3["best_hyp"], s3["best_hyp"])
    np.testing.assert_array_equal(d3["best_transform"], s3["best_transform"])
    np.testing.assert_array_equal(d3["k"], s3["best_row"]["k"])
    assert d3["passes_global"] == s3["stats"]["passes"], (d3["passes_global"], s3["stats"]["passes"])
    dist.barrier()
    if rank == 0:
        print(f"dist_check ok: world={world} plots={len(plots)} hyps={hyp.shape[0]} winners={d['best_hyp'].tolist()} "
              f"passes global={d['passes_global']}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
