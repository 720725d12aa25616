"""Multi-GPU sharding of the batched search: one process per GPU, ``torch.distributed`` (NCCL over
NVLink/NVSwitch) for the only exchange the path has - picking the best registration per plot.

Every (plot, hypothesis) ICP is independent, so the hypotheses of every plot are dealt round-robin
to the ranks (``hyp_shard=(rank, world)``), each rank keeps a replica of the target grid (<= 280 MB
for 1e7 points) and runs the persistent kernel on its share with no data-path collective.  The
exchange step is one all-reduce(MIN) over the packed per-plot keys ``(fp32 score bits << 32) |
hypothesis id`` (8 bytes per plot) followed by one all-reduce(SUM) that carries the winner's pose
from the rank that owns it.  The payload is bytes to kilobytes, i.e. latency-bound, so a fused
compute+collective kernel over peer memory would buy nothing here (DESIGN.md "multi-GPU").
"""
from __future__ import annotations

import numpy as np

from .batch import IcpBatch, TargetIndex, compose_world_transform, decode_best_keys

DETAIL_FIELDS = ("m00", "m01", "m10", "m11", "cx", "cy", "frmsd", "rmse", "k", "passes")


def reduce_best(local_keys, local_detail, group=None):
    """Pick the global winner per plot.

    local_keys   int64 tensor (n_plots,), this rank's best packed key per plot (non-negative).
    local_detail float64 tensor (n_plots, D): payload describing this rank's best hypothesis per plot.
    Returns (global_keys, global_detail) - identical on every rank.  Works on any backend
    (NCCL on GPU tensors, gloo on CPU tensors)."""
    import torch
    import torch.distributed as dist

    keys = local_keys.clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(keys, op=dist.ReduceOp.MIN, group=group)
        mine = (local_keys == keys).to(local_detail.dtype).unsqueeze(1)
        detail = local_detail * mine
        dist.all_reduce(detail, op=dist.ReduceOp.SUM, group=group)
    else:
        detail = local_detail.clone()
    return keys, detail


def shard_of(rank, world):
    """Hypothesis shard of a rank: ids rank, rank + world, rank + 2*world, ..."""
    return (int(rank), int(world))


def plot_shard(n_plots, rank, world):
    """Plots owned by a rank when the batch is sharded over plots (fewer hypotheses than ranks, e.g. one start pose
    per plot): plot ids rank, rank + world, ..."""
    return np.arange(int(rank), int(n_plots), int(world))


def gather_plot_shards(local_detail, local_plots, n_plots, group=None):
    """Assemble per-plot rows computed on different ranks: every rank contributes the rows of the plots it owns,
    everything else is zero, one all-reduce(SUM) puts the table together on all ranks (no arg-min needed)."""
    import torch
    import torch.distributed as dist

    full = torch.zeros((int(n_plots), local_detail.shape[1]), dtype=local_detail.dtype, device=local_detail.device)
    if len(local_plots):
        full[torch.as_tensor(np.asarray(local_plots), device=local_detail.device, dtype=torch.long)] = local_detail
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(full, op=dist.ReduceOp.SUM, group=group)
    return full


def local_best_detail(batch, out):
    """(n_plots, len(DETAIL_FIELDS)) float64 rows of this rank's best hypothesis per plot."""
    det = np.zeros((batch.n_plots, len(DETAIL_FIELDS)), dtype=np.float64)
    for p in range(batch.n_plots):
        j = (int(out["best_hyp"][p]) - batch.hyp_begin) // batch.hyp_stride
        row = out["hyp"][p, j]
        det[p] = [float(row[f]) for f in DETAIL_FIELDS]
    return det


def register_batch_distributed(sources, target, hyp_table, index=None, group=None, device=None, **kw):
    """``register_batch`` sharded over the ranks of an initialised process group (one GPU each).

    Returns the same winner on every rank: best_hyp, best_score, best_transform (n_plots, 2, 3), k,
    rmse, frmsd, plus this rank's pass count and the global pass count."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    hyp_table = np.asarray(hyp_table, dtype=np.float64).reshape(-1, 6)
    if isinstance(sources, np.ndarray) and sources.ndim == 2:
        sources = [sources]
    own = index is None
    if own:
        index = TargetIndex(target)
    try:
        dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        stream = torch.cuda.current_stream()
        n_plots = len(sources)
        if hyp_table.shape[0] >= world:
            # ---- shard the hypotheses of every plot; winner = all-reduce(MIN) over packed keys
            batch = IcpBatch(index, sources, hyp_table, hyp_shard=shard_of(rank, world), **kw)
            try:
                batch.run(stream)
                out = batch.results(stream)
                keys = torch.from_numpy(out["best_key"].astype(np.int64)).to(dev)
                detail = torch.from_numpy(local_best_detail(batch, out)).to(dev)
                centres = batch.centres
                h2d, d2h, local_passes = batch.h2d_bytes, batch.d2h_bytes, out["stats"]["passes"]
            finally:
                batch.close()
            gkeys, gdetail = reduce_best(keys, detail, group)
            gk = gkeys.cpu().numpy().astype(np.uint64)
            gd = gdetail.cpu().numpy()
        else:
            # ---- fewer hypotheses than ranks (e.g. one start pose per plot): shard the PLOTS, gather the rows
            mine = plot_shard(n_plots, rank, world)
            centres_all = kw.pop("centres", None)
            if centres_all is None:
                centres_all = np.array([np.asarray(s, dtype=np.float64)[:, :2].mean(axis=0) for s in sources])
            centres = np.asarray(centres_all, dtype=np.float64).reshape(n_plots, 2)
            ncol = len(DETAIL_FIELDS) + 2   # + the packed key as two exactly-representable halves
            if len(mine):
                batch = IcpBatch(index, [sources[p] for p in mine], hyp_table, centres=centres[mine], **kw)
                try:
                    batch.run(stream)
                    out = batch.results(stream)
                    k64 = out["best_key"].astype(np.uint64)
                    local = np.concatenate([local_best_detail(batch, out), (k64 >> np.uint64(32)).astype(np.float64)[:, None],
                                            (k64 & np.uint64(0xFFFFFFFF)).astype(np.float64)[:, None]], axis=1)
                    h2d, d2h, local_passes = batch.h2d_bytes, batch.d2h_bytes, out["stats"]["passes"]
                finally:
                    batch.close()
            else:
                local, h2d, d2h, local_passes = np.zeros((0, ncol)), 0, 0, 0
            full = gather_plot_shards(torch.from_numpy(local).to(dev), mine, n_plots, group).cpu().numpy()
            gd = full[:, :len(DETAIL_FIELDS)]
            gk = (full[:, -2].astype(np.uint64) << np.uint64(32)) | full[:, -1].astype(np.uint64)
        passes = torch.tensor([local_passes], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(passes, op=dist.ReduceOp.SUM, group=group)
        res = decode_best_keys(gk)
        res["best_key"] = gk
        rows = {f: gd[:, i] for i, f in enumerate(DETAIL_FIELDS)}
        res["best_transform"] = np.stack([compose_world_transform({f: rows[f][p] for f in DETAIL_FIELDS}, centres[p])
                                          for p in range(n_plots)])
        res["k"] = rows["k"].astype(np.int64)
        res["rmse"], res["frmsd"] = rows["rmse"], rows["frmsd"]
        res["passes_local"] = local_passes
        res["passes_global"] = int(passes.item())
        res["h2d_bytes"] = h2d + (int(np.asarray(target).nbytes) if own else 0)
        res["d2h_bytes"] = d2h + int(gk.nbytes + gd.nbytes)
        return res
    finally:
        if own:
            index.close()
