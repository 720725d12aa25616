"""Multi-GPU sharding of the batched search: one process per GPU, ``torch.distributed`` (NCCL over
NVLink/NVSwitch) for the only exchange the path has - picking the best registration per plot.

Every (plot, hypothesis) ICP is independent, so the batch is cut along one of its two axes
(``shard_plan``): whole plots round-robin to the ranks when that balances (a batch of stands), else the
hypotheses of every plot (``hyp_shard=(rank, world)``: ONE stand, strong scaling).  Each rank keeps a
replica of the target grid (<= 320 MB for 1e7 points) and runs the persistent kernel on its share
with no data-path collective.

The exchange is ONE collective on device memory: every rank packs, per plot, its best key
``(fp32 score bits << 32) | hypothesis id``, the 80-byte result row of that hypothesis and its pass
count (``ficp_batch_pack_best_device``, 112 bytes per plot) and the records are all-gathered; the
winner of a plot is the record with the smallest key (ties in score resolve to the lowest
hypothesis id, exactly like the single-GPU ``atomicMin``).  Nothing is read back to the host before
the exchange; the result leaves the device once, as ``world x n_plots x 96`` bytes.  The payload
is bytes to kilobytes, i.e. latency-bound, so a fused compute+collective kernel over peer memory
would have nothing to overlap (DESIGN.md "multi-GPU").
"""
from __future__ import annotations

import numpy as np

from . import _lib
from .batch import IcpBatch, TargetIndex, decode_best_keys, world_transforms

PACK_WORDS = _lib.PACK_WORDS   # int64 words per plot record: key, 10 words of ficp_hyp_result, passes of the rank, world-frame translation (2)


def shard_of(rank, world):
    """Hypothesis shard of a rank: ids rank, rank + world, rank + 2*world, ..."""
    return (int(rank), int(world))


def plot_shard(n_plots, rank, world):
    """Plots owned by a rank when the batch is sharded over plots (fewer hypotheses than ranks, e.g. one start pose
    per plot): plot ids rank, rank + world, ..."""
    return np.arange(int(rank), int(n_plots), int(world))


def shard_plan(n_trees, n_hyp, world):
    """Which axis of the (plot, hypothesis) batch is dealt to the ranks: ``"plots"`` or ``"hypotheses"``.

    Both cuts give every rank the same kind of work; they differ in balance and in what a rank has to prepare.  A rank
    that owns whole plots uploads and stages only those plots (1/world of the host prep, of the upload and of the
    windows its kernel stages - at 8 ranks x 16 stands 1.9 -> 0.3 ms of batch creation and a ~3 % shorter kernel), so
    plots are cut whenever that is at least as balanced as cutting the hypotheses: load of a rank = trees x hypotheses of
    its round-robin share.  Fewer hypotheses than ranks (one start pose per plot) can only be cut by plots; fewer plots
    than ranks (ONE stand) only by hypotheses.  Deterministic in its arguments: every rank takes the same decision."""
    n_trees = np.asarray(n_trees, dtype=np.int64).reshape(-1)
    n_plots, n_hyp, world = int(n_trees.size), int(n_hyp), int(world)
    if world <= 1 or n_hyp < world:
        return "plots"
    if n_plots < world:
        return "hypotheses"
    load = np.array([n_trees[r::world].sum() for r in range(world)], dtype=np.float64)
    imb_plots = float(load.max() * world / max(load.sum(), 1.0))
    imb_hyp = float(-(-n_hyp // world) * world) / n_hyp
    return "plots" if imb_plots <= imb_hyp * 1.02 else "hypotheses"


def gather_packed(packed, group=None):
    """All-gather of the per-rank records: (n, PACK_WORDS) int64 -> (world, n, PACK_WORDS).  The one collective of the
    path; any backend (NCCL on device tensors, gloo on CPU tensors)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    out = torch.empty((world,) + tuple(packed.shape), dtype=packed.dtype, device=packed.device)
    if world > 1:
        dist.all_gather_into_tensor(out.view(-1), packed.contiguous().view(-1), group=group)
    else:
        out[0].copy_(packed)
    return out


def exchange_best(batch, packed, group=None, stream=None):
    """Enqueue-only exchange step of a finished (enqueued) batch: pack on the device, all-gather.  `packed` is a device
    int64 tensor (n_rows >= batch.n_plots, PACK_WORDS) whose unused rows stay as they are (callers zero them once).
    Returns the gathered device tensor (world, n_rows, PACK_WORDS).  bench.py's device-timed step and
    register_batch_distributed both call exactly this."""
    batch.pack_best_to(packed.data_ptr(), stream)
    return gather_packed(packed, group)


def select_best(gathered, by_plots=False, n_plots=None):
    """Host-side decode of the gathered records (numpy int64 (world, n_rows, PACK_WORDS)).

    by_plots=False: every rank holds a record for every plot; the winner is the smallest key.
    by_plots=True : rank r holds the records of plots r, r + world, ... (row i = plot r + world * i).
    Returns best_key (uint64), rows (HYP_RESULT_DTYPE, one per plot) and the pass counts per rank."""
    g = np.ascontiguousarray(np.asarray(gathered, dtype=np.int64))
    world, n_rows = g.shape[0], g.shape[1]
    passes = g[:, 0, 11].astype(np.int64) if n_rows else np.zeros(world, dtype=np.int64)
    if by_plots:
        n_plots = int(n_plots)
        # plot p lives in row p // world of rank p % world
        sel = g[np.arange(n_plots) % world, np.arange(n_plots) // world]
    else:
        keys = g[:, :, 0]                                   # non-negative as int64 (fp32 score bits of a value >= 0)
        win = np.argmin(keys, axis=0)                       # first minimum; equal keys cannot occur across ranks (ids differ)
        sel = g[win, np.arange(n_rows)]
    best_key = sel[:, 0].astype(np.uint64)
    rows = np.ascontiguousarray(sel[:, 1:11]).view(_lib.HYP_RESULT_DTYPE).reshape(-1)
    return best_key, rows, passes, np.ascontiguousarray(sel[:, 12:14]).view(np.float64)


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
        by_plots = shard_plan([len(p) for p in sources], hyp_table.shape[0], world) == "plots"
        if not by_plots:
            # ---- shard the hypotheses of every plot; winner = smallest key among the gathered records
            batch = IcpBatch(index, sources, hyp_table, hyp_shard=shard_of(rank, world), **kw)
            packed = torch.empty((n_plots, PACK_WORDS), dtype=torch.int64, device=dev)
        else:
            # ---- shard the PLOTS (every rank runs all hypotheses of its own plots), gather the rows
            mine = plot_shard(n_plots, rank, world)
            centres_all = kw.pop("centres", None)
            centres_mine = None if centres_all is None else np.asarray(centres_all, dtype=np.float64).reshape(n_plots, 2)[mine]
            n_rows = (n_plots + world - 1) // world
            packed = torch.zeros((n_rows, PACK_WORDS), dtype=torch.int64, device=dev)
            batch = IcpBatch(index, [sources[p] for p in mine], hyp_table, centres=centres_mine, **kw) if len(mine) else None
        try:
            if batch is not None:
                batch.run(stream)
                gathered = exchange_best(batch, packed, group, stream)
                h2d = batch.h2d_bytes
            else:
                gathered = gather_packed(packed, group)
                h2d = 0
            g = gathered.cpu().numpy()                       # the one read-back: world x rows x 112 bytes
        finally:
            if batch is not None:
                batch.close()
        gk, rows, passes, b = select_best(g, by_plots=by_plots, n_plots=n_plots)
        res = decode_best_keys(gk)
        res["best_key"] = gk
        res["best_row"] = rows
        res["best_transform"] = world_transforms(rows, b)
        res["k"] = rows["k"].astype(np.int64)
        res["rmse"], res["frmsd"] = rows["rmse"].copy(), rows["frmsd"].copy()
        res["passes_local"] = int(passes[rank])
        res["passes_global"] = int(passes.sum())
        res["h2d_bytes"] = h2d + (int(np.asarray(target).nbytes) if own else 0)
        res["d2h_bytes"] = int(g.nbytes)
        return res
    finally:
        if own:
            index.close()
