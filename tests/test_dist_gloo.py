"""Multi-rank host logic on CPU: world_size-2 gloo process group, hypothesis sharding and the
best-of-hypotheses exchange (the only collective on the path)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import ficp_oracle as orc


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _pack(keys, rank, n_plots, passes):
    """What ficp_batch_pack_best_device writes: key, ten row words (here: a recognisable payload), passes of the rank."""
    from coregistrationgame_b200 import _lib
    from coregistrationgame_b200.dist import PACK_WORDS
    rec = np.zeros((n_plots, PACK_WORDS), dtype=np.int64)
    rec[:, 0] = keys
    rows = np.zeros(n_plots, dtype=_lib.HYP_RESULT_DTYPE)
    rows["m00"] = 1.0 + rank
    rows["cx"] = (keys.astype(np.uint64) & np.uint64(0xFFFFFFFF)).astype(np.float64)     # the hypothesis id
    rows["k"] = 100 + rank
    rows["passes"] = 7
    rec[:, 1:11] = rows.view(np.int64).reshape(n_plots, 10)
    rec[:, 11] = passes
    return rec


def _worker(rank, world, port, n_plots, n_hyp, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from coregistrationgame_b200.dist import gather_packed, select_best, shard_of
    rng = np.random.default_rng(123)                      # same table on every rank
    score = rng.uniform(0.1, 5.0, (n_plots, n_hyp))
    score[1, :] = np.inf                                  # a plot where every hypothesis is disqualified
    score[2, 5] = score[2, 9] = 0.05                      # exact score tie across ranks -> lowest id wins
    begin, stride = shard_of(rank, world)
    mine = np.arange(begin, n_hyp, stride)
    keys = np.empty(n_plots, dtype=np.int64)
    for p in range(n_plots):
        ks = np.array([orc.pack_best_key(score[p, h], h) for h in mine], dtype=np.uint64)
        keys[p] = np.int64(ks.min())
    g = gather_packed(torch.from_numpy(_pack(keys, rank, n_plots, 1000 + rank)))       # the ONE collective of the path
    gk, rows, passes, _b = select_best(g.numpy())
    q.put((rank, gk.copy(), rows.copy(), passes.copy(), score))
    dist.barrier()
    dist.destroy_process_group()


def test_exchange_best_two_ranks_gloo():
    world, n_plots, n_hyp = 2, 4, 16
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_plots, n_hyp, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    outs.sort(key=lambda t: t[0])
    score = outs[0][4]
    want = np.array([min(int(orc.pack_best_key(score[p, h], h)) for h in range(n_hyp)) for p in range(n_plots)], dtype=np.uint64)
    for rank, gk, rows, passes, _ in outs:
        np.testing.assert_array_equal(gk, want)                     # identical winner on every rank
        win = (gk & np.uint64(0xFFFFFFFF)).astype(int)
        np.testing.assert_array_equal(rows["cx"], win)              # the row comes from the owner of the winner
        np.testing.assert_array_equal(rows["m00"], 1.0 + win % world)
        np.testing.assert_array_equal(rows["k"], 100 + win % world)
        np.testing.assert_array_equal(passes, [1000, 1001])         # pass counts piggy-back on the same collective
    assert (outs[0][1][2] & np.uint64(0xFFFFFFFF)) == 5             # tie -> lowest id
    assert (outs[0][1][1] & np.uint64(0xFFFFFFFF)) == 0             # all disqualified -> id 0, score inf


def _gather_worker(rank, world, port, n_plots, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from coregistrationgame_b200.dist import PACK_WORDS, gather_packed, plot_shard, select_best
    mine = plot_shard(n_plots, rank, world)
    n_rows = (n_plots + world - 1) // world
    packed = np.zeros((n_rows, PACK_WORDS), dtype=np.int64)
    packed[:len(mine)] = _pack(np.array([orc.pack_best_key(0.5 + p, 0) for p in mine], dtype=np.uint64).astype(np.int64), rank, len(mine), 50 + rank)
    gk, rows, passes, _b = select_best(gather_packed(torch.from_numpy(packed)).numpy(), by_plots=True, n_plots=n_plots)
    q.put((rank, gk.copy(), rows.copy(), passes.copy()))
    dist.barrier()
    dist.destroy_process_group()


def test_plot_sharding_gather_two_ranks_gloo():
    world, n_plots = 2, 7
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gather_worker, args=(r, world, port, n_plots, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = np.array([orc.pack_best_key(0.5 + p, 0) for p in range(n_plots)], dtype=np.uint64)
    for _, gk, rows, passes in outs:
        np.testing.assert_array_equal(gk, want)                     # every plot exactly once, from its owner
        np.testing.assert_array_equal(rows["m00"], 1.0 + np.arange(n_plots) % world)
        np.testing.assert_array_equal(passes, [50, 51])


def test_shards_partition_the_hypotheses():
    from coregistrationgame_b200.dist import plot_shard, shard_of
    for world in (1, 2, 4, 8):
        allp = np.concatenate([plot_shard(10001, r, world) for r in range(world)])
        assert sorted(allp.tolist()) == list(range(10001))
        seen = np.concatenate([np.arange(*((b, 4096, s))) for b, s in (shard_of(r, world) for r in range(world))])
        assert sorted(seen.tolist()) == list(range(4096))


def test_single_process_passthrough():
    from coregistrationgame_b200.dist import gather_packed, select_best
    keys = np.array([orc.pack_best_key(0.25, 3), orc.pack_best_key(1.5, 0)], dtype=np.uint64).astype(np.int64)
    g = gather_packed(torch.from_numpy(_pack(keys, 0, 2, 9)))
    from coregistrationgame_b200.dist import PACK_WORDS
    assert tuple(g.shape) == (1, 2, PACK_WORDS)
    gk, rows, passes, _b = select_best(g.numpy())
    np.testing.assert_array_equal(gk.astype(np.int64), keys)
    np.testing.assert_array_equal(rows["cx"], [3, 0])
    assert passes.tolist() == [9]


def test_shard_plan_picks_the_balanced_axis():
    """dist.shard_plan: whole plots per rank for a batch of stands, hypotheses for ONE stand, plots when there are fewer
    start poses than ranks; the same decision on every rank (pure function of sizes)."""
    from coregistrationgame_b200.dist import shard_plan
    assert shard_plan([500] * 128, 4096, 8) == "plots"          # the bench's weak-scaling batch (16 stands per GPU)
    assert shard_plan([500] * 16, 4096, 1) == "plots"           # one rank: everything is its own
    assert shard_plan([500], 4096, 8) == "hypotheses"           # the literal config 3: one stand, strong scaling
    assert shard_plan([500] * 5, 64, 2) == "hypotheses"         # 3 + 2 plots would be 1.2x unbalanced, 32 + 32 poses are not
    assert shard_plan([150] * 10000, 1, 8) == "plots"           # config 4: one start pose per plot
    assert shard_plan([150] * 3, 1, 8) == "plots"               # nothing else to cut
    assert shard_plan([1000, 10, 10, 10], 64, 2) == "hypotheses"   # one big stand among small ones: plots would be 1010 : 20
    assert shard_plan([250] * 8, 7, 4) == "plots"               # 7 poses over 4 ranks = 8/7 unbalanced; plots are even
