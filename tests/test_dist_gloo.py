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


def _worker(rank, world, port, n_plots, n_hyp, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from coregistrationgame_b200.dist import reduce_best, shard_of
    rng = np.random.default_rng(123)                      # same table on every rank
    score = rng.uniform(0.1, 5.0, (n_plots, n_hyp))
    score[1, :] = np.inf                                  # a plot where every hypothesis is disqualified
    score[2, 5] = score[2, 9] = 0.05                      # exact score tie across ranks -> lowest id wins
    begin, stride = shard_of(rank, world)
    mine = np.arange(begin, n_hyp, stride)
    keys = np.empty(n_plots, dtype=np.int64)
    detail = np.zeros((n_plots, 3))
    for p in range(n_plots):
        ks = np.array([orc.pack_best_key(score[p, h], h) for h in mine], dtype=np.uint64)
        j = int(np.argmin(ks))
        keys[p] = np.int64(ks[j])
        detail[p] = [mine[j], score[p, mine[j]], rank]
    gk, gd = reduce_best(torch.from_numpy(keys), torch.from_numpy(detail))
    q.put((rank, gk.numpy().copy(), gd.numpy().copy(), score))
    dist.barrier()
    dist.destroy_process_group()


def test_reduce_best_two_ranks_gloo():
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
    score = outs[0][3]
    want = np.array([min(int(orc.pack_best_key(score[p, h], h)) for h in range(n_hyp)) for p in range(n_plots)], dtype=np.int64)
    for rank, gk, gd, _ in outs:
        np.testing.assert_array_equal(gk, want)                     # identical winner on every rank
        win = (gk.astype(np.uint64) & np.uint64(0xFFFFFFFF)).astype(int)
        np.testing.assert_array_equal(gd[:, 0], win)                # payload comes from the owner of the winner
        np.testing.assert_array_equal(gd[:, 2], win % world)
    assert (outs[0][1].astype(np.uint64)[2] & np.uint64(0xFFFFFFFF)) == 5   # tie -> lowest id
    assert (outs[0][1].astype(np.uint64)[1] & np.uint64(0xFFFFFFFF)) == 0   # all disqualified -> id 0, score inf


def _gather_worker(rank, world, port, n_plots, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from coregistrationgame_b200.dist import gather_plot_shards, plot_shard
    mine = plot_shard(n_plots, rank, world)
    local = np.stack([np.array([p, 10.0 * p + 0.5, rank], dtype=np.float64) for p in mine]) if len(mine) else np.zeros((0, 3))
    full = gather_plot_shards(torch.from_numpy(local), mine, n_plots)
    q.put((rank, full.numpy().copy()))
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
    want = np.stack([np.array([p, 10.0 * p + 0.5, p % world]) for p in range(n_plots)])
    for _, full in outs:
        np.testing.assert_array_equal(full, want)          # every plot exactly once, from its owner


def test_shards_partition_the_hypotheses():
    from coregistrationgame_b200.dist import plot_shard, shard_of
    for world in (1, 2, 4, 8):
        allp = np.concatenate([plot_shard(10001, r, world) for r in range(world)])
        assert sorted(allp.tolist()) == list(range(10001))
        seen = np.concatenate([np.arange(*((b, 4096, s))) for b, s in (shard_of(r, world) for r in range(world))])
        assert sorted(seen.tolist()) == list(range(4096))


def test_reduce_best_single_process_passthrough():
    from coregistrationgame_b200.dist import reduce_best
    k = torch.tensor([5, 3], dtype=torch.int64)
    d = torch.ones(2, 4, dtype=torch.float64)
    gk, gd = reduce_best(k, d)
    assert torch.equal(gk, k) and torch.equal(gd, d)
