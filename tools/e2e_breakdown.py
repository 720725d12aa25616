"""Where an end-to-end step of register_batch_distributed goes (wall clock per rank, synchronised around each part):
target upload + grid build, batch upload, kernel + exchange + read-back, teardown.  Run under torchrun (any N) or alone.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 tools/e2e_breakdown.py"""
import json, os, sys, time
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coregistrationgame_b200 import IcpBatch, TargetIndex, _lib, synthetic as syn
from coregistrationgame_b200.batch import hypothesis_table, translation_lattice
from coregistrationgame_b200.dist import PACK_WORDS, exchange_best, shard_of

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
_lib.check(_lib.load().ficp_set_device(local))
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
n_plots = 16 * world
tgt, plots, _ = syn.synthetic_scene(1_000_000, 500, seed=3, dims=3, n_plots=n_plots, hidden_pose=True)
hyp = hypothesis_table(128, flips=(0, 1), translations=translation_lattice(4, 2.5))
def pinned(a):
    t = torch.empty(a.shape, dtype=torch.float64, pin_memory=True); t.numpy()[...] = a; return t.numpy()
h_tgt, h_plots, h_hyp = pinned(tgt), [pinned(p) for p in plots], pinned(hyp)
stream = torch.cuda.current_stream()
group = dist.group.WORLD if world > 1 else None
acc = np.zeros(5)
reps = 6
for it in range(reps + 2):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    index = TargetIndex(h_tgt)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    batch = IcpBatch(index, h_plots, h_hyp, hyp_shard=shard_of(rank, world))
    torch.cuda.synchronize(); t2 = time.perf_counter()
    packed = torch.empty((n_plots, PACK_WORDS), dtype=torch.int64, device=dev)
    batch.run(stream)
    torch.cuda.synchronize(); t3 = time.perf_counter()
    g = exchange_best(batch, packed, group, stream).cpu().numpy()
    t4 = time.perf_counter()
    batch.close(); index.close()
    torch.cuda.synchronize(); t5 = time.perf_counter()
    if it >= 2:
        acc += [t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4]
acc = acc / reps * 1e3
t = torch.tensor(acc, dtype=torch.float64, device=dev)
if world > 1:
    allt = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(allt, t)
else:
    allt = [t]
if rank == 0:
    names = ["target upload + grid build", "batch create (host prep + upload)", "kernel", "pack + all_gather + read-back", "destroy"]
    rows = np.array([a.cpu().numpy() for a in allt])
    print(json.dumps({"world": world, "plots": n_plots, "build_ms_device": index.info()["build_ms"] if False else None,
                      "ms_mean_over_ranks": dict(zip(names, rows.mean(0).round(3).tolist())),
                      "ms_max_over_ranks": dict(zip(names, rows.max(0).round(3).tolist()))}))

# ---- part 2: the public call itself, as bench.py's e2e leg drives it (no synchronisation between the parts), with wall-clock
# timers around its pieces: where does a step of register_batch_distributed go when the ranks run free?
if os.environ.get("FICP_E2E_PART2", "1") == "1":
    import coregistrationgame_b200.dist as D
    import coregistrationgame_b200.batch as B
    if world > 1:
        pass
    T = {"index_create": 0.0, "index_close": 0.0, "batch_create": 0.0, "batch_close": 0.0, "run_enqueue": 0.0, "exchange+cpu": 0.0}
    def wrap(cls, name, key):
        orig = getattr(cls, name)
        def f(*a, **k):
            t0 = time.perf_counter(); r = orig(*a, **k); T[key] += time.perf_counter() - t0; return r
        setattr(cls, name, f)
    wrap(B.TargetIndex, "__init__", "index_create"); wrap(B.TargetIndex, "close", "index_close")
    wrap(B.IcpBatch, "__init__", "batch_create"); wrap(B.IcpBatch, "close", "batch_close"); wrap(B.IcpBatch, "run", "run_enqueue")
    orig_ex = D.exchange_best
    def ex(*a, **k):
        t0 = time.perf_counter(); r = orig_ex(*a, **k).cpu(); T["exchange+cpu"] += time.perf_counter() - t0; return r
    D.exchange_best = ex
    for mode in ("free", "sync_each_step"):
        for k in T: T[k] = 0.0
        D.register_batch_distributed(h_plots, h_tgt, h_hyp) if world > 1 else B.register_batch(h_plots, h_tgt, h_hyp, per_hypothesis=False)
        torch.cuda.synchronize()
        if world > 1: dist.barrier()
        for k in T: T[k] = 0.0
        t0 = time.perf_counter()
        for _ in range(5):
            D.register_batch_distributed(h_plots, h_tgt, h_hyp) if world > 1 else B.register_batch(h_plots, h_tgt, h_hyp, per_hypothesis=False)
            if mode == "sync_each_step":
                torch.cuda.synchronize()
        torch.cuda.synchronize()
        tot = (time.perf_counter() - t0) / 5 * 1e3
        v = torch.tensor([tot] + [T[k] / 5 * 1e3 for k in T], dtype=torch.float64, device=dev)
        if world > 1:
            allv = [torch.empty_like(v) for _ in range(world)]; dist.all_gather(allv, v)
        else:
            allv = [v]
        if rank == 0:
            rows = np.array([a.cpu().numpy() for a in allv])
            print(json.dumps({"world": world, "mode": mode, "ms_per_step_max": round(float(rows[:, 0].max()), 3),
                              "mean_over_ranks": dict(zip(["step"] + list(T), rows.mean(0).round(3).tolist())),
                              "max_over_ranks": dict(zip(["step"] + list(T), rows.max(0).round(3).tolist()))}))
    if world > 1:
        dist.destroy_process_group()
