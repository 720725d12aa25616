#!/usr/bin/env python
"""Benchmark of the B200 Fractional-ICP hot path (BASELINE.json metric: hypothesis-iterations/s).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host CPU

Workload (config 3 of BASELINE.json, "synthetic stand"): 500 trees vs 1e6 CHM points, 4096 start-pose
hypotheses (128 rotations x 2 flips x 4x4 translations at 2.5 m), XYZ matching; `--plots-per-gpu` such
plots per GPU (weak scaling: the hypotheses of every plot are dealt round-robin to the ranks, so each
GPU always runs plots_per_gpu * 4096 ICPs; the exchange step is the NCCL all-reduce(MIN) of the packed
best keys).  A "step" = every one of those ICPs run to convergence (two stages) against the resident
target grid + the best-of-hypotheses reduction.  One unit = one hypothesis-iteration = one NN pass of
one hypothesis (`find_correspondences` call in the reference, ficp.py:123,137).

value   device-timed (CUDA events on the launching stream), inputs resident in HBM.
e2e     the same metric through the public host-array API (`register_batch[_distributed]`): target +
        plots + hypotheses copied from pinned host memory, grid build, kernel, results read back, all
        inside the timed region (wall clock around synchronised calls).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ALG_BYTES_PER_QUERY = 312.0      # 18 candidate records x 16 B + 3 row-range lookups x 8 B
GRID_BYTES_PER_POINT = 48.0      # grid build traffic per target point


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--points", type=int, default=1_000_000)
    ap.add_argument("--trees", type=int, default=500)
    ap.add_argument("--dims", type=int, default=3, choices=[2, 3])
    ap.add_argument("--plots-per-gpu", type=int, default=16)
    ap.add_argument("--team", type=int, default=0, help="warps per ICP at launch (0 = planner's choice)")
    ap.add_argument("--helpers", choices=["auto", "on", "off"], default="auto",
                    help="elastic kernel (idle warps help the ICPs in flight): planner's choice, forced on, forced off")
    ap.add_argument("--cta", choices=["auto", "on", "off"], default="auto",
                    help="kernel shape: CTA-per-ICP (latency shape, small batches) vs warp-per-ICP; auto = the library's planner")
    ap.add_argument("--fixed-frac", type=float, default=0.0, help="fixed trim fraction (0 = FRMSD-optimal, the reference's behaviour)")
    ap.add_argument("--no-single-stand", action="store_true", help="skip the one-stand strong-scaling leg")
    ap.add_argument("--rotations", type=int, default=128)
    ap.add_argument("--tside", type=int, default=4, help="translation lattice side (tside^2 translations)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="hypotheses in the CPU baseline sample (0 = one per core, <= 32)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = same as --steps")
    ap.add_argument("--warps", type=int, default=0)
    ap.add_argument("--ctas-per-sm", type=int, default=0)
    ap.add_argument("--window-margin", type=float, default=-1.0)
    ap.add_argument("--pts-per-cell", type=float, default=0.0, help="0 = library default")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--workload", default="c3", choices=["c3", "c2", "c4", "c5"],
                    help="c3 (default, the metric's config) | c2: 200 trees vs 1e5 points, 1024 hypotheses | "
                         "c4: 1250 plots/GPU x 150 trees vs 1e7 points, one start pose per plot | "
                         "c5: c3 shape on the adversarial scene (30 %% outlier trees, 30 %% omissions, duplicated + lattice-tied "
                         "CHM points) with the trim-fraction sweep 0.5-0.95 reported beside the FRMSD-optimal mode")
    return ap.parse_args()


def workload(args, n_plots):
    from coregistrationgame_b200 import synthetic as orc   # input generator (the oracle is imported by the CPU legs only)
    from coregistrationgame_b200.batch import hypothesis_matrix, hypothesis_table, translation_lattice
    if args.workload == "c4":
        tgt, plots, _ = orc.synthetic_scene(args.points, args.trees, seed=4, dims=args.dims, n_plots=n_plots,
                                            hidden_pose=False)
        rng = np.random.default_rng(1)
        plots = [orc.apply_pose(p, np.r_[hypothesis_matrix(rng.uniform(-5.0, 5.0), 0).ravel(), rng.uniform(-1.5, 1.5, 2)],
                                   p[:, :2].mean(axis=0)) for p in plots]
        hyp = np.array([[1.0, 0.0, 0.0, 1.0, 0.0, 0.0]])
        name = (f"C4 batched breakout: {n_plots} plots x {args.trees} trees vs a shared {args.points}-point CHM, one ICP per plot, "
                f"{'XYZ' if args.dims == 3 else 'XY'} matching")
        return tgt, plots, hyp, name
    if args.workload == "c5":
        tgt, plots, _ = orc.synthetic_scene(args.points, args.trees, seed=5, dims=args.dims, n_plots=n_plots, hidden_pose=True,
                                            out_frac=0.3, omit_frac=0.3, dup_every=10, lattice_patch=8)
    else:
        tgt, plots, _ = orc.synthetic_scene(args.points, args.trees, seed=3, dims=args.dims, n_plots=n_plots,
                                            hidden_pose=True)
    hyp = hypothesis_table(args.rotations, flips=(0, 1), translations=translation_lattice(args.tside, 2.5))
    tag = {"c3": "C3 synthetic stand", "c2": "C2 synthetic plot", "c5": "C5 adversarial stand (30 % outliers, omissions, duplicated/tied CHM points)"}[args.workload]
    name = (f"{tag}: {args.trees} trees vs {args.points} CHM points, {hyp.shape[0]} hypotheses "
            f"({args.rotations} rot x 2 flips x {args.tside}x{args.tside} translations), {'XYZ' if args.dims == 3 else 'XY'} matching")
    return tgt, plots, hyp, name


def apply_workload_defaults(args):
    """BASELINE.json configs other than the metric's own: only fills values the user left at the c3 defaults."""
    if args.workload == "c2":
        if args.points == 1_000_000: args.points = 100_000
        if args.trees == 500: args.trees = 200
        if args.rotations == 128: args.rotations = 512
        if args.tside == 4: args.tside = 1
    elif args.workload == "c4":
        if args.points == 1_000_000: args.points = 10_000_000
        if args.trees == 500: args.trees = 150
        if args.plots_per_gpu == 16: args.plots_per_gpu = 1250


# ----------------------------------------------------------------------------------- CPU reference arm
_REF_CLS = None


def reference_class():
    """The UNMODIFIED reference class, loaded from oracle/_ref/ficp.py (a byte copy of /root/reference/ficp.py staged by
    tools/vendor_ref.sh; the reference tree itself does not exist on the GPU box).  None when it has not been staged."""
    global _REF_CLS
    if _REF_CLS is None:
        path = os.path.join(ROOT, "oracle", "_ref", "ficp.py")
        if not os.path.exists(path):
            _REF_CLS = False
        else:
            import importlib.util
            spec = importlib.util.spec_from_file_location("ficp_reference_unmodified", path)
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            _REF_CLS = mod.FractionalICP
    return _REF_CLS or None


class _Budget(Exception):
    pass


def _cpu_one(job):
    """One hypothesis on the CPU.  hoist=False: `FractionalICP(pre_transform(src, h), tgt).run()` of the unmodified
    reference (kd-tree rebuilt on every pass, O(N^2) FRMSD loop - ficp.py:122-154) under a subclass that only counts
    `find_correspondences` calls (= hypothesis-iterations) and stops between passes once the step's time budget is
    spent; the passes completed so far are the units of work done.  hoist=True: the oracle port with the kd-tree built
    once and a cumsum FRMSD (the "fair" CPU baseline)."""
    from oracle import ficp_oracle as orc
    src, tgt, row, centre, hoist, deadline = job
    s0 = orc.pre_transform(src, row, centre)
    ref = reference_class()
    if not hoist and ref is not None:
        class Counting(ref):
            n_passes = 0

            def find_correspondences(self, source, target):
                if time.perf_counter() > deadline:
                    raise _Budget()
                out = super().find_correspondences(source, target)
                self.n_passes += 1
                return out
        icp = Counting(s0, tgt)
        try:
            icp.run()
        except _Budget:
            pass
        return icp.n_passes
    tr = orc.RunTrace(deadline=deadline, light=True)
    try:
        if hoist:
            orc.ficp_run(s0, tgt, nn="tree", hoist_tree=True, trace=tr, closed_form=True)
        else:
            orc.ficp_run(s0, tgt, nn="reference", hoist_tree=False, trace=tr, pairwise=True)
    except orc.TimeBudgetExceeded:
        pass
    return tr.passes


_POOL = None


def cpu_step(src, tgt, hyp, n_sample, procs, hoist=False, budget_s=8.0):
    """Runs `n_sample` strided hypotheses of one plot on `procs` processes for at most ~`budget_s` seconds of wall
    time (a pass in flight is finished); returns (passes completed, seconds)."""
    global _POOL
    import multiprocessing as mp
    if _POOL is None or _POOL[1] != procs:
        if _POOL is not None:
            _POOL[0].terminate()
        _POOL = (mp.get_context("fork").Pool(procs), procs)
    centre = src[:, :2].mean(axis=0)
    ids = np.linspace(0, hyp.shape[0] - 1, n_sample).astype(int)
    t0 = time.perf_counter()
    jobs = [(src, tgt, hyp[h], centre, hoist, t0 + budget_s) for h in ids]
    passes = _POOL[0].map(_cpu_one, jobs, chunksize=1)
    return int(sum(passes)), time.perf_counter() - t0


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    tgt, plots, hyp, name = workload(args, 1)
    procs = max(1, min(host_threads(), 32))
    n_sample = args.cpu_sample or procs
    for _ in range(args.warmup):
        cpu_step(plots[0], tgt, hyp, min(n_sample, procs), procs, budget_s=1.0)
    tot_p, tot_t = 0, 0.0
    for _ in range(args.steps):
        p, t = cpu_step(plots[0], tgt, hyp, n_sample, procs)
        tot_p += p
        tot_t += t
    val = tot_p / tot_t
    kind = "reference" if reference_class() is not None else "port"
    sample = (f"{n_sample} strided hypotheses of one plot per step on {procs} processes, ~8 s time budget per step; "
              + ("unmodified reference ficp.py (oracle/_ref, staged by tools/vendor_ref.sh): FractionalICP(pre_transform(src, h), tgt).run()"
                 if kind == "reference" else "oracle port of the reference as shipped (oracle/_ref not staged)"))
    line = {"impl": "reference", "metric": "FICP hypothesis-iterations/sec", "value": val, "unit": "hyp-iter/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": name, "sample": sample},
            "nn_queries_per_s": val * args.trees,
            "cpu_baseline": {"value": val, "unit": "hyp-iter/s", "cores": procs, "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": "hyp-iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ----------------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz = index, False, [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {getattr(nv, n): n.replace("nvmlClocksEventReason", "").replace("nvmlClocksThrottleReason", "")
                 for n in dir(nv) if n.startswith("nvmlClocksThrottleReason") or n.startswith("nvmlClocksEventReason")}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if isinstance(bit, int) and bit and (r & bit) and bin(bit).count("1") == 1:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.02)

    def summary(self):
        s = sorted(self.samples)
        keep = [r for r in self.reasons if r not in ("GpuIdle", "None", "ApplicationsClocksSetting")]
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(keep),
                "samples": len(s)}


# ----------------------------------------------------------------------------------- CUDA arm
def kernel_source_hash():
    """sha256 (16 hex) of the sources the warp-per-ICP persistent kernel is compiled from: ncu-derived figures in
    profiles/traffic.json (all of that kernel) are only used when they were captured on exactly this code.  (icp_team.cu, the
    CTA-per-ICP kernel, is a separate translation unit and does not enter the default bench line's dominant kernel.)"""
    import hashlib
    h = hashlib.sha256()
    for f in ("icp_persistent.cu", "icp_shared.cuh", "nn_search.cuh", "ficp_common.cuh"):
        h.update(open(os.path.join(ROOT, "coregistrationgame_b200", "csrc", f), "rb").read())
    return h.hexdigest()[:16]


def run_b200(args):
    import ctypes as C
    import torch
    import torch.distributed as dist
    from coregistrationgame_b200 import IcpBatch, TargetIndex, _lib, register_batch
    from coregistrationgame_b200.dist import PACK_WORDS, exchange_best, register_batch_distributed, shard_of, shard_plan

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    _lib.require_device()
    _lib.check(_lib.load().ficp_set_device(local))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    n_plots = args.plots_per_gpu * world
    tgt, plots, hyp, name = workload(args, n_plots)
    one_pose = hyp.shape[0] < world or args.workload == "c4"     # one start pose per plot (config 4)
    # the same cut register_batch_distributed takes (dist.shard_plan): whole plots per rank when that balances
    by_plots = one_pose or (world > 1 and shard_plan([len(p) for p in plots], hyp.shape[0], world) == "plots")
    props = _lib.device_props()
    cta = {"auto": None, "on": True, "off": False}[args.cta]
    helpers = {'auto': None, 'on': True, 'off': False}[args.helpers]
    bkw = dict(warps_per_cta=args.warps, ctas_per_sm=args.ctas_per_sm, window_margin=args.window_margin, team_warps=args.team,
               helpers=helpers, cta_per_icp=cta)
    if args.fixed_frac > 0:
        bkw["fixed_frac"] = args.fixed_frac

    # ---- resident inputs
    index = TargetIndex(tgt, pts_per_cell=(args.pts_per_cell or None))
    tinfo = index.info()
    if by_plots:
        mine = list(range(rank, n_plots, world))
        batch = IcpBatch(index, [plots[p] for p in mine], hyp, **({"min_k": 0} if one_pose else {}), **bkw)
        n_rows = (n_plots + world - 1) // world
    else:
        batch = IcpBatch(index, plots, hyp, hyp_shard=shard_of(rank, world), **bkw)
        n_rows = n_plots
    packed = torch.zeros((n_rows, PACK_WORDS), dtype=torch.int64, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2
    stream = torch.cuda.current_stream()
    group = dist.group.WORLD if world > 1 else None

    def step(evk=None):
        """One step on resident inputs = what register_batch_distributed enqueues after the upload: the persistent kernel
        and THE exchange (device-side pack + one all_gather; dist.exchange_best, the shipped function)."""
        batch.run(stream)
        if evk is not None:
            evk.record(stream)
        return exchange_best(batch, packed, group, stream)

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    passes_per_step = batch.results(stream, per_hypothesis=False)["stats"]["passes"]

    sampler = ClockSampler(local)
    sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    evs = []
    for _ in range(args.steps):
        flush.fill_(1)                           # evict L2 between timed iterations
        e0, ek, e1 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record(stream)
        step(ek)
        e1.record(stream)
        evs.append((e0, ek, e1))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler.stop_flag = True
    step_ms = sum(a.elapsed_time(c) for a, _, c in evs)
    kern_ms = sum(a.elapsed_time(b) for a, b, _ in evs) / args.steps
    stats = batch.results(stream, per_hypothesis=False)["stats"]
    t = torch.tensor([step_ms], dtype=torch.float64, device=dev)
    units = torch.tensor([float(stats["passes"])], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(units, op=dist.ReduceOp.SUM)
    total_ms = float(t.item())
    passes_all = float(units.item())             # all ranks, one step
    value = passes_all * args.steps / (total_ms * 1e-3)

    # ---- end to end through the public API, host (pinned) buffers in, results out
    def pinned(a):
        tbuf = torch.empty(a.shape, dtype=torch.float64, pin_memory=True)
        tbuf.numpy()[...] = a
        return tbuf.numpy()
    h_tgt, h_plots, h_hyp = pinned(tgt), [pinned(p) for p in plots], pinned(hyp)
    e2e_steps = args.e2e_steps or args.steps

    ekw = dict(warps_per_cta=args.warps, ctas_per_sm=args.ctas_per_sm, cta_per_icp=cta)
    if args.fixed_frac > 0:
        ekw["fixed_frac"] = args.fixed_frac
    if one_pose:
        ekw["min_k"] = 0

    def e2e_step():
        if world > 1:
            return register_batch_distributed(h_plots, h_tgt, h_hyp, **ekw)
        r = register_batch(h_plots, h_tgt, h_hyp, per_hypothesis=False, **ekw)
        r["passes_global"] = r["stats"]["passes"]
        return r
    if args.no_e2e:
        e2e_steps = 0
    r = {"h2d_bytes": 0, "d2h_bytes": 0}
    for _ in range(2 if e2e_steps else 0):      # untimed: the device memory pool reaches its steady size
        r = e2e_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    # like timeit: no cyclic garbage collection inside the timed loops (with torch imported a full collection walks millions of
    # objects); collected once before instead.  (The 14-40 ms step that showed up once in ~25 end-to-end steps was not this but
    # the memory pool growing when a handle's buffers were freed on another stream than they were allocated on - fixed in the
    # library, ficp_internal.h release_stream; tools/e2e_outlier_probe.py: 118 steps, max 48.6 ms.)
    import gc
    gc.collect()
    gc.disable()
    t0 = time.perf_counter()
    e2e_passes = 0
    e2e_ms_list = []
    for _ in range(e2e_steps):
        ts = time.perf_counter()
        r = e2e_step()                      # synchronous: the winners are on the host when it returns
        e2e_ms_list.append((time.perf_counter() - ts) * 1e3)
        e2e_passes += r["passes_global"]
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_val = e2e_passes / float(e2e_s.item()) if e2e_steps else None
    h2d, d2h = r["h2d_bytes"], r["d2h_bytes"]

    # the same call with the target index already resident (a caller registering many batches against one CHM layer keeps
    # its TargetIndex): plots + hypotheses up, winners back - no target upload, no grid build
    e2e_res = None
    if e2e_steps:
        def e2e_step_resident():
            if world > 1:
                return register_batch_distributed(h_plots, None, h_hyp, index=index, **ekw)
            rr = register_batch(h_plots, None, h_hyp, index=index, per_hypothesis=False, **ekw)
            rr["passes_global"] = rr["stats"]["passes"]
            return rr
        rr = e2e_step_resident()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        rp = 0
        for _ in range(e2e_steps):
            rr = e2e_step_resident()
            rp += rr["passes_global"]
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        rs = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(rs, op=dist.ReduceOp.MAX)
        e2e_res = {"value": rp / float(rs.item()), "unit": "hyp-iter/s", "h2d_bytes_per_step": int(rr["h2d_bytes"]),
                   "d2h_bytes_per_step": int(rr["d2h_bytes"]), "what": "register_batch(..., index=resident TargetIndex)"}
        if one_pose and world == 1:
            # thousands of plots: the per-plot Python work of a LIST of arrays bounds the step; the stacked input form
            # register_batch((rows, offsets), ...) leaves one pass over one array
            st_rows = pinned(np.vstack(h_plots))
            st_offs = np.concatenate([[0], np.cumsum([len(p) for p in h_plots])]).astype(np.int64)
            register_batch((st_rows, st_offs), None, h_hyp, index=index, per_hypothesis=False, **ekw)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            sp = 0
            for _ in range(e2e_steps):
                sp += register_batch((st_rows, st_offs), None, h_hyp, index=index, per_hypothesis=False, **ekw)["stats"]["passes"]
            torch.cuda.synchronize()
            e2e_res["stacked_input"] = {"value": sp / (time.perf_counter() - t0), "unit": "hyp-iter/s",
                                        "what": "register_batch((rows, offsets), index=resident TargetIndex): plots already stacked"}

    gc.enable()

    # ---- the named config taken literally: ONE stand x all its hypotheses, strong-scaled over the ranks
    # (time to register a single stand; the headline above is the throughput of a batch of stands)
    def time_single(shard, with_exchange):
        sb = IcpBatch(index, [plots[0]], hyp, hyp_shard=shard, team_warps=args.team, helpers=helpers, cta_per_icp=cta,
                      **({"fixed_frac": args.fixed_frac} if args.fixed_frac > 0 else {}))
        spk = torch.zeros((1, PACK_WORDS), dtype=torch.int64, device=dev)

        def sstep():
            sb.run(stream)
            if with_exchange:
                exchange_best(sb, spk, group, stream)
        for _ in range(3):
            sstep()
        torch.cuda.synchronize()
        if with_exchange and world > 1:
            dist.barrier()
        sev = []
        for _ in range(max(args.steps, 5)):
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            sstep()
            b.record(stream)
            sev.append((a, b))
        torch.cuda.synchronize()
        ms = sorted(a.elapsed_time(b) for a, b in sev)
        res = (ms[len(ms) // 2], float(sb.results(stream, per_hypothesis=False)["stats"]["passes"]), dict(sb.info))
        sb.close()
        return res

    single = None
    if not one_pose and not args.no_single_stand:
        ms_n, p_n, info_n = time_single(shard_of(rank, world), True)
        sms_ = torch.tensor([ms_n], dtype=torch.float64, device=dev)
        sp_ = torch.tensor([p_n], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(sms_, op=dist.ReduceOp.MAX)
            dist.all_reduce(sp_, op=dist.ReduceOp.SUM)
        single = {"workload": "one stand x %d hypotheses (plot 0), hypotheses sharded over %d GPU(s)" % (hyp.shape[0], world),
                  "scaling": "strong", "ms": float(sms_.item()), "hyp_iterations": float(sp_.item()),
                  "hyp_iter_per_s": float(sp_.item()) / (float(sms_.item()) * 1e-3), "team_warps": info_n["team_warps"],
                  "cta_per_icp": info_n["cta_per_icp"], "warps_per_cta": info_n["warps_per_cta"], "ctas": info_n["ctas"],
                  "timing": "median of %d launches (kernel + exchange), L2 flushed before each" % max(args.steps, 5)}
        if world > 1:
            # the same stand on ONE GPU of this box, in this run: the denominator of the strong-scaling figure
            if rank == 0:
                ms_1, _, info_1 = time_single((0, 1), False)
                single["ms_1gpu_same_run"] = ms_1
                single["speedup_vs_1gpu"] = ms_1 / single["ms"]
                single["kernel_1gpu"] = "cta_per_icp" if info_1["cta_per_icp"] else "warp_per_icp"
            dist.barrier()

    # ---- C5: trim-fraction sweep beside the FRMSD-optimal mode (device-timed, resident inputs, this rank's shard)
    sweep = None
    if args.workload == "c5" and rank == 0:
        sweep = {}
        for frac in (None, 0.5, 0.6, 0.7, 0.8, 0.9, 0.95):
            sb = IcpBatch(index, plots[:min(len(plots), 4)], hyp, cta_per_icp=cta, **({"fixed_frac": frac} if frac else {}))
            for _ in range(2):
                sb.run(stream)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            flush.fill_(1)
            a.record(stream)
            sb.run(stream)
            b.record(stream)
            torch.cuda.synchronize()
            st = sb.results(stream, per_hypothesis=False)["stats"]
            sweep["auto" if frac is None else str(frac)] = {"ms": a.elapsed_time(b), "hyp_iterations": st["passes"],
                                                             "hyp_iter_per_s": st["passes"] / (a.elapsed_time(b) * 1e-3),
                                                             "searched_share": st["searched_queries"] / max(1, st["queries"])}
            sb.close()

    # ---- standalone kernels (reported, not the headline): bulk NN query and grid build; L2 read peak of this run
    extra = {}
    lib = _lib.load()
    l2 = C.c_double(0.0)
    _lib.check(lib.ficp_measure_l2_read_gbs(32 << 20, 20, C.byref(l2)))
    if rank == 0:
        nq = 1 << 22
        rng = np.random.default_rng(1)
        bb = tinfo["bbox"]
        q = np.empty((nq, args.dims))
        q[:, 0] = rng.uniform(bb[0], bb[1], nq)
        q[:, 1] = rng.uniform(bb[2], bb[3], nq)
        if args.dims == 3:
            q[:, 2] = rng.uniform(5, 35, nq)
        dq = torch.from_numpy(q).to(dev)
        didx = torch.empty(nq, dtype=torch.int32, device=dev)
        ddist = torch.empty(nq, dtype=torch.float64, device=dev)
        sp = C.c_void_p(stream.cuda_stream)
        qindex = TargetIndex(tgt, pts_per_cell=(args.pts_per_cell or None), purpose="query")   # bulk-query grid density

        def time_nn(kernel):
            call = lambda: _lib.check(lib.ficp_nn_query_device_ex(qindex.handle, C.c_void_p(dq.data_ptr()), nq, args.dims, int(args.dims == 3),
                                                                  C.c_void_p(didx.data_ptr()), C.c_void_p(ddist.data_ptr()), kernel, None, sp))
            for _ in range(2):
                call()
            torch.cuda.synchronize()
            times = []
            for _ in range(5):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                flush.fill_(1)
                a.record(stream)
                call()
                b.record(stream)
                torch.cuda.synchronize()
                times.append(a.elapsed_time(b))
            return sorted(times)[len(times) // 2]
        ms_thread = time_nn(1)
        ms_bulk = time_nn(2)
        ms = time_nn(0)                 # the planner's choice (what ficp_nn_query[_device] runs)
        cnt = (C.c_uint64 * 3)()
        _lib.check(lib.ficp_nn_query_device_ex(qindex.handle, C.c_void_p(dq.data_ptr()), nq, args.dims, int(args.dims == 3),
                                               C.c_void_p(didx.data_ptr()), C.c_void_p(ddist.data_ptr()), 2, cnt, sp))
        alg = nq * ALG_BYTES_PER_QUERY / (ms * 1e-3) / 1e9
        frac = lambda t: (nq * ALG_BYTES_PER_QUERY / (t * 1e-3) / 1e9 / l2.value) if l2.value else None
        extra["nn_query_kernel"] = {"queries": nq, "ms": ms, "queries_per_s": nq / (ms * 1e-3),
                                    "kernel": "planner's choice (ficp_nn_query_device); whole call timed, L2 flushed before each, median of 5",
                                    "alg_GBps_L2_level": alg, "l2_read_peak_GBps_measured": l2.value,
                                    "frac_of_l2_peak": alg / l2.value if l2.value else None,
                                    "hbm_compulsory_GBps": nq * (8.0 * args.dims + 12.0) / (ms * 1e-3) / 1e9,
                                    "thread_per_query_kernel": {"ms": ms_thread, "queries_per_s": nq / (ms_thread * 1e-3), "frac_of_l2_peak": frac(ms_thread)},
                                    "bulk_kernel": {"ms": ms_bulk, "queries_per_s": nq / (ms_bulk * 1e-3), "frac_of_l2_peak": frac(ms_bulk),
                                                    "what": "qbin + qscan + qscatter + qplan + nn_bulk_kernel (cell-ordered queries, windows staged by cp.async.bulk) + ring finish",
                                                    "resolved": {"window": int(cnt[0]), "global_grid": int(cnt[1]), "rings": int(cnt[2])}},
                                    "bound": "instruction issue / FP64 pipe, not bandwidth: ~25-55 fp64 candidates per query (DESIGN.md 4.2)",
                                    "cell_m": qindex.info()["cell"]}
        qindex.close()
        # grid build of the bench target: device time of its launch chain (CUDA events inside ficp_target_create), median of
        # five warm builds - the build that made `index` above was the process's first launches (module load, pool growth)
        gb = []
        for _ in range(6):
            t2 = TargetIndex(tgt, pts_per_cell=(args.pts_per_cell or None))
            gb.append(t2.info()["build_ms"])
            t2.close()
        gb_ms = sorted(gb[1:])[2]
        extra["grid_build"] = {"points": int(tinfo["m"]), "ms": gb_ms, "ms_first_build_of_the_process": tinfo["build_ms"],
                               "alg_GBps": tinfo["m"] * GRID_BYTES_PER_POINT / (gb_ms * 1e-3) / 1e9,
                               "timing": "CUDA events around the build's launch chain, median of 5 warm builds",
                               "grid": [tinfo["grid_w"], tinfo["grid_h"]], "cell_m": tinfo["cell"]}

    # ---- roofline of the dominant kernel (the persistent ICP kernel), SURVEY 8(d): the working set (cell-sorted target,
    # 32 MB at 1e6 points) is L2-resident, so the bound is the L2 -> SM read path.  achieved = algorithmic bytes at the L2
    # level (hypothesis-iterations x trees x 312 B: 18 candidate records x 16 B + 3 row-range lookups x 8 B per query) per
    # launch / kernel time; peak = the L2 read bandwidth measured in THIS run by ficp_measure_l2_read_gbs (32 MB buffer
    # swept 20 times by 148 x 8 CTAs); frac = achieved / peak.  The kernel serves most of those bytes from shared memory
    # and skips ~93 % of the searches outright, so frac is an EFFECTIVE figure; the actual traffic (ncu, same workload) is
    # reported beside it, and the real limiter (issue slots) in `on_chip`.
    alg_bytes = passes_per_step * args.trees * ALG_BYTES_PER_QUERY          # this rank, one launch
    achieved = alg_bytes / (kern_ms * 1e-3) / 1e9
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm_peak = json.load(open(peaks_path))["hbm_gbs"] if os.path.exists(peaks_path) else 6650.0
    prof, prof_ok = {}, False
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            prof = json.load(open(tpath))
            prof_ok = (prof.get("kernel_source_sha16") == kernel_source_hash() and args.workload == "c3" and args.dims == 3
                       and args.trees == 500 and args.plots_per_gpu == 16 and not batch.info["cta_per_icp"])
        except Exception:
            prof = {}
    roofline = {"bound": "l2", "achieved": achieved, "peak": l2.value, "unit": "GB/s",
                "frac": (achieved / l2.value) if l2.value else None,
                "traffic": prof.get("icp_kernel_dram_bytes_per_launch") if prof_ok else None,
                "l2_bytes_per_launch": prof.get("icp_kernel_lts_bytes_per_launch") if prof_ok else None,
                "kernel": ("icp_team_kernel (CTA-per-ICP two-stage FICP)" if batch.info["cta_per_icp"] else "icp_kernel (persistent two-stage FICP, warp per ICP)"),
                "kernel_ms": kern_ms,
                "peak_source": "L2 read bandwidth measured in this run (ficp_measure_l2_read_gbs: 32 MB buffer, 20 sweeps, 148 x 8 CTAs x 256 threads, 16 B loads)",
                "hbm_peak_GBps": hbm_peak, "frac_of_hbm_peak": achieved / hbm_peak,
                "profile_commit_matches_head": prof_ok,
                "note": ("algorithmic bytes = hypothesis-iterations x trees x 312 B (SURVEY 8d, L2 level, as if every candidate "
                         "record of every query came from L2 on every pass); the kernel serves them from a shared-memory window and "
                         "proves ~93 % of the queries unchanged without a search, so frac is an effective figure - `traffic` "
                         "(DRAM) and `l2_bytes_per_launch` are what ncu measured on this workload at this source hash "
                         "(null when the kernel sources changed since the capture)")}

    # on-chip view of the same kernel: share of the SM issue slots it uses, from the warp-instructions per
    # hypothesis-iteration ncu counted on this workload at this source hash and the live kernel time
    on_chip = None
    ipp = prof.get("icp_kernel_warp_instructions_per_hyp_iteration") if prof_ok else None
    if ipp:
        issue_peak = float(props["sms"]) * 4.0 * float(props["clock_khz"]) * 1e3   # 4 schedulers/SM, 1 warp-instr/cycle
        issued = float(ipp) * float(passes_per_step) / (kern_ms * 1e-3)
        on_chip = {"bound": "issue slots", "achieved": issued, "peak": issue_peak, "unit": "warp-instr/s",
                   "frac": issued / issue_peak, "source": prof.get("source")}
    roofline["on_chip"] = on_chip

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        procs = max(1, min(host_threads(), 32))
        n_sample = args.cpu_sample or procs
        p_asis, t_asis = cpu_step(plots[0], tgt, hyp, n_sample, procs, hoist=False, budget_s=12.0)
        p_h, t_h = cpu_step(plots[0], tgt, hyp, n_sample, procs, hoist=True, budget_s=12.0)
        sample = f"{n_sample} strided hypotheses of plot 0 on {procs} processes, <= ~12 s time budget"
        # single core (SURVEY 8d asks for both): one hypothesis on one process, ~6 s budget each
        p1, t1 = cpu_step(plots[0], tgt, hyp, 1, 1, hoist=False, budget_s=6.0)
        p1h, t1h = cpu_step(plots[0], tgt, hyp, 1, 1, hoist=True, budget_s=6.0)
        kind = "reference" if reference_class() is not None else "port"
        cpu_baseline = {"value": p_asis / t_asis, "unit": "hyp-iter/s", "cores": procs, "kind": kind,
                        "sample": sample + (" (unmodified reference ficp.py from oracle/_ref: kd-tree rebuilt every pass, O(N^2) FRMSD loop)"
                                            if kind == "reference" else " (oracle port of the reference as shipped)"),
                        "seconds": t_asis,
                        "index_hoisted": {"value": p_h / t_h, "unit": "hyp-iter/s", "seconds": t_h, "kind": "port",
                                          "sample": sample + " (oracle port: kd-tree built once, cumsum FRMSD)"},
                        "single_core": {"as_shipped": p1 / t1, "index_hoisted": p1h / t1h, "unit": "hyp-iter/s",
                                        "sample": "hypothesis 0 of plot 0 on 1 process, ~6 s budget each"}}

    if rank == 0:
        line = {"metric": "FICP hypothesis-iterations/sec", "value": value, "unit": "hyp-iter/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": name, "plots_per_gpu": args.plots_per_gpu, "plots": n_plots,
                           "icps_per_gpu_per_step": (args.plots_per_gpu if one_pose else args.plots_per_gpu * hyp.shape[0]),
                           "hyp_iterations_per_step": passes_all, "parallelism": (f"plots round-robin over {world} GPU(s)" if by_plots else f"hypotheses round-robin over {world} GPU(s)"),
                           "exchange": "device-side pack + ONE all_gather of 112 B per plot (dist.exchange_best), inside the timed step",
                           "l2": "flushed between timed steps (256 MB write)", "launch": batch.info,
                           "device": props},
                "nn_queries_per_s": value * args.trees,
                "e2e": {"value": e2e_val, "unit": "hyp-iter/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                        "steps": e2e_steps, "api": "register_batch_distributed" if world > 1 else "register_batch",
                        "ms_per_step_rank0": [round(x, 3) for x in e2e_ms_list],
                        "note": "value = all steps' hypothesis-iterations / wall clock of the whole loop (max over ranks); the per-step list "
                                "(rank 0) shows whether one slow step - a host or PCIe hiccup on the box - moved it"},
                "e2e_resident_index": e2e_res,
                "gpu_launches": args.steps * 2,      # per timed step: the persistent ICP kernel + the pack kernel of the exchange
                "roofline": roofline, "clocks": sampler.summary(),
                "path_stats": stats}
        if cpu_baseline:
            line["cpu_baseline"] = cpu_baseline
        if single:
            line["single_stand"] = single
        if sweep:
            line["trim_fraction_sweep"] = sweep
        line.update(extra)
        print(json.dumps(line))
    batch.close()
    index.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    apply_workload_defaults(args)
    try:
        if args.impl == "reference":
            run_reference(args)
        else:
            run_b200(args)
    finally:
        if _POOL is not None:
            _POOL[0].terminate()
            _POOL[0].join()


if __name__ == "__main__":
    main()
