"""Bulk NN query: thread-per-query vs bulk kernel over grid densities (points per cell), whole call device-timed.
    python tools/nn_bulk_probe.py [dims] [log2 queries]"""
import ctypes as C, json, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coregistrationgame_b200 import TargetIndex, _lib, synthetic as syn

dims = int(sys.argv[1]) if len(sys.argv) > 1 else 3
nq = 1 << (int(sys.argv[2]) if len(sys.argv) > 2 else 22)
lib = _lib.load()
tgt, _, _ = syn.synthetic_scene(1_000_000, 50, seed=3, dims=3, n_plots=1, hidden_pose=False)
tgt = np.ascontiguousarray(tgt[:, :dims])
rng = np.random.default_rng(1)
lo, hi = tgt[:, :2].min(0), tgt[:, :2].max(0)
q = np.empty((nq, dims))
q[:, 0] = rng.uniform(lo[0], hi[0], nq); q[:, 1] = rng.uniform(lo[1], hi[1], nq)
if dims == 3:
    q[:, 2] = rng.uniform(5, 35, nq)
dev = torch.device("cuda")
dq = torch.from_numpy(q).to(dev)
didx = torch.empty(nq, dtype=torch.int32, device=dev); ddist = torch.empty(nq, dtype=torch.float64, device=dev)
ref = None
stream = torch.cuda.current_stream(); sp = C.c_void_p(stream.cuda_stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for ppc in (1.0, 1.5, 2.0, 3.0, 4.0, 6.0):
    ti = TargetIndex(tgt, pts_per_cell=ppc)
    row = {"dims": dims, "queries": nq, "pts_per_cell": ppc, "cell_m": ti.info()["cell"]}
    for name, k in (("thread", 1), ("bulk", 2)):
        call = lambda: _lib.check(lib.ficp_nn_query_device_ex(ti.handle, C.c_void_p(dq.data_ptr()), nq, dims, int(dims == 3),
                                                              C.c_void_p(didx.data_ptr()), C.c_void_p(ddist.data_ptr()), k, None, sp))
        call(); call(); torch.cuda.synchronize()
        ms = []
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            flush.fill_(1); a.record(stream); call(); b.record(stream); torch.cuda.synchronize()
            ms.append(a.elapsed_time(b))
        row[name + "_ms"] = float(np.median(ms)); row[name + "_Gqps"] = nq / np.median(ms) / 1e6
        got = (didx.cpu().numpy().copy(), ddist.cpu().numpy().copy())
        if ref is None:
            ref = got
        row[name + "_same_bits"] = bool(np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1]))
    cnt = (C.c_uint64 * 3)()
    _lib.check(lib.ficp_nn_query_device_ex(ti.handle, C.c_void_p(dq.data_ptr()), nq, dims, int(dims == 3), C.c_void_p(didx.data_ptr()),
                                           C.c_void_p(ddist.data_ptr()), 2, cnt, sp))
    row["resolved"] = {"window": int(cnt[0]), "global_grid": int(cnt[1]), "rings": int(cnt[2])}
    print(json.dumps(row), flush=True)
    ti.close()
