"""Host side of the batched Fractional-ICP search (extension over the reference API).

The reference runs one plot from one start pose per call (``app.py:630-661``).  This module
batches ``FractionalICP(pre_transform(src, h), tgt).run()`` over many plots and many start-pose
hypotheses ``h`` (rotation grid x flip x coarse translation, with the semantics of
``trees.py:165-222``) and picks the best registration per plot.  All arithmetic happens in the
CUDA kernels behind ``libficp_b200.so``; this file only marshals arrays.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _lib

STAGE2_LAMBDA = {2: 1.3, 3: 0.95}  # ficp.py:152


# --------------------------------------------------------------------------- hypothesis tables
def hypothesis_matrix(theta_deg, flip):
    """2x2 linear part of a start pose: y-flip (optional) then CCW rotation, both about the plot
    centre - ``Plot.coordinate_flip`` / ``Plot.rotate_plot`` (trees.py:165-222)."""
    th = np.radians(theta_deg)
    c, s = np.cos(th), np.sin(th)
    if flip:
        return np.array([[c, s], [s, -c]])
    return np.array([[c, -s], [s, c]])


def hypothesis_table(n_rot, flips=(0, 1), translations=((0.0, 0.0),)):
    """(H, 6) float64 rows ``[m00 m01 m10 m11 dx dy]``; translation-major, then flip, then rotation."""
    rows = []
    for (dx, dy) in translations:
        for f in flips:
            for r in range(n_rot):
                m = hypothesis_matrix(360.0 * r / n_rot, f)
                rows.append([m[0, 0], m[0, 1], m[1, 0], m[1, 1], dx, dy])
    return np.array(rows, dtype=np.float64).reshape(-1, 6)


def translation_lattice(n_side, pitch):
    off = (np.arange(n_side) - (n_side - 1) / 2.0) * pitch
    return [(float(dx), float(dy)) for dy in off for dx in off]


IDENTITY_HYPOTHESIS = np.array([[1.0, 0.0, 0.0, 1.0, 0.0, 0.0]])


def plot_centres(src, offsets):
    """Centre of every plot = ``rows[:, :2].mean(axis=0)`` of its rows (what the oracle and the reference-side callers
    compute; `Plot.rotate_plot` / `coordinate_flip` turn about it, trees.py:201-222), for ten thousand plots without ten
    thousand numpy calls: ``ficp_plot_centres`` adds the rows of a plot in order and divides once - the same additions in
    the same order, bit for bit (pinned by tests/test_host_cabi.py) - on a few host threads (no device involved)."""
    src = np.ascontiguousarray(np.asarray(src, dtype=np.float64))
    offsets = np.ascontiguousarray(np.asarray(offsets, dtype=np.int64))
    out = np.empty((offsets.shape[0] - 1, 2), dtype=np.float64)
    if out.shape[0]:
        _lib.check(_lib.load().ficp_plot_centres(_lib.ptr(src), src.shape[1], _lib.ptr(offsets), out.shape[0], _lib.ptr(out)),
                   "ficp_plot_centres")
    return out


def stack_plots(sources):
    """Plots as ONE C-contiguous float64 array + offsets + sizes (plot p owns ``rows[offsets[p]:offsets[p + 1]]``).

    ``sources``: a list of (N_p, D) arrays, one (N, D) array (a single plot), or - the form for thousands of plots, where
    the per-plot Python work of a list is what bounds config 4 end to end - a tuple ``(rows, offsets)`` already stacked."""
    if isinstance(sources, tuple) and len(sources) == 2 and isinstance(sources[0], np.ndarray) and sources[0].ndim == 2 \
            and np.ndim(sources[1]) == 1:
        rows, offs = sources
        src = np.ascontiguousarray(np.asarray(rows, dtype=np.float64))
        offsets = np.ascontiguousarray(np.asarray(offs, dtype=np.int64))
        sizes = np.diff(offsets)
        if offsets.size < 2 or offsets[0] != 0 or offsets[-1] != src.shape[0] or (sizes <= 0).any():
            raise ValueError("offsets must start at 0, end at the number of rows and describe non-empty plots")
        return src, offsets, sizes
    if isinstance(sources, np.ndarray) and sources.ndim == 2:
        sources = [sources]
    srcs = list(sources)
    if not all(type(s) is np.ndarray and s.dtype == np.float64 for s in srcs):
        srcs = [np.asarray(s, dtype=np.float64) for s in srcs]
    if not srcs or any(s.ndim != 2 or s.shape[0] == 0 for s in srcs):
        raise ValueError("every plot must be a non-empty 2D array (N, D)")
    sizes = np.fromiter(map(len, srcs), dtype=np.int64, count=len(srcs))
    offsets = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    try:
        # one copy; checks the column counts itself (Fortran-ordered plots give a Fortran-ordered stack: made C after)
        src = np.ascontiguousarray(np.concatenate(srcs, axis=0))
    except ValueError:
        raise ValueError("all plots must have the same number of columns") from None
    return src, offsets, sizes


def frmsd_weights(n, lam):
    """w[k-1] = 1/((k/n)**lam) with Python-float semantics - the very expression of ficp.py:60,81,
    so the table is bit-identical to what the reference multiplies with."""
    return np.array([1.0 / ((k / n) ** lam) for k in range(1, n + 1)], dtype=np.float64)


def fixed_fraction_k(n, frac):
    return max(1, min(n, int(math.floor(frac * n + 1e-9))))


def _stream_ptr(stream):
    if stream is None:
        return None
    if hasattr(stream, "cuda_stream"):   # torch.cuda.Stream
        return C.c_void_p(stream.cuda_stream)
    return C.c_void_p(int(stream))


# --------------------------------------------------------------------------- target index
class TargetIndex:
    """Device-resident uniform-grid index over the Layer-2 (CHM) points; built once, reused by
    every pass of every hypothesis (the reference rebuilds its kd-tree on every pass, ficp.py:69)."""

    # points per grid cell, measured on B200 (profiles/r01_summary.md).  The ICP kernel wants cells wide enough that the
    # 3x3 block around a query settles its search even for poor start poses (3-D residuals of 6-8 m) and gives its
    # skip test a long leash; bulk one-shot queries (no previous neighbour to bound the search) want small cells.
    PTS_PER_CELL = {("icp", True): 6.0, ("icp", False): 3.0, ("query", True): 3.0, ("query", False): 2.0}

    def __init__(self, target, use_z=None, pts_per_cell=None, stream=None, purpose="icp"):
        lib = _lib.load()
        arr = np.ascontiguousarray(np.asarray(target, dtype=np.float64))
        if arr.ndim != 2 or arr.shape[1] < 2:
            raise ValueError("source and target must be 2D arrays (N, D).")
        self.m, self.ld = int(arr.shape[0]), int(arr.shape[1])
        self.has_z = bool(arr.shape[1] >= 3) if use_z is None else bool(use_z)
        if pts_per_cell is None:
            pts_per_cell = self.PTS_PER_CELL[(purpose, self.has_z)]
        self._h = C.c_void_p()
        _lib.require_device()
        _lib.check(lib.ficp_target_create(_lib.ptr(arr), self.m, self.ld, int(self.has_z), float(pts_per_cell),
                                          _stream_ptr(stream), C.byref(self._h)), "ficp_target_create")

    @property
    def handle(self):
        if not self._h:
            raise _lib.FicpError("TargetIndex already closed")
        return self._h

    def info(self):
        ti = _lib.TargetInfo()
        _lib.check(_lib.load().ficp_target_get_info(self.handle, C.byref(ti)))
        return {"m": ti.m, "has_z": bool(ti.has_z), "grid_w": ti.grid_w, "grid_h": ti.grid_h, "cell": ti.cell,
                "x0": ti.x0, "y0": ti.y0, "bbox": tuple(ti.bbox), "build_ms": ti.build_ms, "clamped": bool(ti.clamped),
                "max_cell_pts": int(ti.max_cell_pts)}

    QUERY_KERNELS = {"auto": 0, "thread": 1, "bulk": 2}

    def query(self, points, use_z=None, stream=None, kernel="auto", counters=None):
        """Exact NN of every row: (original target index int64, Euclidean distance float64).
        kernel: "auto" (bulk kernel for large, dense batches), "thread" (one thread per query), "bulk" (cell-ordered
        queries against shared-memory windows); identical bits either way.  counters: optional dict, filled by the bulk
        kernel with how its queries were resolved."""
        q = np.ascontiguousarray(np.asarray(points, dtype=np.float64))
        if q.ndim != 2 or q.shape[1] < 2:
            raise ValueError("query points must be a 2D array (N, D>=2)")
        z = self.has_z and q.shape[1] >= 3 if use_z is None else bool(use_z)
        md = 3 if z else 2
        if not np.isfinite(q[:, :md]).all():
            raise ValueError("'x' must be finite, check for nan or inf values")
        n = q.shape[0]
        idx = np.empty(n, dtype=np.int64)
        dist = np.empty(n, dtype=np.float64)
        if n and self.m:
            cnt = (C.c_uint64 * 3)() if counters is not None else None
            _lib.check(_lib.load().ficp_nn_query_ex(self.handle, _lib.ptr(q), n, q.shape[1], int(z), _lib.ptr(idx),
                                                    _lib.ptr(dist), self.QUERY_KERNELS[kernel], cnt, _stream_ptr(stream)),
                       "ficp_nn_query")
            if counters is not None:
                counters.update(window=int(cnt[0]), global_grid=int(cnt[1]), rings=int(cnt[2]))
        return idx, dist

    def close(self):
        if getattr(self, "_h", None):
            _lib.load().ficp_target_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# --------------------------------------------------------------------------- batch
class IcpBatch:
    """Plots x hypotheses uploaded once; ``run()`` launches the persistent kernel, ``results()`` reads back."""

    def __init__(self, index, sources, hyp_table=None, centres=None, lambda_val=3.0, stage2_lambda=None, n_stages=2,
                 threshold=1e-6, max_iterations=1000, allow_reflection=False, min_k=3, fixed_frac=None,
                 hyp_shard=(0, 1), want_final_xy=False, window_margin=-1.0, warps_per_cta=0, ctas_per_sm=0,
                 disable_window=False, team_warps=0, helpers=None, trace_passes=0, cta_per_icp=None, stream=None):
        lib = _lib.load()
        self.src, self.offsets, self.sizes = stack_plots(sources)
        ld = self.src.shape[1]
        self.index = index
        self.match_dims = 3 if (ld >= 3 and index.has_z) else 2
        self.n_plots = int(self.sizes.shape[0])
        self.hyp = np.ascontiguousarray(IDENTITY_HYPOTHESIS if hyp_table is None else
                                        np.asarray(hyp_table, dtype=np.float64).reshape(-1, 6))
        # centres=None: the library takes each plot's own centroid while it prepares the rows (ficp_batch_create with
        # centres = NULL); the `centres` property computes the same bits on first use (winners-only callers never need them)
        self._centres = None if centres is None else \
            np.ascontiguousarray(np.asarray(centres, dtype=np.float64).reshape(self.n_plots, 2))
        self.n_stages = int(n_stages)
        lam2 = STAGE2_LAMBDA[self.match_dims] if stage2_lambda is None else stage2_lambda
        self.lambdas = [lambda_val, lam2][: self.n_stages]
        # FRMSD weight tables, one per distinct plot size
        uniq = np.unique(self.sizes).tolist()
        tab_of = {n: i for i, n in enumerate(uniq)}
        tabs, offs = [], [0]
        for n in uniq:
            for lam in self.lambdas:
                tabs.append(frmsd_weights(n, lam))
            offs.append(offs[-1] + self.n_stages * n)
        self.weights = np.ascontiguousarray(np.concatenate(tabs))
        self.weight_offsets = np.array(offs, dtype=np.int64)
        self.plot_tab = np.searchsorted(np.array(uniq, dtype=np.int64), self.sizes).astype(np.int32)
        self.fixed_k = None
        if fixed_frac is not None:
            k_of = {n: fixed_fraction_k(n, fixed_frac) for n in uniq}
            self.fixed_k = np.array([k_of[int(n)] for n in self.sizes], dtype=np.int32)
        self.hyp_begin, self.hyp_stride = int(hyp_shard[0]), int(hyp_shard[1])
        prm = _lib.BatchParams(self.n_stages, int(max_iterations), int(bool(allow_reflection)), int(min_k),
                               float(threshold), float(window_margin), int(warps_per_cta), int(ctas_per_sm),
                               int(bool(disable_window)), int(team_warps), (0 if helpers is None else (2 if helpers else 1)),
                               int(trace_passes), (0 if cta_per_icp is None else (2 if cta_per_icp else 1)), 0)
        self._h = C.c_void_p()
        _lib.check(lib.ficp_batch_create(index.handle, _lib.ptr(self.src), ld, int(self.match_dims == 3),
                                         _lib.ptr(self.offsets), self.n_plots, _lib.ptr(self._centres),
                                         _lib.ptr(self.hyp), self.hyp.shape[0], self.hyp_begin, self.hyp_stride,
                                         _lib.ptr(self.weights), _lib.ptr(self.weight_offsets),
                                         _lib.ptr(self.plot_tab), len(uniq), _lib.ptr(self.fixed_k), C.byref(prm),
                                         int(bool(want_final_xy)), _stream_ptr(stream), C.byref(self._h)),
                   "ficp_batch_create")
        bi = _lib.BatchInfo()
        _lib.check(lib.ficp_batch_get_info(self._h, C.byref(bi)))
        self.info = {f: getattr(bi, f) for f, _ in bi._fields_}
        self.n_hyp = self.hyp.shape[0]
        self.n_hyp_local = bi.n_hyp_local
        self.want_final_xy = bool(want_final_xy) and self.n_hyp_local == 1
        self.h2d_bytes = int(self.src.nbytes + self.hyp.nbytes + 16 * self.n_plots + self.weights.nbytes)

    @property
    def centres(self):
        """(n_plots, 2) point the start poses of each plot turn about (given, or each plot's centroid)."""
        if self._centres is None:
            self._centres = plot_centres(self.src, self.offsets)
        return self._centres

    def run(self, stream=None):
        _lib.check(_lib.load().ficp_batch_run(self._h, _stream_ptr(stream)), "ficp_batch_run")
        return self

    def copy_best_keys_to(self, device_ptr, stream=None):
        _lib.check(_lib.load().ficp_batch_copy_best_keys_device(self._h, C.c_void_p(int(device_ptr)),
                                                                _stream_ptr(stream)))

    def pack_best_to(self, device_ptr, stream=None):
        """Enqueue only: this GPU's best registration per plot as (n_plots, 12) int64 words in device memory at
        ``device_ptr`` - key, the 80-byte result row, this GPU's hypothesis-iterations, the world-frame translation
        (see dist.PACK_WORDS)."""
        _lib.check(_lib.load().ficp_batch_pack_best_device(self._h, C.c_void_p(int(device_ptr)), _stream_ptr(stream)))

    def best(self, stream=None):
        """Synchronise and read back only the winner per plot (112 bytes each) - the per-hypothesis table stays on the device.
        Returns ``best_key``, ``best_hyp``, ``best_score``, ``best_row`` (HYP_RESULT_DTYPE per plot) and ``stats``."""
        packed = np.empty((self.n_plots, _lib.PACK_WORDS), dtype=np.uint64)
        stats = np.zeros(8, dtype=np.uint64)
        _lib.check(_lib.load().ficp_batch_best(self._h, _lib.ptr(packed), _lib.ptr(stats), _stream_ptr(stream)), "ficp_batch_best")
        keys = np.ascontiguousarray(packed[:, 0])
        out = {"best_key": keys, "best_row": np.ascontiguousarray(packed[:, 1:11]).view(_lib.HYP_RESULT_DTYPE).reshape(-1),
               "best_b": np.ascontiguousarray(packed[:, 12:14]).view(np.float64),   # world-frame translation: final = M p + b
               "hyp": None, "final_xy": None,
               "stats": {"passes": int(stats[0]), "global_path_queries": int(stats[1]),
                         "windows_disabled": int(stats[2]), "fixup_rounds": int(stats[3]), "queries": int(stats[4]),
                         "searched_queries": int(stats[5]), "deferred_queries": int(stats[6]),
                         "order_rebuilds": int(stats[7])}}
        out.update(decode_best_keys(keys))
        self.d2h_bytes = int(packed.nbytes + stats.nbytes)
        return out

    def results(self, stream=None, per_hypothesis=True):
        """Synchronise and read back.  Returns a dict:
        ``hyp`` structured array (n_plots, n_hyp_local) of per-hypothesis outcomes (see HYP_RESULT_DTYPE),
        ``hyp_ids`` global hypothesis id of each local column, ``best_key`` uint64 per plot, ``best_hyp``,
        ``best_score``, ``stats`` and (optionally) ``final_xy``."""
        lib = _lib.load()
        res = np.empty((self.n_plots, self.n_hyp_local), dtype=_lib.HYP_RESULT_DTYPE) if per_hypothesis else None
        keys = np.empty(self.n_plots, dtype=np.uint64)
        stats = np.zeros(8, dtype=np.uint64)
        final = np.empty((int(self.offsets[-1]), 2), dtype=np.float64) if self.want_final_xy else None
        _lib.check(lib.ficp_batch_results(self._h, _lib.ptr(res), _lib.ptr(keys), _lib.ptr(final), _lib.ptr(stats),
                                          _stream_ptr(stream)), "ficp_batch_results")
        out = {"hyp": res, "best_key": keys, "final_xy": final,
               "hyp_ids": self.hyp_begin + self.hyp_stride * np.arange(self.n_hyp_local),
               "stats": {"passes": int(stats[0]), "global_path_queries": int(stats[1]),
                         "windows_disabled": int(stats[2]), "fixup_rounds": int(stats[3]), "queries": int(stats[4]),
                         "searched_queries": int(stats[5]), "deferred_queries": int(stats[6]),
                         "order_rebuilds": int(stats[7])}}
        out.update(decode_best_keys(keys))
        self.d2h_bytes = int(keys.nbytes + stats.nbytes + (res.nbytes if res is not None else 0)
                             + (final.nbytes if final is not None else 0))
        return out

    def trace(self, stream=None):
        """Per-pass trace of a batch created with ``trace_passes > 0`` (test instrument).  Returns arrays indexed
        ``[plot, local hypothesis, pass, tree]``: ``idx`` original target row of every tree's nearest neighbour
        (ficp.py:70), ``d2`` its squared distance, ``inlier`` membership in the trimmed subset (ficp.py:62-63,133), and
        ``k`` / ``frmsd`` ``[plot, local hypothesis, pass]``.  Only the first ``min(passes, trace_passes)`` passes of an
        ICP and the first ``n`` trees of a plot are meaningful."""
        cap, stride = int(self.info["trace_passes"]), int(self.info["trace_stride"])
        if cap <= 0:
            raise ValueError("batch was created without trace_passes")
        shape = (self.n_plots, self.n_hyp_local, cap)
        idx = np.empty(shape + (stride,), dtype=np.int32)
        d2 = np.empty(shape + (stride,), dtype=np.float64)
        inl = np.empty(shape + (stride,), dtype=np.uint8)
        k = np.empty(shape, dtype=np.int32)
        f = np.empty(shape, dtype=np.float64)
        _lib.check(_lib.load().ficp_batch_trace(self._h, _lib.ptr(idx), _lib.ptr(d2), _lib.ptr(inl), _lib.ptr(k),
                                                _lib.ptr(f), _stream_ptr(stream)), "ficp_batch_trace")
        return {"idx": idx, "d2": d2, "inlier": inl.astype(bool), "k": k, "frmsd": f}

    def transform_of(self, plot, hyp_row):
        """2x3 world-coordinate transform [A | b] of one result row: final = A p + b."""
        return compose_world_transform(hyp_row, self.centres[plot])

    def close(self):
        if getattr(self, "_h", None):
            _lib.load().ficp_batch_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def decode_best_keys(keys):
    keys = np.asarray(keys, dtype=np.uint64)
    score = (keys >> np.uint64(32)).astype(np.uint32).view(np.float32).astype(np.float64)
    return {"best_hyp": (keys & np.uint64(0xFFFFFFFF)).astype(np.int64), "best_score": score}


def compose_world_transforms(rows, centres):
    """final = M (p - centre) + c  ->  [M | c - M centre] for many result rows at once: (n, 2, 3).  Elementwise arithmetic in
    one fixed order, so one row or ten thousand give the same bits."""
    f = lambda name: np.asarray(rows[name], dtype=np.float64).reshape(-1)
    c = np.asarray(centres, dtype=np.float64).reshape(-1, 2)
    m00, m01, m10, m11 = f("m00"), f("m01"), f("m10"), f("m11")
    out = np.empty((m00.shape[0], 2, 3), dtype=np.float64)
    out[:, 0, 0], out[:, 0, 1], out[:, 1, 0], out[:, 1, 1] = m00, m01, m10, m11
    out[:, 0, 2] = f("cx") - (m00 * c[:, 0] + m01 * c[:, 1])
    out[:, 1, 2] = f("cy") - (m10 * c[:, 0] + m11 * c[:, 1])
    return out


def world_transforms(rows, b):
    """[M | b] per plot from winners' rows and the world-frame translations the device packed with them (it evaluates
    b = c - M centre in the arithmetic of compose_world_transforms: same bits, and a receiver needs no plot centres)."""
    out = np.empty((rows.shape[0], 2, 3), dtype=np.float64)
    out[:, 0, 0], out[:, 0, 1], out[:, 1, 0], out[:, 1, 1] = rows["m00"], rows["m01"], rows["m10"], rows["m11"]
    out[:, :, 2] = b
    return out


def compose_world_transform(row, centre):
    """2x3 world transform [A | b] of ONE result row (final = A p + b)."""
    return compose_world_transforms(row, centre)[0]


def register_batch(sources, target, hyp_table=None, index=None, per_hypothesis=True, **kw):
    """One-call batched registration from HOST arrays (the end-to-end path bench.py times as ``e2e``):
    build the target index (unless one is passed), upload plots + hypotheses, run, read back.

    Returns per plot: best hypothesis id, its score (final FRMSD, fp32-rounded), its 2x3 transform,
    trimmed size ``k``, RMSE and the number of passes (``best_row``); with ``per_hypothesis`` (default) also the whole
    per-hypothesis table ``hyp`` - without it only the winners (112 bytes per plot) leave the device."""
    own = index is None
    if own:
        index = TargetIndex(target)
    try:
        batch = IcpBatch(index, sources, hyp_table, **kw)
        try:
            if per_hypothesis:
                out = batch.run().results()
                j = (out["best_hyp"] - batch.hyp_begin) // batch.hyp_stride
                out["best_row"] = out["hyp"][np.arange(batch.n_plots), j].copy()
            else:
                out = batch.run().best()
            out["best_transform"] = (compose_world_transforms(out["best_row"], batch.centres) if per_hypothesis
                                     else world_transforms(out["best_row"], out["best_b"]))
            out["h2d_bytes"] = batch.h2d_bytes + (int(np.asarray(target).nbytes) if own else 0)
            out["d2h_bytes"] = batch.d2h_bytes
            out["launch"] = dict(batch.info)
            return out
        finally:
            batch.close()
    finally:
        if own:
            index.close()
