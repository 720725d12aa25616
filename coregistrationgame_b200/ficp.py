"""Drop-in ``FractionalICP`` backed by the sm_100a kernels in ``libficp_b200.so``.

Mirrors the public surface of the reference class (``/root/reference/ficp.py:5-154``): same
constructor arguments, attributes, method names, return shapes/dtypes and error behaviour, so
``app.py:658-661`` (``FractionalICP(src, tgt); run(); icp.source[:, :2]``) and the reference's own
tests work unchanged.  No numerics are done here - every method marshals numpy arrays into the
C ABI (``include/ficp_b200.h``).  There is no CPU fallback.

Differences a caller can observe (all documented in DESIGN.md):
  * nearest-neighbour and trim-order ties resolve to the lowest index (the reference's choice is
    traversal-order dependent);
  * after ``run()`` / ``_iterate()`` the attributes ``transform_`` (composed 3x3), ``frmsd_``,
    ``rmse_``, ``k_`` and ``n_passes_`` describe the result (the reference exposes none).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .batch import IDENTITY_HYPOTHESIS, STAGE2_LAMBDA, IcpBatch, TargetIndex, frmsd_weights

_KERNEL_MAX_N = 1024      # persistent kernel: trees per plot
_STEPWISE_MAX_N = 1 << 24  # stage kernels (one CTA in shared memory up to 8192 rows, global-scratch sort / scan above)


class FractionalICP:
    def __init__(self, source, target, lambda_val=3.0, threshold=1e-6, max_iterations=1000,
                 allow_reflection=False):
        """Fractional ICP, rigid in the plane: correspondences / FRMSD use XYZ when both inputs have a third
        column, the fitted transform moves XY only (ficp.py:6-44)."""
        self.source = np.array(source, dtype=float)
        self.target = np.array(target, dtype=float)
        if self.source.ndim != 2 or self.target.ndim != 2:
            raise ValueError("source and target must be 2D arrays (N, D).")
        self.match_dims = 3 if (self.source.shape[1] >= 3 and self.target.shape[1] >= 3) else 2
        self.lambda_val = lambda_val
        self.threshold = threshold
        self.max_iterations = max_iterations
        self.allow_reflection = allow_reflection
        self._index = None
        self._index_key = None
        self.transform_ = np.eye(3)
        self.frmsd_ = float("inf")
        self.rmse_ = float("nan")
        self.k_ = 0
        self.n_passes_ = 0

    # ----------------- helpers (ficp.py:47-51) -----------------
    def _xy(self, pts):
        return np.ascontiguousarray(pts[:, :2])

    def _xyz_or_xy(self, pts):
        return np.ascontiguousarray(pts[:, :self.match_dims])

    def _target_index(self, target):
        """Grid index over `target`, cached for self.target (built once, not once per pass)."""
        if target is self.target:
            # keyed on the CONTENT of the matched columns (xor + wrapping sum of the float64 bit patterns: two passes over
            # the array, ~5 ms per 1e6 points), so an in-place edit of self.target between calls rebuilds the grid - the
            # reference rebuilds its kd-tree from the current contents on every call (ficp.py:69)
            bits = self._xyz_or_xy(target).view(np.uint64)
            key = (id(target), target.shape, int(np.bitwise_xor.reduce(bits, axis=None)) if bits.size else 0,
                   int(bits.sum(dtype=np.uint64)) if bits.size else 0)
            if self._index is None or self._index_key != key:
                if self._index is not None:
                    self._index.close()
                self._index = TargetIndex(self._xyz_or_xy(target), use_z=(self.match_dims == 3))
                self._index_key = key
            return self._index, False
        return TargetIndex(self._xyz_or_xy(np.asarray(target, dtype=float)), use_z=(self.match_dims == 3)), True

    # ----------------- FRMSD & matching -----------------
    def frmsd(self, fraction, num_elements, subset_source, corresponding_targets):
        """Fractional RMSD in XYZ (XY without Z) - ficp.py:54-60; the sum runs on the device."""
        if num_elements == 0:
            return float("inf")
        a = self._xyz_or_xy(np.asarray(subset_source, dtype=float))
        b = self._xyz_or_xy(np.asarray(corresponding_targets, dtype=float))
        if a.shape != b.shape:
            raise ValueError(f"operands could not be broadcast together with shapes {a.shape} {b.shape}")
        out = C.c_double(0.0)
        _lib.require_device()
        _lib.check(_lib.load().ficp_sumsq(_lib.ptr(a), a.shape[1], _lib.ptr(b), b.shape[1], a.shape[0],
                                          self.match_dims, C.byref(out)), "ficp_sumsq")
        rmse = np.sqrt(out.value / num_elements)
        return (1.0 / (fraction ** self.lambda_val)) * rmse

    def _order(self, distances):
        d = np.ascontiguousarray(np.asarray(distances, dtype=np.float64).ravel())
        n = d.shape[0]
        order = np.empty(n, dtype=np.int64)
        if n:
            _lib.require_device()
            _lib.check(_lib.load().ficp_select_fraction(None, 0, None, 0, _lib.ptr(d), n, self.match_dims, None, 0,
                                                        None, None, _lib.ptr(order)), "ficp_select_fraction")
        return order

    def get_n_first_elements(self, num_elements, distances):
        """Indices of the `num_elements` smallest distances (stable order) - ficp.py:62-63."""
        return self._order(distances)[:num_elements]

    def find_correspondences(self, source, target):
        """Nearest target row of every source row and its distance - ficp.py:65-71."""
        if len(target) == 0 or len(source) == 0:
            return np.empty((0, target.shape[1])), np.array([])
        index, temporary = self._target_index(target)
        try:
            q = self._xyz_or_xy(np.asarray(source, dtype=float))
            idx, dists = index.query(q, use_z=(self.match_dims == 3))
        finally:
            if temporary:
                index.close()
        return target[idx], dists

    def find_optimal_fraction(self, corresponding_targets, distances):
        """Subset size minimising FRMSD over the distance-sorted prefixes - ficp.py:73-86."""
        n = len(self.source)
        if n == 0 or len(distances) == 0:
            return 0.0, 0
        if len(distances) != n or len(corresponding_targets) != n:
            raise ValueError("distances / correspondences must have one row per source point")
        src = self._xyz_or_xy(self.source)
        corr = self._xyz_or_xy(np.asarray(corresponding_targets, dtype=float))
        d = np.ascontiguousarray(np.asarray(distances, dtype=np.float64).ravel())
        w = frmsd_weights(n, self.lambda_val)
        k, val = C.c_int64(0), C.c_double(0.0)
        _lib.require_device()
        _lib.check(_lib.load().ficp_select_fraction(_lib.ptr(src), src.shape[1], _lib.ptr(corr), corr.shape[1],
                                                    _lib.ptr(d), n, self.match_dims, _lib.ptr(w), 0, C.byref(k),
                                                    C.byref(val), None), "ficp_select_fraction")
        return (k.value / n if k.value else 0.0), int(k.value)

    # ----------------- rigid 2D transform -----------------
    def compute_optimal_transform_2d(self, source_subset, target_subset):
        """Least-squares rotation + translation in the plane (no scale) - ficp.py:89-110."""
        a = self._xy(np.asarray(source_subset, dtype=float))
        b = self._xy(np.asarray(target_subset, dtype=float))
        if a.shape != b.shape:
            raise ValueError(f"shapes {a.shape} and {b.shape} not aligned")
        t9 = np.empty(9, dtype=np.float64)
        _lib.require_device()
        _lib.check(_lib.load().ficp_fit_rigid2d(_lib.ptr(a), 2, _lib.ptr(b), 2, a.shape[0],
                                                int(bool(self.allow_reflection)), _lib.ptr(t9)), "ficp_fit_rigid2d")
        return t9.reshape(3, 3)

    def apply_transform_2d_xy_only(self, points, T):
        """Moves XY by T, keeps Z and every other column bit-identical - ficp.py:112-119."""
        p = np.ascontiguousarray(np.asarray(points, dtype=np.float64))
        out = np.empty_like(p)
        t9 = np.ascontiguousarray(np.asarray(T, dtype=np.float64).reshape(9))
        if p.shape[0]:
            _lib.require_device()
            _lib.check(_lib.load().ficp_apply_xy(_lib.ptr(p), _lib.ptr(out), p.shape[0], p.shape[1], _lib.ptr(t9)),
                       "ficp_apply_xy")
        return out

    # ----------------- ICP loop -----------------
    def _run_stages(self, lambdas):
        n = len(self.source)
        if n == 0 or len(self.target) == 0:
            return self.source  # ficp.py:66-68,76-77,125-126: nothing to match, nothing moves
        if not np.isfinite(self.source[:, :self.match_dims]).all():
            raise ValueError("'x' must be finite, check for nan or inf values")
        if n > _STEPWISE_MAX_N:
            raise NotImplementedError(f"plots above {_STEPWISE_MAX_N} rows are not supported by the stage kernels "
                                      f"(got {n}); split the plot")
        if n > _KERNEL_MAX_N:
            return self._run_stages_resident(lambdas)
        index, _ = self._target_index(self.target)
        batch = IcpBatch(index, [self.source], IDENTITY_HYPOTHESIS, centres=np.zeros((1, 2)),
                         lambda_val=lambdas[0], stage2_lambda=(lambdas[1] if len(lambdas) > 1 else None),
                         n_stages=len(lambdas), threshold=self.threshold, max_iterations=self.max_iterations,
                         allow_reflection=self.allow_reflection, min_k=0, want_final_xy=True)
        try:
            out = batch.run().results()
        finally:
            batch.close()
        row = out["hyp"][0, 0]
        moved = self.source.copy()
        moved[:, :2] = out["final_xy"]
        self.source = moved
        step = np.eye(3)
        step[:2, :] = batch.transform_of(0, row)
        self.transform_ = step @ self.transform_
        self.frmsd_, self.rmse_, self.k_ = float(row["frmsd"]), float(row["rmse"]), int(row["k"])
        self.n_passes_ += int(row["passes"])
        return self.source

    def _run_stages_resident(self, lambdas):
        """Plots above the persistent kernels' 1024-tree limit: the loop of ficp.py:122-147 driven from here pass by pass
        (the convergence test stays in the reference's own expressions) over arrays that stay on the device
        (``ficp_stepper_*``).  Same stage kernels on the same values as `_iterate_stepwise` - same bits - without its six
        host round trips per pass."""
        lib = _lib.load()
        _lib.require_device()
        n, md = len(self.source), self.match_dims
        index, _ = self._target_index(self.target)
        src = self._xyz_or_xy(self.source)
        h = C.c_void_p()
        _lib.check(lib.ficp_stepper_create(index.handle, _lib.ptr(src), n, md, md, C.byref(h)), "ficp_stepper_create")
        k, ss = C.c_int64(0), C.c_double(0.0)
        t9 = np.empty(9, dtype=np.float64)

        def one_pass():
            _lib.check(lib.ficp_stepper_pass(h, 0, C.byref(k), C.byref(ss)), "ficp_stepper_pass")
            self.n_passes_ += 1
            if k.value == 0:
                return 0, float("inf")
            frac = k.value / n
            rmse = np.sqrt(ss.value / k.value)                       # the expressions of frmsd(), ficp.py:54-60
            self.rmse_ = float(rmse)
            return int(k.value), (1.0 / (frac ** self.lambda_val)) * rmse
        try:
            for lam in lambdas:
                self.lambda_val = lam
                w = frmsd_weights(n, lam)
                _lib.check(lib.ficp_stepper_set_weights(h, _lib.ptr(w)), "ficp_stepper_set_weights")
                kk, score = one_pass()
                if kk == 0:
                    continue
                for _ in range(self.max_iterations):
                    _lib.check(lib.ficp_stepper_fit_apply(h, int(bool(self.allow_reflection)), _lib.ptr(t9)), "ficp_stepper_fit_apply")
                    self.transform_ = t9.reshape(3, 3).copy() @ self.transform_
                    kk, new_score = one_pass()
                    self.frmsd_, self.k_ = float(new_score), int(kk)
                    if score - new_score <= self.threshold or kk == 0:
                        break
                    score = new_score
            xy = np.empty((n, 2), dtype=np.float64)
            _lib.check(lib.ficp_stepper_read_xy(h, _lib.ptr(xy)), "ficp_stepper_read_xy")
        finally:
            lib.ficp_stepper_destroy(h)
        moved = self.source.copy()
        moved[:, :2] = xy
        self.source = moved
        return self.source

    def _pass_stepwise(self):
        """One NN pass + trimming through the stage kernels: (correspondences, trimmed rows, k, FRMSD at k)."""
        matched, dist = self.find_correspondences(self.source, self.target)
        self.n_passes_ += 1
        frac, k = self.find_optimal_fraction(matched, dist)
        if k == 0:
            return matched, None, 0, float("inf")
        rows = self.get_n_first_elements(k, dist)
        return matched, rows, k, self.frmsd(frac, k, self.source[rows], matched[rows])

    def _iterate_stepwise(self):
        """One stage driven from the host over the stage kernels, for plots above the persistent kernel's 1024-tree
        limit: pass, then (fit, move, pass) until the FRMSD stops improving by more than `threshold` - the loop of
        ficp.py:122-147, including its habit of keeping the pose of the last (possibly worse) pass."""
        matched, rows, k, score = self._pass_stepwise()
        if k == 0:
            return self.source
        for _ in range(self.max_iterations):
            step = self.compute_optimal_transform_2d(self.source[rows], matched[rows])
            self.source = self.apply_transform_2d_xy_only(self.source, step)
            self.transform_ = step @ self.transform_
            matched, rows, k, new_score = self._pass_stepwise()
            self.frmsd_, self.k_ = float(new_score), int(k)
            if score - new_score <= self.threshold:
                break
            score = new_score
        return self.source

    def _iterate(self):
        """One stage with the current ``lambda_val`` - ficp.py:122-147."""
        return self._run_stages([self.lambda_val])

    def run(self):
        """Two-stage Fractional ICP - ficp.py:149-154.  Leaves ``lambda_val`` at the stage-2 value, like the
        reference."""
        lam2 = STAGE2_LAMBDA[self.match_dims]
        self._run_stages([self.lambda_val, lam2])
        self.lambda_val = lam2
        return self.source

    def __del__(self):
        try:
            if self._index is not None:
                self._index.close()
        except Exception:
            pass
