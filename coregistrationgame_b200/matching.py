"""Steps right after the Fractional ICP in the reference application.

``remove_matches``   - ``CHMPlot.remove_matches`` (chm_plot.py:223-285): each tree of a confirmed plot, in order,
                       takes its nearest remaining CHM tree and removes it when it is closer than
                       ``min_dist_percent`` % of the tree's height.  CUDA: one exact grid search per tree with the
                       already-removed points masked out (``ficp_match_remove``).
``transform_record`` - ``Plot.get_transform`` + ``App.store_transformations`` (trees.py:248-280, app.py:884-924): the
                       rigid transform original -> final coordinates as the CSV record
                       ``{tx, ty, r00, r01, r10, r11, flip}``; the Procrustes fit runs in ``ficp_fit_rigid2d``.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .batch import TargetIndex, _stream_ptr


def _heights_ok(h):
    h = np.asarray(h, dtype=np.float64)
    return bool(np.isfinite(h).all())


def match_thresholds(heights, min_dist_percent, use_3d):
    """Per-tree distance threshold, the expression of chm_plot.py:262,281: ``(pct / 100.0) * height``; in the XY
    fallback a missing height counts as 10 m (chm_plot.py:274-280)."""
    h = np.asarray(heights, dtype=np.float64).copy()
    if not use_3d:
        h[~np.isfinite(h)] = 10.0
    return np.array([(min_dist_percent / 100.0) * float(v) for v in h], dtype=np.float64)


def remove_matches(plot_trees, chm, min_dist_percent=15, index=None, stream=None):
    """Greedy match-and-remove for ONE plot.

    plot_trees : (n, 3) rows ``[currentx, currenty, height]`` in plot order (height may be NaN).
    chm        : (M, 3) rows ``[currentx, currenty, height]`` of the CHM layer in its current order.
    Returns ``matched`` (n,) int64: the CHM row removed by each tree or -1; the removed rows in removal order are
    ``matched[matched >= 0]`` (what the reference appends to ``removed_stems``)."""
    out = remove_matches_batch([plot_trees], chm, min_dist_percent, index=index, stream=stream)
    return out[0]


def remove_matches_batch(plots, chm, min_dist_percent=15, index=None, stream=None):
    """Same for many plots at once, each against the SAME (unmodified) CHM layer: plots do not see each other's
    removals - use it for plots that cannot compete for the same CHM trees, or call ``remove_matches`` plot by plot
    like the reference does."""
    plots = [np.ascontiguousarray(np.asarray(p, dtype=np.float64).reshape(-1, 3)) for p in plots]
    chm = np.ascontiguousarray(np.asarray(chm, dtype=np.float64).reshape(-1, 3))
    # 3-D only when every height on both sides is present (chm_plot.py:240-249)
    # (a resident index built WITH heights has already refused non-finite ones at build time: no second scan of the layer -
    # 15 of the 21 ms of a one-plot call against 1e7 CHM points)
    chm_ok = True if (index is not None and index.has_z) else _heights_ok(chm[:, 2])
    use_3d = chm_ok and all(_heights_ok(p[:, 2]) for p in plots)
    sizes = np.array([len(p) for p in plots], dtype=np.int64)
    offsets = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    rows = int(offsets[-1])
    matched = np.full(rows, -1, dtype=np.int64)
    if rows and len(chm):
        trees = np.ascontiguousarray(np.vstack(plots))
        thr = np.concatenate([match_thresholds(p[:, 2], min_dist_percent, use_3d) for p in plots])
        if not use_3d:
            trees = np.ascontiguousarray(trees[:, :2])
        own = index is None
        if own:
            index = TargetIndex(chm[:, :3] if use_3d else chm[:, :2], use_z=use_3d, purpose="query")
        try:
            _lib.check(_lib.load().ficp_match_remove(index.handle, _lib.ptr(trees), _lib.ptr(offsets), len(plots),
                                                     trees.shape[1], int(use_3d), _lib.ptr(thr), _lib.ptr(matched),
                                                     _stream_ptr(stream)), "ficp_match_remove")
        finally:
            if own:
                index.close()
    return [matched[offsets[i]:offsets[i + 1]] for i in range(len(plots))]


def transform_record(original_xy, current_xy, flipped=False):
    """Rigid transform original -> current as the reference's CSV record (app.py:901-912).  ``current ~ R @ original + t``;
    a reflection is allowed only when the plot was flipped (trees.py:273-276)."""
    a = np.ascontiguousarray(np.asarray(original_xy, dtype=np.float64)[:, :2])
    b = np.ascontiguousarray(np.asarray(current_xy, dtype=np.float64)[:, :2])
    if a.shape != b.shape or a.shape[0] == 0:
        raise ValueError("No trees available to compute transform.")
    t9 = np.empty(9, dtype=np.float64)
    _lib.require_device()
    _lib.check(_lib.load().ficp_fit_rigid2d(_lib.ptr(a), 2, _lib.ptr(b), 2, a.shape[0], int(bool(flipped)), _lib.ptr(t9)),
               "ficp_fit_rigid2d")
    T = t9.reshape(3, 3)
    return {"tx": float(T[0, 2]), "ty": float(T[1, 2]), "r00": float(T[0, 0]), "r01": float(T[0, 1]),
            "r10": float(T[1, 0]), "r11": float(T[1, 1]), "flip": bool(flipped)}


def radial_crop(index_or_points, x, y, dist, stream=None):
    """Rows of the CHM layer within `dist` of (x, y) in XY, in their original order - the crop of
    ``CHMPlot`` / ``SavedPlot`` (chm_plot.py:144-148, :306-311: ``cdist(coordinates, [[x, y]]) <= dist``).
    Accepts a built ``TargetIndex`` (cell-range query, O(points near the disc)) or an (M, >=2) array."""
    own = not isinstance(index_or_points, TargetIndex)
    index = TargetIndex(np.asarray(index_or_points, dtype=np.float64)[:, :2], use_z=False, purpose="query") if own else index_or_points
    try:
        mask = np.zeros(index.m, dtype=np.uint8)
        if index.m:
            _lib.check(_lib.load().ficp_radial_crop(index.handle, float(x), float(y), float(dist), _lib.ptr(mask),
                                                    _stream_ptr(stream)), "ficp_radial_crop")
        return np.flatnonzero(mask)
    finally:
        if own:
            index.close()


def gui_hypothesis_table(rot_steps=range(-36, 36), trans_steps=(0,), flips=(0,), rot_step_deg=5.0, translate_step=0.5):
    """Start poses reachable with the reference's keys: rotations in multiples of 5 degrees
    (``App.rotate_plot``, app.py:618-624), translations in multiples of ``TRANSLATE_STEP`` = 0.5 m
    (``App.shift_plot``, app.py:604-616), optional flip (``Plot.coordinate_flip``, trees.py:213-222).
    Rows ``[m00 m01 m10 m11 dx dy]`` in the order translation (dy-major) x flip x rotation."""
    from .batch import hypothesis_matrix
    rows = []
    for ty in trans_steps:
        for tx in trans_steps:
            for f in flips:
                for r in rot_steps:
                    m = hypothesis_matrix(rot_step_deg * r, f)
                    rows.append([m[0, 0], m[0, 1], m[1, 0], m[1, 1], translate_step * tx, translate_step * ty])
    return np.array(rows, dtype=np.float64).reshape(-1, 6)


def key_hypothesis(rot_steps=0, flip=0, tx_steps=0, ty_steps=0, rot_step_deg=5.0, translate_step=0.5):
    """The start pose a user reaches with the reference's keys, as ONE hypothesis row ``[m00 m01 m10 m11 dx dy]``:
    `rot_steps` presses of rotate-left (+5 degrees each, negative = rotate-right; ``App.rotate_plot``, app.py:618-624),
    `flip` presses of flip modulo 2 (``Plot.coordinate_flip``, trees.py:213-222), `tx_steps` / `ty_steps` presses of
    right / down (+0.5 m each; ``App.shift_plot``, app.py:604-616).  Rotations and flips act about the plot centroid
    (``Plot.current_center``), which only the shifts move, so any interleaving of the keys gives
    ``x' = R(5 rot_steps) F^flip (x - c) + c + 0.5 (tx_steps, ty_steps)`` - checked against the unmodified ``Plot`` class in
    tests/test_plot_keys.py (goldens: tests/golden/make_golden_keys.py)."""
    from .batch import hypothesis_matrix
    m = hypothesis_matrix(rot_step_deg * rot_steps, int(flip) % 2)
    return np.array([m[0, 0], m[0, 1], m[1, 0], m[1, 1], translate_step * tx_steps, translate_step * ty_steps], dtype=np.float64)


def write_back(plot, new_xy):
    """``Plot.update_tree_positions`` (trees.py:296-314) for any plot-like object: sets ``currentx`` / ``currenty`` of every
    tree from the (n, 2) array the registration returned (``icp.source[:, :2]`` in app.py:658-661), then recomputes the
    centroid exactly like ``Plot._update_centroid`` (trees.py:149-153).  Same error for a length mismatch."""
    new_xy = np.asarray(new_xy)
    if len(plot.trees) != new_xy.shape[0]:
        raise ValueError('Update array length does not match number of trees in the plot')
    for tree, (x, y) in zip(plot.trees, new_xy):
        tree.currentx = x
        tree.currenty = y
    if plot.trees:
        plot.current_center = np.mean(np.array([(tree.currentx, tree.currenty) for tree in plot.trees]), axis=0)
    else:
        plot.current_center = plot.center
    return plot


def apply_registration(plot, transform):
    """Moves a plot by a registration result: `transform` is the 2x3 world transform ``[A | b]`` of
    ``register_batch(...)["best_transform"][p]`` (final = A p + b for the coordinates the batch was given, i.e. the
    plot's current coordinates) - the batched counterpart of app.py:658-661."""
    t = np.asarray(transform, dtype=np.float64).reshape(2, 3)
    cur = np.array([(tree.currentx, tree.currenty) for tree in plot.trees], dtype=np.float64).reshape(-1, 2)
    return write_back(plot, cur @ t[:, :2].T + t[:, 2])

