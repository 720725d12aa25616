"""CPU oracle for the Fractional-ICP hot path.  TEST INFRASTRUCTURE ONLY.

This module is the *checker*: a numpy/scipy restatement of the algorithm in the
reference's ``ficp.py`` plus a tie-canonical variant.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import it.  Nothing under ``coregistrationgame_b200/`` does.

Parity status: PINNED.  ``tests/golden/make_golden.py`` runs the unmodified
reference (``/root/reference/ficp.py``) in the build container and stores its
inputs/outputs under ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks this
restatement against every one of those vectors (per-iteration NN indices, trimmed
subset size, FRMSD and the final aligned array).

Reference map (all ``ficp.py``):
  nn_assign*            <- find_correspondences            :65-71
  stable_order          <- get_n_first_elements / argsort  :62-63, :78
  frmsd_value           <- frmsd                           :54-60
  select_fraction*      <- find_optimal_fraction           :73-86
  fit_rigid2d*          <- compute_optimal_transform_2d    :89-110
  apply_xy              <- apply_transform_2d_xy_only      :112-119
  icp_stage             <- _iterate                        :122-147
  ficp_run              <- run                             :149-154
  pre_transform         <- trees.py Plot.rotate_plot/coordinate_flip/translate_plot :165-222

Tie rule (stricter than the reference): the nearest neighbour of a
query is the target with the smallest squared distance ``((dx*dx)+(dy*dy))[+(dz*dz)]``
(IEEE fp64, no FMA - bit-identical to what scipy's cKDTree returns) and, among exact
ties, the lowest original index.  The trimmed order is the stable order of
``(d2, source index)``.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
from scipy.spatial import cKDTree

STAGE2_LAMBDA = {2: 1.3, 3: 0.95}  # ficp.py:152


# --------------------------------------------------------------------------- inputs
def as_points(a):
    """float64 copy, must be 2-D (ficp.py:34-38)."""
    arr = np.array(a, dtype=float)
    if arr.ndim != 2:
        raise ValueError("source and target must be 2D arrays (N, D).")
    return arr


def match_dims_of(source, target):
    """ficp.py:40 - XYZ only if both sides carry a third column."""
    return 3 if (source.shape[1] >= 3 and target.shape[1] >= 3) else 2


# --------------------------------------------------------------------------- NN
def sqdist_canonical(q, t):
    """Squared distance in the canonical operation order (no FMA); q, t broadcastable (.., md)."""
    dx = q[..., 0] - t[..., 0]
    dy = q[..., 1] - t[..., 1]
    s = dx * dx + dy * dy
    if q.shape[-1] >= 3 and t.shape[-1] >= 3:
        dz = q[..., 2] - t[..., 2]
        s = s + dz * dz
    return s


def nn_assign_bruteforce(src, tgt, md, chunk=256):
    """Exact NN with lowest-index tie-break.  O(N*M); for small/medium cases.

    Returns (idx int64 (N,), d2 float64 (N,))."""
    n = src.shape[0]
    idx = np.empty(n, dtype=np.int64)
    d2 = np.empty(n, dtype=np.float64)
    t = np.ascontiguousarray(tgt[:, :md])
    for lo in range(0, n, chunk):
        q = src[lo:lo + chunk, :md]
        dx = q[:, None, 0] - t[None, :, 0]
        dy = q[:, None, 1] - t[None, :, 1]
        s = dx * dx + dy * dy
        if md == 3:
            dz = q[:, None, 2] - t[None, :, 2]
            s = s + dz * dz
        j = np.argmin(s, axis=1)  # np.argmin returns the FIRST minimum -> lowest index
        idx[lo:lo + chunk] = j
        d2[lo:lo + chunk] = s[np.arange(s.shape[0]), j]
    return idx, d2


def nn_assign_tree(src, tgt, md, tree=None):
    """Exact NN with lowest-index tie-break using a kd-tree for candidate generation.

    The kd-tree gives the NN distance; every target within that distance (closed ball,
    slightly inflated) is then re-scored in the canonical arithmetic and the lowest index
    among the exact minima is kept.  Returns (idx, d2)."""
    t = np.ascontiguousarray(tgt[:, :md])
    q = np.ascontiguousarray(src[:, :md])
    if tree is None:
        tree = cKDTree(t)
    d, j = tree.query(q, k=1)
    idx = j.astype(np.int64)
    d2 = sqdist_canonical(q, t[idx])
    # candidates that could tie: anything within d*(1+eps)
    r = d * (1.0 + 1e-12) + 1e-300
    balls = tree.query_ball_point(q, r)
    for i, cand in enumerate(balls):
        if len(cand) > 1:
            c = np.asarray(cand, dtype=np.int64)
            s = sqdist_canonical(q[i][None, :], t[c])
            m = s.min()
            best = c[s == m].min()
            idx[i] = best
            d2[i] = m
    return idx, d2


def nn_assign_reference_style(src, tgt, md):
    """What the reference does verbatim: build a kd-tree, query k=1 (ficp.py:69-70).
    Tie choice is whatever the tree traversal yields.  Returns (idx, dist)."""
    tree = cKDTree(np.ascontiguousarray(tgt[:, :md]))
    d, j = tree.query(np.ascontiguousarray(src[:, :md]), k=1)
    return j.astype(np.int64), d


# --------------------------------------------------------------------------- trimming
def stable_order(d2):
    """Trim order: ascending d2, ties by source index (stable)."""
    return np.argsort(d2, kind="stable")


def frmsd_weight(k, n, lam):
    """1 / (k/n)**lam with Python-float semantics, exactly ficp.py:60,81."""
    return 1.0 / ((k / n) ** lam)


def frmsd_weights(n, lam):
    """Table w[k-1] = frmsd_weight(k, n, lam) for k = 1..n."""
    return np.array([frmsd_weight(k, n, lam) for k in range(1, n + 1)], dtype=np.float64)


def frmsd_value(k, n, lam, sum_sq):
    """(1/(k/n)^lam) * sqrt(sum_sq / k); inf for k == 0 (ficp.py:54-60)."""
    if k == 0:
        return float("inf")
    return frmsd_weight(k, n, lam) * math.sqrt(sum_sq / k)


def select_fraction_cumsum(d2, lam, order=None):
    """FRMSD-optimal subset size from sorted squared distances (prefix sums).

    First strict minimum over k = 1..N (ficp.py:80-85).  Returns (k, value, order)."""
    n = d2.shape[0]
    if n == 0:
        return 0, float("inf"), np.empty(0, dtype=np.int64)
    if order is None:
        order = stable_order(d2)
    s = np.cumsum(d2[order])
    k_arr = np.arange(1, n + 1, dtype=np.float64)
    vals = frmsd_weights(n, lam) * np.sqrt(s / k_arr)
    k = int(np.argmin(vals)) + 1  # first minimum
    return k, float(vals[k - 1]), order


def select_fraction_pairwise(src_md, corr_md, order, lam):
    """Same as select_fraction_cumsum but re-deriving every prefix sum from the coordinate
    differences with numpy's own summation, i.e. the arithmetic of ficp.py:58-59,80-85
    (O(N^2)).  Returns (k, value)."""
    n = src_md.shape[0]
    best_v, best_k = float("inf"), 0
    for k in range(1, n + 1):
        sel = order[:k]
        diff = src_md[sel] - corr_md[sel]
        v = frmsd_weight(k, n, lam) * float(np.sqrt(np.sum(diff ** 2) / k))
        if v < best_v:
            best_v, best_k = v, k
    return best_k, best_v


def fixed_fraction_k(n, frac):
    """Extension (not in the reference): subset size for a fixed trim fraction."""
    return max(1, min(n, int(math.floor(frac * n + 1e-9))))


# --------------------------------------------------------------------------- rigid fit
def fit_rigid2d_svd(src_xy, tgt_xy, allow_reflection=False):
    """Kabsch in the plane via SVD (ficp.py:89-110).  Returns 3x3 homogeneous T."""
    mu_s = src_xy.mean(axis=0)
    mu_t = tgt_xy.mean(axis=0)
    h = (src_xy - mu_s).T @ (tgt_xy - mu_t)
    u, _, vt = np.linalg.svd(h)
    r = vt.T @ u.T
    if (not allow_reflection) and np.linalg.det(r) < 0:
        vt[-1, :] *= -1
        r = vt.T @ u.T
    out = np.eye(3)
    out[:2, :2] = r
    out[:2, 2] = mu_t - mu_s @ r.T
    return out


def fit_rigid2d_closed(src_xy, tgt_xy, allow_reflection=False):
    """Same optimum without SVD: the normalised (H00+H11, H01-H10) pair.
    H == 0 gives the identity rotation."""
    mu_s = src_xy.mean(axis=0)
    mu_t = tgt_xy.mean(axis=0)
    xc = src_xy - mu_s
    yc = tgt_xy - mu_t
    h00 = float(np.sum(xc[:, 0] * yc[:, 0]))
    h01 = float(np.sum(xc[:, 0] * yc[:, 1]))
    h10 = float(np.sum(xc[:, 1] * yc[:, 0]))
    h11 = float(np.sum(xc[:, 1] * yc[:, 1]))
    # reflection only when det(H) is negative beyond rounding noise (for det(H) == 0 the SVD's pick is arbitrary)
    if allow_reflection and (h00 * h11 - h01 * h10) < -1e-14 * (abs(h00 * h11) + abs(h01 * h10)):
        a, b = h00 - h11, h01 + h10
        nrm = math.hypot(a, b)
        c, s = (1.0, 0.0) if nrm == 0 else (a / nrm, b / nrm)
        r = np.array([[c, s], [s, -c]])
    else:
        a, b = h00 + h11, h01 - h10
        nrm = math.hypot(a, b)
        c, s = (1.0, 0.0) if nrm == 0 else (a / nrm, b / nrm)
        r = np.array([[c, -s], [s, c]])
    out = np.eye(3)
    out[:2, :2] = r
    out[:2, 2] = mu_t - r @ mu_s
    return out


def apply_xy(points, t):
    """Move XY by the homogeneous 3x3, leave every other column bit-identical (ficp.py:112-119)."""
    out = points.copy()
    x = points[:, 0]
    y = points[:, 1]
    out[:, 0] = t[0, 0] * x + t[0, 1] * y + t[0, 2]
    out[:, 1] = t[1, 0] * x + t[1, 1] * y + t[1, 2]
    return out


# --------------------------------------------------------------------------- ICP loop
@dataclass
class IterRecord:
    idx: np.ndarray       # NN index per source point
    d2: np.ndarray        # squared NN distance
    k: int                # trimmed subset size
    value: float          # FRMSD at k
    inliers: np.ndarray   # sorted source indices of the trimmed subset


class TimeBudgetExceeded(Exception):
    """Raised between passes when a RunTrace carries a wall-clock deadline (bounded CPU baseline samples)."""


@dataclass
class RunTrace:
    records: list = field(default_factory=list)   # one per NN pass (= hypothesis-iteration)
    deadline: float = None                        # time.perf_counter() value after which no new pass starts
    light: bool = False                           # keep only counters, not the per-pass arrays
    transform: np.ndarray = field(default_factory=lambda: np.eye(3))  # composed T_total
    stage_passes: list = field(default_factory=list)

    @property
    def passes(self):
        return len(self.records)


def _pass(src, tgt, md, lam, nn, tree, fixed_k, trace, pairwise):
    if trace is not None and trace.deadline is not None:
        import time
        if time.perf_counter() > trace.deadline:
            raise TimeBudgetExceeded()
    if nn == "tree":
        idx, d2 = nn_assign_tree(src, tgt, md, tree)
    elif nn == "brute":
        idx, d2 = nn_assign_bruteforce(src, tgt, md)
    elif nn == "reference":
        idx, d = nn_assign_reference_style(src, tgt, md)
        d2 = sqdist_canonical(src[:, :md], tgt[idx, :md])
    else:
        raise ValueError(nn)
    n = src.shape[0]
    if fixed_k is not None:
        order = stable_order(d2)
        k = fixed_k
        value = frmsd_value(k, n, lam, float(np.cumsum(d2[order])[k - 1]))
    elif pairwise:
        order = stable_order(d2)
        k, value = select_fraction_pairwise(src[:, :md], tgt[idx, :md], order, lam)
    else:
        k, value, order = select_fraction_cumsum(d2, lam)
    inl = order[:k]
    if trace is not None:
        trace.records.append(None if trace.light else IterRecord(idx.copy(), d2.copy(), k, value, np.sort(inl)))
    return idx, inl, k, value


def icp_stage(src, tgt, md, lam, threshold=1e-6, max_iterations=1000, allow_reflection=False,
              nn="tree", tree=None, fixed_k=None, trace=None, pairwise=False, closed_form=False):
    """One stage of the loop (ficp.py:122-147).  Returns the moved source array."""
    if src.shape[0] == 0 or tgt.shape[0] == 0:
        return src
    fit = fit_rigid2d_closed if closed_form else fit_rigid2d_svd
    n0 = trace.passes if trace is not None else 0
    idx, inl, k, cur = _pass(src, tgt, md, lam, nn, tree, fixed_k, trace, pairwise)
    if k == 0:
        return src
    it = 0
    while it < max_iterations:
        t = fit(src[inl, :2], tgt[idx[inl], :2], allow_reflection)
        src = apply_xy(src, t)
        if trace is not None:
            trace.transform = t @ trace.transform
        idx, inl, k, new = _pass(src, tgt, md, lam, nn, tree, fixed_k, trace, pairwise)
        if cur - new <= threshold:
            break
        cur = new
        it += 1
    if trace is not None:
        trace.stage_passes.append(trace.passes - n0)
    return src


def ficp_run(source, target, lambda_val=3.0, threshold=1e-6, max_iterations=1000,
             allow_reflection=False, nn="tree", hoist_tree=True, fixed_frac=None,
             trace=None, pairwise=False, closed_form=False, stage2_lambda=None):
    """Two-stage Fractional ICP (ficp.py:149-154).  Returns the aligned (N, D) array.

    nn="reference" + hoist_tree=False reproduces the reference's cost profile (kd-tree rebuilt
    on every pass).  nn="tree"/"brute" use the lowest-index tie rule."""
    src = as_points(source)
    tgt = as_points(target)
    md = match_dims_of(src, tgt)
    tree = None
    if nn == "tree" and hoist_tree and tgt.shape[0] > 0:
        tree = cKDTree(np.ascontiguousarray(tgt[:, :md]))
    fixed_k = fixed_fraction_k(src.shape[0], fixed_frac) if (fixed_frac is not None and src.shape[0]) else None
    lam2 = STAGE2_LAMBDA[md] if stage2_lambda is None else stage2_lambda
    for lam in (lambda_val, lam2):
        src = icp_stage(src, tgt, md, lam, threshold, max_iterations, allow_reflection,
                        nn=nn, tree=tree, fixed_k=fixed_k, trace=trace, pairwise=pairwise,
                        closed_form=closed_form)
    return src


# --------------------------------------------------------------------------- hypotheses
def hypothesis_matrix(theta_deg, flip):
    """2x2 linear part of a start-pose hypothesis: rotate CCW by theta after an optional
    y-flip, both about the plot centroid (trees.py:165-222: x' = R(theta) F^f (x-c) + c + d)."""
    th = np.radians(theta_deg)
    c, s = np.cos(th), np.sin(th)
    if flip:
        return np.array([[c, s], [s, -c]])      # R @ diag(1,-1)
    return np.array([[c, -s], [s, c]])


def hypothesis_table(n_rot, flips=(0, 1), translations=((0.0, 0.0),)):
    """(H, 6) table [m00 m01 m10 m11 dx dy]; order: translation-major, then flip, then rotation."""
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


def pre_transform(src, hyp_row, centre):
    """Start pose of one hypothesis, elementwise in THIS operation order (the device kernel
    evaluates the same expression without FMA, so both sides start from identical bits):
        u = p - c ;  x' = (m00*ux + m01*uy) + (cx + dx) ;  y' = (m10*ux + m11*uy) + (cy + dy)."""
    m00, m01, m10, m11, dx, dy = [float(v) for v in hyp_row]
    out = src.copy()
    ux = src[:, 0] - centre[0]
    uy = src[:, 1] - centre[1]
    ox = centre[0] + dx
    oy = centre[1] + dy
    out[:, 0] = (m00 * ux + m01 * uy) + ox
    out[:, 1] = (m10 * ux + m11 * uy) + oy
    return out


def pack_best_key(score, hyp_id):
    """uint64 ranking key: fp32 bit pattern of the (non-negative) score in the high word,
    hypothesis id in the low word; min() over keys = best score, ties to the lowest id."""
    bits = np.float32(score).view(np.uint32)
    return (np.uint64(bits) << np.uint64(32)) | np.uint64(hyp_id)


def run_hypotheses(source, target, hyp_table, centre=None, min_k=3, trace_all=False, **kw):
    """Oracle for the batched search over start poses of ONE plot.

    Each hypothesis h equals ficp_run(pre_transform(source, hyp_table[h], centre), target).
    Score = final FRMSD (stage-2 lambda); hypotheses ending with k < min_k are disqualified.
    Returns dict with per-hypothesis aligned arrays, score, k, passes, and the winner."""
    src = as_points(source)
    tgt = as_points(target)
    if centre is None:
        centre = src[:, :2].mean(axis=0)
    md = match_dims_of(src, tgt)
    tree = cKDTree(np.ascontiguousarray(tgt[:, :md])) if tgt.shape[0] else None
    res = {"aligned": [], "score": [], "k": [], "passes": [], "traces": []}
    best_key = None
    for h in range(hyp_table.shape[0]):
        s0 = pre_transform(src, hyp_table[h], centre)
        tr = RunTrace()
        out = s0
        lam2 = STAGE2_LAMBDA[md]
        fk = kw.get("fixed_frac")
        fixed_k = fixed_fraction_k(src.shape[0], fk) if fk is not None else None
        for lam in (kw.get("lambda_val", 3.0), lam2):
            out = icp_stage(out, tgt, md, lam, kw.get("threshold", 1e-6), kw.get("max_iterations", 1000),
                            kw.get("allow_reflection", False), nn="tree", tree=tree, fixed_k=fixed_k, trace=tr,
                            closed_form=kw.get("closed_form", False))
        last = tr.records[-1]
        score = last.value if last.k >= min_k else float("inf")
        res["aligned"].append(out)
        res["score"].append(score)
        res["k"].append(last.k)
        res["passes"].append(tr.passes)
        if trace_all:
            res["traces"].append(tr)
        key = pack_best_key(score, h)
        if best_key is None or key < best_key:
            best_key = key
    res["best_key"] = best_key
    res["best_hyp"] = int(best_key & np.uint64(0xFFFFFFFF))
    return res


# --------------------------------------------------------------------------- synthetic scenes
# The generator lives in the package (bench.py's CUDA arm must not import the oracle); re-exported for the tests.
from coregistrationgame_b200.synthetic import synthetic_scene  # noqa: E402,F401


# --------------------------------------------------------------------------- after the ICP
def remove_matches_oracle(plot_trees, chm, min_dist_percent=15):
    """CHMPlot.remove_matches (chm_plot.py:223-285) on arrays.  plot_trees (n,3) / chm (M,3): x, y, height.
    Returns matched (n,) int64: index (into the ORIGINAL chm rows) removed by each tree, or -1."""
    plot_trees = np.asarray(plot_trees, dtype=float).reshape(-1, 3)
    chm = np.asarray(chm, dtype=float).reshape(-1, 3)
    use_3d = bool(np.isfinite(plot_trees[:, 2]).all() and np.isfinite(chm[:, 2]).all())
    md = 3 if use_3d else 2
    alive = np.arange(len(chm))
    matched = np.full(len(plot_trees), -1, dtype=np.int64)
    for t, row in enumerate(plot_trees):
        if len(alive) == 0:
            break
        d = np.sqrt(sqdist_canonical(row[None, :md], chm[alive, :md]))
        j = int(np.argmin(d))                      # first minimum = lowest remaining index
        h = row[2] if (use_3d or np.isfinite(row[2])) else 10.0
        if d[j] < (min_dist_percent / 100.0) * float(h):
            matched[t] = alive[j]
            alive = np.delete(alive, j)
    return matched


def transform_record_oracle(original_xy, current_xy, flipped=False):
    """Plot.get_transform (trees.py:248-280) + the record of App.store_transformations (app.py:901-912)."""
    T = fit_rigid2d_svd(np.asarray(original_xy, dtype=float)[:, :2], np.asarray(current_xy, dtype=float)[:, :2], bool(flipped))
    return {"tx": float(T[0, 2]), "ty": float(T[1, 2]), "r00": float(T[0, 0]), "r01": float(T[0, 1]),
            "r10": float(T[1, 0]), "r11": float(T[1, 1]), "flip": bool(flipped)}
