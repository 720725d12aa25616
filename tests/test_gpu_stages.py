"""GPU parity tests of the stage kernels (through the C ABI / FractionalICP methods) against the
CPU oracle.  Integer/index results must be bit-exact; fp64 values are compared at the tolerances
written next to each check."""
import numpy as np
import pytest

from oracle import ficp_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    from coregistrationgame_b200 import _lib
    _lib.require_device()
    return _lib


def _scene(m, n, seed, dims, **kw):
    tgt, plots, _ = orc.synthetic_scene(m, n, seed=seed, dims=dims, **kw)
    return tgt, plots[0]


# ------------------------------------------------------------------------------- kernel 1a/1b
@pytest.mark.parametrize("dims", [2, 3])
@pytest.mark.parametrize("offset", [(0.0, 0.0), (420000.0, 6483000.0)])
def test_nn_query_bit_exact(gpu, dims, offset):
    from coregistrationgame_b200 import TargetIndex
    tgt, src = _scene(20000, 400, seed=21, dims=dims, dup_every=9, lattice_patch=7)
    rng = np.random.default_rng(5)
    lo, hi = tgt[:, :2].min(0), tgt[:, :2].max(0)
    q = np.empty((3000, dims))
    q[:, :2] = rng.uniform(lo - 0.3 * (hi - lo), hi + 0.3 * (hi - lo), (3000, 2))  # also outside the grid
    if dims == 3:
        q[:, 2] = rng.uniform(0, 40, 3000)
    q[:400] = src
    # exact ties: centres of lattice cells (4 equidistant targets) and exact duplicates
    q[400:436, 0] = tgt[:36, 0] + 0.5
    q[400:436, 1] = tgt[:36, 1] + 0.5
    q[436:500] = tgt[-64:]
    off = np.zeros(dims)
    off[:2] = offset
    tgt, q = tgt + off, q + off
    ti = TargetIndex(tgt)
    idx, dist = ti.query(q)
    ref_idx, ref_d2 = orc.nn_assign_bruteforce(q, tgt, dims)
    np.testing.assert_array_equal(idx, ref_idx)                 # bit-exact indices incl. lowest-index ties
    np.testing.assert_array_equal(dist, np.sqrt(ref_d2))         # identical fp64 bits
    info = ti.info()
    assert info["m"] == len(tgt) and info["grid_w"] * info["grid_h"] >= 1
    ti.close()


@pytest.mark.parametrize("dims", [2, 3])
def test_nn_query_bulk_kernel_bit_exact(gpu, dims):
    """The bulk kernel (cell-ordered queries, windows of cells staged in shared memory by cp.async.bulk) returns the bits
    of the thread-per-query kernel and of brute force: lattice ties, exact duplicates, off-grid queries, UTM offsets."""
    from coregistrationgame_b200 import TargetIndex
    tgt, src = _scene(20000, 400, seed=21, dims=dims, dup_every=9, lattice_patch=7)
    rng = np.random.default_rng(6)
    lo, hi = tgt[:, :2].min(0), tgt[:, :2].max(0)
    nq = 70000
    q = np.empty((nq, dims))
    q[:, :2] = rng.uniform(lo - 0.05 * (hi - lo), hi + 0.05 * (hi - lo), (nq, 2))
    q[:2000, :2] = rng.uniform(lo - 0.6 * (hi - lo), hi + 0.6 * (hi - lo), (2000, 2))   # far outside the grid
    if dims == 3:
        q[:, 2] = rng.uniform(0, 40, nq)
        q[5000:9000, 2] = rng.uniform(-80, 150, 4000)          # wide searches: rings >= 2 on the global grid
    q[2000:2400] = src
    q[2400:2436, 0] = tgt[:36, 0] + 0.5                        # centres of lattice cells: 4 equidistant targets
    q[2400:2436, 1] = tgt[:36, 1] + 0.5
    q[2436:2500] = tgt[-64:]                                   # exact duplicates of targets
    q[2500:3000, :2] = tgt[rng.integers(0, 49, 500), :2] + rng.integers(-2, 3, (500, 2)) * 0.5   # lattice edge / corner ties
    off = np.zeros(dims)
    off[:2] = (420000.0, 6483000.0)
    tgt, q = tgt + off, q + off
    for purpose in ("query", "icp"):
        ti = TargetIndex(tgt, purpose=purpose)
        cnt = {}
        idx_b, dist_b = ti.query(q, kernel="bulk", counters=cnt)
        idx_t, dist_t = ti.query(q, kernel="thread")
        np.testing.assert_array_equal(idx_b, idx_t)
        np.testing.assert_array_equal(dist_b, dist_t)
        assert cnt["window"] + cnt["global_grid"] == nq, cnt
        assert cnt["window"] > 0.9 * nq, cnt                    # dense batch: resolved from shared-memory windows
        sel = np.r_[0:6000, rng.integers(0, nq, 4000)]
        ref_idx, ref_d2 = orc.nn_assign_bruteforce(q[sel], tgt, dims)
        np.testing.assert_array_equal(idx_b[sel], ref_idx)
        np.testing.assert_array_equal(dist_b[sel], np.sqrt(ref_d2))
        idx_a, dist_a = ti.query(q)                             # auto: bulk kernel on the dense (ICP) grid, thread kernel otherwise
        np.testing.assert_array_equal(idx_a, idx_b)
        ti.close()


def test_nn_query_bulk_kernel_sparse_clustered_and_degenerate(gpu):
    """Query batches the windows do not fit: sparse (chunks spanning many rows -> global grid, still exact), clustered
    (thousands of queries in a handful of cells), one-row / one-cell grids, and a batch that is not a multiple of 256."""
    from coregistrationgame_b200 import TargetIndex
    rng = np.random.default_rng(8)
    tgt, _ = _scene(60000, 50, seed=5, dims=3)
    ti = TargetIndex(tgt, purpose="query")
    lo, hi = tgt[:, :2].min(0), tgt[:, :2].max(0)
    # sparse: far fewer queries than cells
    q = np.column_stack([rng.uniform(lo[0], hi[0], 3001), rng.uniform(lo[1], hi[1], 3001), rng.uniform(5, 35, 3001)])
    cnt = {}
    ib, db = ti.query(q, kernel="bulk", counters=cnt)
    it, dt = ti.query(q, kernel="thread")
    np.testing.assert_array_equal(ib, it)
    np.testing.assert_array_equal(db, dt)
    assert cnt["global_grid"] > 0
    # clustered: 66 000 queries inside a 30 m square
    c = 0.5 * (lo + hi)
    q = np.column_stack([rng.uniform(c[0] - 15, c[0] + 15, 66000), rng.uniform(c[1] - 15, c[1] + 15, 66000), rng.uniform(5, 35, 66000)])
    q[::7] = q[3]                                               # thousands of identical queries
    cnt = {}
    ib, db = ti.query(q, kernel="bulk", counters=cnt)
    it, dt = ti.query(q, kernel="thread")
    np.testing.assert_array_equal(ib, it)
    np.testing.assert_array_equal(db, dt)
    assert cnt["window"] == 66000, cnt
    ti.close()
    # degenerate grids: a single point, all-identical points, collinear points (one row / one column of cells)
    q2 = rng.normal(size=(70001, 2)) * 30
    for t2 in (np.array([[1.0, 2.0]]), np.repeat(np.array([[3.0, -1.0]]), 50, axis=0),
               np.stack([np.linspace(-100, 100, 3000), np.zeros(3000)], 1),
               np.stack([np.zeros(3000), np.linspace(-100, 100, 3000)], 1)):
        ti = TargetIndex(t2, purpose="query")
        ib, db = ti.query(q2, kernel="bulk")
        ref_idx, ref_d2 = orc.nn_assign_bruteforce(q2[:4000], t2, 2)
        np.testing.assert_array_equal(ib[:4000], ref_idx)
        np.testing.assert_array_equal(db[:4000], np.sqrt(ref_d2))
        it, dt = ti.query(q2, kernel="thread")
        np.testing.assert_array_equal(ib, it)
        np.testing.assert_array_equal(db, dt)
        ti.close()


def test_nn_query_degenerate_targets(gpu):
    from coregistrationgame_b200 import TargetIndex
    rng = np.random.default_rng(0)
    q = rng.normal(size=(100, 2)) * 5
    for tgt in (np.array([[1.0, 2.0]]),                                 # single point
                np.repeat(np.array([[3.0, -1.0]]), 50, axis=0),          # all identical
                np.stack([np.linspace(0, 100, 300), np.zeros(300)], 1),  # collinear
                np.stack([np.zeros(300), np.linspace(0, 100, 300)], 1)):
        ti = TargetIndex(tgt)
        idx, dist = ti.query(q)
        ref_idx, ref_d2 = orc.nn_assign_bruteforce(q, tgt, 2)
        np.testing.assert_array_equal(idx, ref_idx)
        np.testing.assert_array_equal(dist, np.sqrt(ref_d2))
        ti.close()


def _skewed_targets(dims):
    """Targets the mean-density grid of round 1 could not take: one stray coordinate, long tails, clusters
    far apart, heavy duplicates."""
    rng = np.random.default_rng(77)
    utm = np.array([420000.0, 6483000.0])
    core = np.column_stack([rng.uniform(0, 700, 30000), rng.uniform(0, 500, 30000)]) + utm
    z = lambda n: rng.uniform(5, 35, (n, 1))
    out = {}
    t = core.copy()
    t[17] = [0.0, 0.0]                                        # the (0, 0) placeholder row of a UTM export
    t[4242] = utm + [9.0e4, -3.0e4]                           # and a stray point 90 km away
    out["placeholder_row"] = t
    out["gaussian_tails"] = rng.normal(0.0, 60.0, (40000, 2)) * [1.0, 0.2] + utm
    two = np.vstack([core[:15000], core[15000:] + [250000.0, 40000.0]])
    out["two_stands_250km_apart"] = two
    dup = np.vstack([core[:20000], np.repeat(core[:1], 2500, axis=0), np.repeat(core[5:6] + 0.25, 40, axis=0)])
    out["heavy_duplicates"] = dup                              # one cell of > 2500 points, one of ~40
    out["all_identical_above_heavy_cap"] = np.repeat(core[:1], 6000, axis=0)
    if dims == 3:
        out = {k: np.hstack([v, z(len(v))]) for k, v in out.items()}
    return out


@pytest.mark.parametrize("dims", [2, 3])
def test_nn_query_skewed_targets_are_exact_and_fast_to_build(gpu, dims):
    """Robust grid extent + clamped border cells: exact neighbours (bit-exact indices and distances) for queries inside
    the core, next to the far-off points, off the map; the heaviest cell stays small where the skew is an outlier."""
    from coregistrationgame_b200 import TargetIndex
    rng = np.random.default_rng(3)
    for name, tgt in _skewed_targets(dims).items():
        ti = TargetIndex(tgt)
        info = ti.info()
        lo, hi = np.percentile(tgt[:, :2], 1, axis=0), np.percentile(tgt[:, :2], 99, axis=0)
        q = np.empty((1500, dims))
        q[:, :2] = rng.uniform(lo - 0.2 * (hi - lo + 1), hi + 0.2 * (hi - lo + 1), (1500, 2))
        far = tgt[np.argsort(np.abs(tgt[:, :2] - np.median(tgt[:, :2], axis=0)).sum(1))[-200:], :2]
        q[:200, :2] = far + rng.normal(0, 3.0, (200, 2))                      # next to the most remote target points
        q[200:260, :2] = rng.uniform(-1e6, 1e7, (60, 2))                      # anywhere
        if dims == 3:
            q[:, 2] = rng.uniform(0, 40, 1500)
        idx, dist = ti.query(q)
        ref_idx, ref_d2 = orc.nn_assign_bruteforce(q, tgt, dims)
        np.testing.assert_array_equal(idx, ref_idx, err_msg=name)
        np.testing.assert_array_equal(dist, np.sqrt(ref_d2), err_msg=name)
        if name in ("placeholder_row", "two_stands_250km_apart"):
            assert info["clamped"] if name == "placeholder_row" else True
        if name == "placeholder_row":
            assert info["max_cell_pts"] <= 64, info                         # round 1: ~all 30 000 points in one cell
        assert info["build_ms"] < 50.0, (name, info)                          # round 1: O(c^2) insertion sort per heavy cell
        ti.close()


def test_icp_on_a_skewed_target_matches_oracle(gpu):
    """The whole ICP against a target with a placeholder row and a stray point: pass counts, k and poses as the oracle."""
    from coregistrationgame_b200 import IcpBatch, TargetIndex
    from coregistrationgame_b200.batch import compose_world_transform
    tgt, plots, _ = orc.synthetic_scene(30000, 150, seed=41, dims=3, hidden_pose=True, out_frac=0.1)
    tgt = tgt.copy()
    tgt[:, :2] += [420000.0, 6483000.0]
    src = plots[0].copy()
    src[:, :2] += [420000.0, 6483000.0]
    tgt[11] = [0.0, 0.0, 10.0]
    tgt[12] = [420000.0 + 5.0e4, 6483000.0, 12.0]
    hyp = orc.hypothesis_table(6, flips=(0, 1))
    ti = TargetIndex(tgt)
    assert ti.info()["clamped"]
    for cta in (False, True):
        b = IcpBatch(ti, [src], hyp, cta_per_icp=cta)
        out = b.run().results()
        ref = orc.run_hypotheses(src, tgt, hyp, centre=b.centres[0], min_k=3, closed_form=True)
        np.testing.assert_array_equal(out["hyp"]["passes"][0], ref["passes"])
        np.testing.assert_array_equal(out["hyp"]["k"][0], ref["k"])
        for h in range(hyp.shape[0]):
            A = compose_world_transform(out["hyp"][0, h], b.centres[0])
            np.testing.assert_allclose(src[:, :2] @ A[:, :2].T + A[:, 2], ref["aligned"][h][:, :2], rtol=0, atol=1e-6)
        assert int(out["best_hyp"][0]) == ref["best_hyp"]
        b.close()
    ti.close()


def test_nonfinite_inputs_raise_value_error(gpu):
    from coregistrationgame_b200 import TargetIndex
    from ficp import FractionalICP
    bad = np.array([[0.0, 1.0], [np.nan, 2.0]])
    with pytest.raises(ValueError):
        TargetIndex(bad)
    good = np.random.default_rng(1).normal(size=(10, 2))
    with pytest.raises(ValueError):
        FractionalICP(bad, good).run()
    with pytest.raises(ValueError):
        FractionalICP(good, bad).run()
    with pytest.raises(ValueError):
        FractionalICP(np.zeros(3), good)


# ------------------------------------------------------------------------------- kernel 2
@pytest.mark.parametrize("n", [1, 2, 31, 150, 500, 1024, 3000])
@pytest.mark.parametrize("lam", [3.0, 1.3, 0.95])
def test_select_fraction_matches_oracle(gpu, n, lam):
    from ficp import FractionalICP
    rng = np.random.default_rng(n)
    src = rng.normal(size=(n, 3)) * 10
    corr = src + rng.normal(size=(n, 3)) * rng.choice([0.05, 0.5, 5.0], size=(n, 1))
    if n > 10:
        src[3], corr[3] = src[7], corr[7]        # duplicated tree -> exact tie in the trim order (index decides)
    icp = FractionalICP(src, corr, lambda_val=lam)
    d2 = orc.sqdist_canonical(src, corr)
    dist = np.sqrt(d2)
    frac, k = icp.find_optimal_fraction(corr, dist)
    order = orc.stable_order(dist)
    k_ref, val_ref = orc.select_fraction_pairwise(src, corr, order, lam)
    assert k == k_ref
    assert frac == k_ref / n
    np.testing.assert_array_equal(icp.get_n_first_elements(k, dist), order[:k])
    val = icp.frmsd(frac, k, src[order[:k]], corr[order[:k]])
    assert val == pytest.approx(val_ref, rel=1e-13)


@pytest.mark.parametrize("n", [8193, 20000, 70001])
@pytest.mark.parametrize("lam", [3.0, 0.95])
def test_select_fraction_large_plots(gpu, n, lam):
    """Plots above the one-CTA kernel's 8192 rows (the reference accepts any N, ficp.py:73-86): global-scratch bitonic sort +
    scan.  The trim order is exact (array_equal with the stable (distance, index) order, ties included); k is the first
    strict minimum of the oracle's cumsum FRMSD curve unless that minimum is within summation rounding of another k."""
    from ficp import FractionalICP
    rng = np.random.default_rng(n)
    src = rng.normal(size=(n, 3)) * 50
    corr = src + rng.normal(size=(n, 3)) * rng.choice([0.05, 0.5, 5.0], size=(n, 1))
    src[3], corr[3] = src[7], corr[7]
    src[n - 1], corr[n - 1] = src[11], corr[11]      # exact ties far apart in index
    icp = FractionalICP(src, corr, lambda_val=lam)
    d2 = orc.sqdist_canonical(src, corr)
    dist = np.sqrt(d2)
    order = orc.stable_order(dist)
    np.testing.assert_array_equal(icp.get_n_first_elements(n, dist), order)
    frac, k = icp.find_optimal_fraction(corr, dist)
    k_ref, val_ref, _ = orc.select_fraction_cumsum(d2, lam, order)
    s = np.cumsum(d2[order])
    vals = orc.frmsd_weights(n, lam) * np.sqrt(s / np.arange(1, n + 1))
    assert k == k_ref or abs(vals[k - 1] - val_ref) <= 1e-12 * val_ref, (k, k_ref, vals[k - 1], val_ref)
    assert frac == k / n
    # fixed-size trimming reads the k-th element of the same order
    lib_k, lib_f = ctypes_select_fixed(src, corr, dist, n // 2, lam)
    assert lib_k == n // 2
    assert lib_f == pytest.approx(vals[n // 2 - 1], rel=1e-12)


def ctypes_select_fixed(src, corr, dist, fixed_k, lam):
    import ctypes as C
    from coregistrationgame_b200 import _lib
    from coregistrationgame_b200.ficp import frmsd_weights
    n = len(src)
    w = frmsd_weights(n, lam)
    k, val = C.c_int64(0), C.c_double(0.0)
    d = np.ascontiguousarray(dist, dtype=np.float64)
    s3, c3 = np.ascontiguousarray(src), np.ascontiguousarray(corr)
    _lib.check(_lib.load().ficp_select_fraction(_lib.ptr(s3), 3, _lib.ptr(c3), 3, _lib.ptr(d), n, 3, _lib.ptr(w), fixed_k,
                                                C.byref(k), C.byref(val), None), "ficp_select_fraction")
    return int(k.value), float(val.value)


def test_run_on_a_plot_above_the_persistent_kernels_limit(gpu):
    """A 12 000-tree plot (above the persistent kernel's 1024 and the one-CTA trim kernel's 8192): the host-stepped path
    over the stage kernels recovers the hidden pose like the oracle's run on the same data."""
    from ficp import FractionalICP
    rng = np.random.default_rng(5)
    m, n = 60000, 12000
    tgt = np.column_stack([rng.uniform(0, 800, m), rng.uniform(0, 800, m), rng.uniform(5, 35, m)])
    pick = rng.choice(m, n, replace=False)
    src = tgt[pick] + np.column_stack([rng.normal(0, 0.05, (n, 2)), rng.normal(0, 0.3, n)])
    th = np.deg2rad(0.2)
    r = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
    c = src[:, :2].mean(0)
    moved = src.copy()
    moved[:, :2] = (src[:, :2] - c) @ r.T + c + np.array([0.3, -0.2])
    icp = FractionalICP(moved, tgt, max_iterations=30)
    out = icp.run()
    tr = orc.RunTrace(light=True)
    ref = orc.ficp_run(moved, tgt, max_iterations=30, trace=tr)
    assert icp.n_passes_ == tr.passes
    np.testing.assert_allclose(out[:, :2], ref[:, :2], atol=1e-6)
    assert np.abs(out[:, :2] - src[:, :2]).max() < 0.05
    np.testing.assert_array_equal(out[:, 2], moved[:, 2])


def test_select_fraction_zero_and_tied_distances(gpu):
    from ficp import FractionalICP
    src = np.arange(40, dtype=float).reshape(20, 2)
    icp = FractionalICP(src, src.copy())
    frac, k = icp.find_optimal_fraction(src.copy(), np.zeros(20))
    assert (frac, k) == (1 / 20, 1)             # all FRMSD(k) == 0 -> first strict minimum is k = 1
    np.testing.assert_array_equal(icp.get_n_first_elements(20, np.zeros(20)), np.arange(20))


# ------------------------------------------------------------------------------- kernel 3
@pytest.mark.parametrize("refl", [False, True])
def test_fit_rigid2d_matches_svd(gpu, refl):
    from ficp import FractionalICP
    rng = np.random.default_rng(3)
    for k in (1, 2, 3, 17, 500, 5000):
        a = rng.normal(size=(k, 3)) * 30 + 1000.0
        th = rng.uniform(-np.pi, np.pi)
        r = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
        if refl and k > 2:
            r = r @ np.diag([1.0, -1.0])
        b = a.copy()
        b[:, :2] = a[:, :2] @ r.T + rng.normal(size=2) * 20 + rng.normal(size=(k, 2)) * 0.1
        icp = FractionalICP(a, b, allow_reflection=refl)
        T = icp.compute_optimal_transform_2d(a, b)
        if refl and k <= 2:
            continue  # det(H) == 0: the SVD's reflection choice is arbitrary
        T_ref = orc.fit_rigid2d_svd(a[:, :2], b[:, :2], refl)
        np.testing.assert_allclose(T[:2, :2], T_ref[:2, :2], atol=1e-11)     # rotation: << 1e-6 rad
        np.testing.assert_allclose(T[:2, 2], T_ref[:2, 2], atol=1e-8)
        np.testing.assert_allclose(T[:2, :2].T @ T[:2, :2], np.eye(2), atol=1e-14)
        np.testing.assert_array_equal(T[2], [0.0, 0.0, 1.0])


def test_apply_xy_only_bit_exact(gpu):
    from ficp import FractionalICP
    rng = np.random.default_rng(8)
    pts = rng.normal(size=(1000, 5)) * 100
    T = orc.fit_rigid2d_closed(rng.normal(size=(5, 2)), rng.normal(size=(5, 2)))
    icp = FractionalICP(pts, pts)
    out = icp.apply_transform_2d_xy_only(pts, T)
    ref = orc.apply_xy(pts, T)
    np.testing.assert_array_equal(out, ref)           # same operation order, no FMA
    np.testing.assert_array_equal(out[:, 2:], pts[:, 2:])
    assert out is not pts


# ------------------------------------------------------------------------------- reference test shapes
def _cloud(n, seed):
    rng = np.random.default_rng(seed)
    xy = rng.normal(size=(n, 2)) @ np.array([[1.0, 0.3], [0.0, 0.6]]).T
    z = np.linspace(0.0, 20.0, n)[:, None] + rng.normal(scale=0.02, size=(n, 1))
    return np.hstack([xy, z])


def _move(src, deg, t):
    th = np.deg2rad(deg)
    r = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
    return np.hstack([src[:, :2] @ r.T + np.asarray(t), src[:, 2:]])


def _nn_rmsd(a, b):
    _, d2 = orc.nn_assign_tree(a, b, a.shape[1])
    return np.sqrt(d2.mean())


def test_reference_acceptance_pose_recovery(gpu):
    """Same properties the reference's tests/test_ficp.py:39-101 assert, on the same cloud shapes."""
    from ficp import FractionalICP
    src = _cloud(150, 1)
    tgt = _move(src, 27.0, [1.6, -2.2])
    icp = FractionalICP(src.copy(), tgt)
    out = icp.run()
    np.testing.assert_array_equal(out[:, 2], src[:, 2])
    ang = np.rad2deg(np.arctan2(icp.transform_[1, 0], icp.transform_[0, 0]))
    assert abs(((ang - 27.0 + 180) % 360) - 180) < 0.2
    assert _nn_rmsd(out, tgt) < 2e-3
    assert icp.lambda_val == 0.95

    src = _cloud(200, 2)
    full = _move(src, 31.0, [2.5, -1.8])
    keep = np.random.default_rng(123).choice(200, 100, replace=False)
    out = FractionalICP(src.copy(), full[keep]).run()
    assert _nn_rmsd(out, full[keep]) < 0.4 * _nn_rmsd(src, full[keep])

    src = _cloud(200, 3)
    clean = _move(src, -22.0, [-1.2, 2.0])
    rng = np.random.default_rng(7)
    t = clean[rng.choice(200, 100, replace=False)]
    no = int(0.3 * len(t))
    t = np.vstack([t, np.hstack([rng.uniform(-20, 20, (no, 2)), rng.uniform(-5, 25, (no, 1))])])
    out = FractionalICP(src.copy(), t).run()
    assert _nn_rmsd(out, t) < 0.5 * _nn_rmsd(src, t)
    _, d2 = orc.nn_assign_tree(out[:, :2], clean[:, :2], 2)
    assert np.mean(np.sqrt(d2) < 0.12) > 0.90


def test_empty_inputs_follow_reference_conventions(gpu):
    """tests/test_ficp.py:104-126 of the reference."""
    from ficp import FractionalICP
    tgt = _cloud(5, 42)
    out = FractionalICP(np.empty((0, 3)), tgt).run()
    assert out.shape == (0, 3)
    src = _cloud(4, 24)
    icp = FractionalICP(src.copy(), np.empty((0, 3)))
    corr, dist = icp.find_correspondences(src, np.empty((0, 3)))
    assert corr.shape == (0, 3) and dist.size == 0
    assert icp.find_optimal_fraction(corr, dist) == (0.0, 0)
    np.testing.assert_array_equal(icp.run(), src)
    assert icp.lambda_val == 0.95


def test_mixed_dims_fall_back_to_xy(gpu):
    from ficp import FractionalICP
    src = _cloud(30, 5)
    tgt = _move(src, 5.0, [0.2, 0.1])[:, :2]
    icp = FractionalICP(src, tgt)
    assert icp.match_dims == 2
    out = icp.run()
    assert out.shape == src.shape and icp.lambda_val == 1.3
    np.testing.assert_array_equal(out[:, 2], src[:, 2])


def test_target_edited_in_place_rebuilds_the_index(gpu):
    """the cached grid index is keyed on the CONTENT of the matched columns - the reference rebuilds its
    kd-tree from the current `target` on every call (ficp.py:69), so an in-place edit between calls must be seen."""
    from coregistrationgame_b200 import FractionalICP
    tgt, src = _scene(3000, 60, seed=2, dims=3)
    icp = FractionalICP(src, tgt)
    m1, d1 = icp.find_correspondences(icp.source, icp.target)
    icp.target[:, :2] += 7.5                                   # same object, same shape, new contents
    m2, d2 = icp.find_correspondences(icp.source, icp.target)
    ref_idx, ref_d2 = orc.nn_assign_bruteforce(icp.source, icp.target, 3)
    np.testing.assert_array_equal(m2, icp.target[ref_idx])
    np.testing.assert_array_equal(d2, np.sqrt(ref_d2))
    assert not np.array_equal(d1, d2)
