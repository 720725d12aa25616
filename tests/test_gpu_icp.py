"""GPU parity tests of the persistent batched ICP kernel against the CPU oracle and against the
golden vectors recorded from the unmodified reference (tests/golden/).

Tolerances (BASELINE.json north_star): NN indices / trimmed subsets bit-exact (checked through the
pass counts, subset sizes and final poses they determine, and directly in test_gpu_stages.py);
rotation within 1e-6 rad; translation within 1e-5 of the scene extent; RMSE / FRMSD within 1e-6 relative."""
import glob
import os

import numpy as np
import pytest

from oracle import ficp_oracle as orc

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz"))
               if not os.path.basename(p).startswith(("c1_", "next_")))
NOISE_FLOOR = 1e-9


@pytest.fixture(scope="module")
def gpu():
    from coregistrationgame_b200 import _lib
    _lib.require_device()
    return _lib


def _pose_of(src_xy, out_xy):
    """rotation angle and translation of the rigid map src -> out (exact for a rigid map)."""
    T = orc.fit_rigid2d_closed(src_xy, out_xy)
    return np.arctan2(T[1, 0], T[0, 0]), T[:2, 2], T


def _assert_same_pose(src, got, want, extent, what=""):
    a1, t1, T1 = _pose_of(src[:, :2], got[:, :2])
    a2, t2, T2 = _pose_of(src[:, :2], want[:, :2])
    dang = abs(((a1 - a2 + np.pi) % (2 * np.pi)) - np.pi)
    assert dang < 1e-6, f"{what}: rotation differs by {dang} rad"
    # compare the translation at the plot centre (t itself is ill-conditioned far from the origin)
    c = src[:, :2].mean(axis=0)
    p1 = T1[:2, :2] @ c + t1
    p2 = T2[:2, :2] @ c + t2
    assert np.abs(p1 - p2).max() < 1e-5 * extent, f"{what}: translation differs by {np.abs(p1 - p2).max()}"
    assert np.abs(got[:, :2] - want[:, :2]).max() < 1e-5 * extent


@pytest.mark.parametrize("case", CASES)
def test_run_matches_reference_golden(gpu, case):
    """FractionalICP.run() on the inputs the unmodified reference was run on."""
    from ficp import FractionalICP
    g = np.load(os.path.join(GOLDEN, case + ".npz"))
    icp = FractionalICP(g["source"], g["target"], lambda_val=float(g["lambda_val"]),
                        allow_reflection=bool(g["allow_reflection"]))
    out = icp.run()
    extent = max(1.0, float(np.ptp(g["target"][:, :2], axis=0).max()))
    _assert_same_pose(g["source"], out, g["aligned"], extent, case)
    np.testing.assert_array_equal(out[:, 2:], g["aligned"][:, 2:])
    assert icp.lambda_val == float(g["lambda_after"])
    if not (g["val"] < NOISE_FLOOR).any():          # see tests/test_oracle_golden.py about the noise floor
        assert icp.n_passes_ == len(g["k"])
        assert icp.k_ == int(g["k"][-1])
        assert icp.frmsd_ == pytest.approx(float(g["val"][-1]), rel=1e-6)


def test_real_data_c1_matches_reference(gpu):
    """Config 1: Data/2014 plots vs Data/2019 layer (2-D), all 16 plots in ONE batch launch."""
    from coregistrationgame_b200 import IcpBatch, TargetIndex
    g = np.load(os.path.join(GOLDEN, "c1_real_2d.npz"))
    offs = g["offsets"]
    plots = [g["source"][offs[p]:offs[p + 1]] for p in range(len(offs) - 1)]
    ti = TargetIndex(g["target"])
    b = IcpBatch(ti, plots, None, centres=np.zeros((len(plots), 2)), min_k=0, want_final_xy=True)
    out = b.run().results()
    np.testing.assert_array_equal(out["hyp"]["passes"][:, 0], g["passes"])
    np.testing.assert_array_equal(out["hyp"]["k"][:, 0], g["k_final"])
    extent = float(np.ptp(g["target"], axis=0).max())
    for p, s in enumerate(plots):
        _assert_same_pose(s, out["final_xy"][offs[p]:offs[p + 1]], g["aligned"][offs[p]:offs[p + 1]], extent, f"plot {p}")
    b.close()
    ti.close()


def _check_batch_against_oracle(tgt, plots, hyp, atol_xy=1e-6, **kw):
    from coregistrationgame_b200 import IcpBatch, TargetIndex
    from coregistrationgame_b200.batch import compose_world_transform
    kw.setdefault("cta_per_icp", False)      # small test batches would pick the CTA kernel by themselves: test_cta_per_icp_*
    ti = TargetIndex(tgt)
    b = IcpBatch(ti, plots, hyp, **kw)
    out = b.run().results()
    okw = {k: v for k, v in kw.items() if k in ("lambda_val", "threshold", "max_iterations", "allow_reflection", "fixed_frac")}
    for p, src in enumerate(plots):
        ref = orc.run_hypotheses(src, tgt, hyp, centre=b.centres[p], min_k=kw.get("min_k", 3), closed_form=True,
                                 trace_all=True, **okw)
        rows = out["hyp"][p]
        sc = np.array(ref["score"])
        # residuals at the rounding-noise floor (~1e-15 m: trees sitting exactly on CHM points) make k a coin
        # flip for ANY two implementations (tests/test_oracle_golden.py): compare k / passes above the floor only
        real = ~(np.array([t_.records[-1].value for t_ in ref["traces"]]) < NOISE_FLOOR)
        np.testing.assert_array_equal(rows["passes"][real], np.array(ref["passes"])[real], err_msg=f"plot {p}: passes per hypothesis")
        np.testing.assert_array_equal(rows["k"][real], np.array(ref["k"])[real], err_msg=f"plot {p}: trimmed subset size")
        raw = np.array([v for v in rows["frmsd"]])
        fin = np.isfinite(sc) & real
        np.testing.assert_allclose(raw[fin], sc[fin], rtol=1e-6)
        for h in range(hyp.shape[0]):
            A = compose_world_transform(rows[h], b.centres[p])
            got = src[:, :2] @ A[:, :2].T + A[:, 2]
            np.testing.assert_allclose(got, ref["aligned"][h][:, :2], rtol=0, atol=atol_xy)
        if real.all():
            assert int(out["best_hyp"][p]) == ref["best_hyp"]
            assert out["best_key"][p] == ref["best_key"]
    stats = out["stats"]
    assert stats["passes"] == int(out["hyp"]["passes"].sum())
    b.close()
    ti.close()
    return out


@pytest.mark.parametrize("dims", [2, 3])
def test_batch_hypotheses_match_oracle_c2_shape(gpu, dims):
    """Config 2 shape (200 trees vs 1e5 CHM points), a 64-hypothesis slice of the 1024 grid."""
    tgt, plots, _ = orc.synthetic_scene(100000, 200, seed=2, dims=dims, hidden_pose=True)
    hyp = orc.hypothesis_table(16, flips=(0, 1), translations=[(0.0, 0.0), (2.5, -2.5)])
    out = _check_batch_against_oracle(tgt, plots, hyp)
    assert out["stats"]["windows_disabled"] == 0
    assert out["stats"]["global_path_queries"] < 0.05 * out["stats"]["queries"]


def test_batch_window_and_global_paths_agree(gpu):
    """The shared-memory window is an optimisation only: disabling it must not change one bit."""
    from coregistrationgame_b200 import IcpBatch, TargetIndex
    tgt, plots, _ = orc.synthetic_scene(50000, 150, seed=8, dims=3, n_plots=3, hidden_pose=True, out_frac=0.2)
    hyp = orc.hypothesis_table(12, flips=(0, 1), translations=orc.translation_lattice(2, 3.0))
    ti = TargetIndex(tgt)
    outs = []
    for kw in (dict(), dict(disable_window=True), dict(window_margin=0.0), dict(warps_per_cta=4, ctas_per_sm=2)):
        b = IcpBatch(ti, plots, hyp, **kw)
        outs.append(b.run().results())
        b.close()
    assert outs[1]["stats"]["global_path_queries"] == outs[1]["stats"]["queries"]
    assert outs[0]["stats"]["global_path_queries"] < outs[2]["stats"]["global_path_queries"] + 1
    for o in outs[1:]:
        assert o["hyp"].tobytes() == outs[0]["hyp"].tobytes() or _rows_equal_except_flags(o["hyp"], outs[0]["hyp"])
        np.testing.assert_array_equal(o["best_key"], outs[0]["best_key"])
    ti.close()


@pytest.mark.parametrize("dims,dup", [(3, 0), (2, 0), (3, 5)])
def test_skip_test_is_bit_identical_to_searching_every_query(gpu, dims, dup):
    """The skip test (a searched query keeps its runner-up and a lower bound on every other target point; later passes
    re-evaluate the two and skip the search while the nearer one stays inside the bound) must not change one bit.
    Reference: the same batch on the global-grid path, where no query carries a bound and every query is searched on
    every pass.  With exact duplicates among the targets (ties in every neighbourhood) the test must keep deferring to
    the search (lowest original index wins there) and still agree."""
    from coregistrationgame_b200 import IcpBatch, TargetIndex
    tgt, plots, _ = orc.synthetic_scene(60000, 180, seed=31 + dims, dims=dims, n_plots=2, hidden_pose=True, out_frac=0.1,
                                        dup_every=dup)
    hyp = orc.hypothesis_table(16, flips=(0, 1), translations=orc.translation_lattice(2, 2.5))
    ti = TargetIndex(tgt)
    fast = IcpBatch(ti, plots, hyp)
    a = fast.run().results()
    centre0 = fast.centres[0].copy()
    fast.close()
    slow = IcpBatch(ti, plots, hyp, disable_window=True)
    b = slow.run().results()
    slow.close()
    ti.close()
    assert _rows_equal_except_flags(a["hyp"], b["hyp"])
    np.testing.assert_array_equal(a["best_key"], b["best_key"])
    sa, sb = a["stats"], b["stats"]
    assert sa["passes"] == sb["passes"] and sa["queries"] == sb["queries"]
    assert sb["searched_queries"] == sb["queries"]                 # no bound without the window: everything searched
    # most queries skip the search (with duplicated targets every query whose neighbour has a twin keeps searching) ...
    assert sa["searched_queries"] < (0.5 if dup == 0 else 1.0) * sa["queries"]
    assert sa["searched_queries"] >= plots[0].shape[0] * 2 * hyp.shape[0]   # ... but never on the first pass
    # and the oracle agrees with both (per-hypothesis pass counts and trimmed sizes of plot 0)
    ref = orc.run_hypotheses(plots[0], tgt, hyp[:6], centre=centre0, min_k=3, closed_form=True)
    np.testing.assert_array_equal(a["hyp"]["passes"][0, :6], ref["passes"])
    np.testing.assert_array_equal(a["hyp"]["k"][0, :6], ref["k"])


def _rows_equal_except_flags(a, b):
    for f in a.dtype.names:
        if f in ("flags", "pad"):
            continue
        if not np.array_equal(a[f], b[f]):
            return False
    return True


def test_batch_adversarial_and_fixed_fraction(gpu):
    """Config 5: 30 % outlier trees, omissions, duplicated and lattice-tied CHM points; FRMSD-auto mode and
    the fixed trim-fraction sweep 0.5-0.95."""
    tgt, plots, _ = orc.synthetic_scene(20000, 120, seed=5, dims=3, out_frac=0.3, omit_frac=0.3, dup_every=10,
                                        lattice_patch=8, hidden_pose=True)
    hyp = orc.hypothesis_table(8, flips=(0, 1))
    _check_batch_against_oracle(tgt, plots, hyp)
    for frac in (0.5, 0.6, 0.7, 0.8, 0.9, 0.95):
        _check_batch_against_oracle(tgt, plots, hyp[:6], fixed_frac=frac)


def test_batch_exact_ties_everywhere(gpu):
    """Source trees sitting exactly on duplicated / lattice CHM points: many d2 == 0 and equal-distance ties."""
    tgt, _, _ = orc.synthetic_scene(5000, 10, seed=9, dims=2, dup_every=5, lattice_patch=10, hidden_pose=False)
    src = np.vstack([tgt[:60], tgt[:20] + np.array([0.5, 0.5])])       # on lattice nodes + on cell centres
    src = np.vstack([src, src[:7]])                                     # duplicated trees
    hyp = orc.hypothesis_table(4, flips=(0,), translations=[(0.0, 0.0), (1.0, 0.0)])
    _check_batch_against_oracle(tgt, [src], hyp)


@pytest.mark.parametrize("n", [5, 28, 33, 64, 100, 257, 500, 1000])
def test_batch_all_plot_size_classes(gpu, n):
    """Every elements-per-lane instantiation (N <= 32, 64, ..., 1024), 2-D and 3-D, mixed in one batch."""
    tgt, plots, _ = orc.synthetic_scene(30000, n, seed=n, dims=3, n_plots=2, hidden_pose=True, out_frac=0.1)
    small = plots[1][: max(3, n // 2)]
    hyp = orc.hypothesis_table(4, flips=(0, 1))
    _check_batch_against_oracle(tgt, [plots[0], small], hyp)
    _check_batch_against_oracle(tgt[:, :2], [plots[0][:, :2], small[:, :2]], hyp[:3])


def test_batch_reflection_and_single_stage(gpu):
    tgt, plots, _ = orc.synthetic_scene(20000, 90, seed=12, dims=2, hidden_pose=True)
    hyp = orc.hypothesis_table(6, flips=(0, 1))
    _check_batch_against_oracle(tgt, plots, hyp, allow_reflection=True)
    _check_batch_against_oracle(tgt, plots, hyp, lambda_val=1.0, max_iterations=3)


def test_iterate_single_stage_and_stepwise_path(gpu):
    """_iterate() = one stage; the host-stepped stage-kernel path (used above 1024 trees) equals the persistent kernel."""
    from ficp import FractionalICP
    tgt, plots, _ = orc.synthetic_scene(20000, 300, seed=14, dims=3, hidden_pose=False)
    src = orc.pre_transform(plots[0], np.r_[orc.hypothesis_matrix(4.0, 0).ravel(), 1.0, -0.5], plots[0][:, :2].mean(0))
    a = FractionalICP(src, tgt)
    out_a = a._iterate()
    tr = orc.RunTrace()
    ref = orc.icp_stage(src.copy(), tgt, 3, 3.0, trace=tr, closed_form=True)
    assert a.n_passes_ == tr.passes and a.k_ == tr.records[-1].k
    np.testing.assert_allclose(out_a[:, :2], ref[:, :2], atol=1e-7)
    b = FractionalICP(src, tgt)
    b._iterate_stepwise()
    assert b.n_passes_ == a.n_passes_ and b.k_ == a.k_
    np.testing.assert_allclose(b.source[:, :2], out_a[:, :2], atol=1e-7)
    np.testing.assert_allclose(b.transform_, a.transform_, atol=1e-7)


def test_large_plot_uses_stepwise_path(gpu):
    from ficp import FractionalICP
    tgt, plots, _ = orc.synthetic_scene(40000, 1500, seed=15, dims=2, hidden_pose=False)
    src = orc.pre_transform(plots[0], np.r_[orc.hypothesis_matrix(1.0, 0).ravel(), 0.5, 0.3], plots[0][:, :2].mean(0))
    icp = FractionalICP(src, tgt)
    out = icp.run()
    tr = orc.RunTrace()
    ref = orc.ficp_run(src, tgt, trace=tr, closed_form=True)
    assert icp.n_passes_ == tr.passes
    np.testing.assert_allclose(out, ref, atol=1e-6)


@pytest.mark.parametrize("n,dims,extra", [(1500, 2, 0), (3000, 3, 2), (9000, 3, 0)])
def test_device_resident_stepper_equals_host_stepped_path(gpu, n, dims, extra):
    """Plots above 1024 trees: run() drives the loop over device-resident arrays (ficp_stepper_*); the host-stepped path over
    the host-buffer stage entry points (`_iterate_stepwise`) runs the same kernels on the same values - every observable
    must agree BIT FOR BIT, incl. the extra columns the transform never touches and the passes taken."""
    from ficp import FractionalICP
    from coregistrationgame_b200.batch import STAGE2_LAMBDA
    tgt, plots, _ = orc.synthetic_scene(60000, n, seed=21 + n, dims=dims, hidden_pose=False)
    src = orc.pre_transform(plots[0], np.r_[orc.hypothesis_matrix(0.7, 0).ravel(), 0.4, -0.3], plots[0][:, :2].mean(0))
    if extra:
        src = np.column_stack([src, np.random.default_rng(n).normal(size=(n, extra))])
    a = FractionalICP(src, tgt, max_iterations=25)
    out_a = a.run()
    b = FractionalICP(src, tgt, max_iterations=25)
    for lam in (3.0, STAGE2_LAMBDA[b.match_dims]):
        b.lambda_val = lam
        b._iterate_stepwise()
    assert a.n_passes_ == b.n_passes_ and a.k_ == b.k_ and a.n_passes_ >= 4
    np.testing.assert_array_equal(out_a, b.source)
    np.testing.assert_array_equal(a.transform_, b.transform_)
    assert a.frmsd_ == b.frmsd_
    np.testing.assert_array_equal(out_a[:, 2:], src[:, 2:])


def test_determinism_bitwise(gpu):
    from coregistrationgame_b200 import register_batch
    tgt, plots, _ = orc.synthetic_scene(30000, 200, seed=3, dims=3, n_plots=4, hidden_pose=True)
    hyp = orc.hypothesis_table(32, flips=(0, 1))
    a = register_batch(plots, tgt, hyp)
    b = register_batch(plots, tgt, hyp, warps_per_cta=8)
    assert _rows_equal_except_flags(a["hyp"], b["hyp"])
    np.testing.assert_array_equal(a["best_key"], b["best_key"])


@pytest.mark.parametrize("n,dims", [(40, 2), (150, 3), (200, 2), (500, 3), (1000, 3)])
def test_helper_warps_are_bit_identical(gpu, n, dims):
    """Elastic kernel: warps without an ICP of their own (small batches: one stand over 8 GPUs, one start pose per
    plot; or a plot running out of hypotheses) take nearest-neighbour rounds of the ICPs in flight; trimming and fit
    keep their arithmetic, so the plain kernel and every team size return the same bits."""
    from coregistrationgame_b200 import IcpBatch, TargetIndex
    tgt, plots, _ = orc.synthetic_scene(40000, n, seed=21 + n, dims=dims, n_plots=3, hidden_pose=True, out_frac=0.15,
                                        dup_every=7)
    hyp = orc.hypothesis_table(6, flips=(0, 1), translations=[(0.0, 0.0), (60.0, -45.0)])   # second half: off the window
    ti = TargetIndex(tgt)
    base = None
    for kw in (dict(team_warps=1, helpers=False), dict(team_warps=1), dict(team_warps=2), dict(team_warps=4),
               dict(team_warps=8), dict(team_warps=0), dict(team_warps=4, disable_window=True),
               dict(team_warps=2, warps_per_cta=6), dict(team_warps=1, warps_per_cta=3, ctas_per_sm=1)):
        b = IcpBatch(ti, plots, hyp, cta_per_icp=False, **kw)
        out = b.run().results()
        assert b.info["helpers"] == int(kw.get("helpers", True))          # auto: 36 ICPs -> elastic
        want = kw["team_warps"]
        if want in (1, 2, 4, 8):
            assert b.info["team_warps"] == min(want, b.info["elems_per_lane"], max(1, b.info["warps_per_cta"]))
        else:
            assert b.info["team_warps"] >= 2          # 36 ICPs on a 148-SM GPU: the planner teams up by itself
        b.close()
        if base is None:
            base = out
            ref = orc.run_hypotheses(plots[0], tgt, hyp[:3], centre=b.centres[0], min_k=3, closed_form=True)
            np.testing.assert_array_equal(out["hyp"]["passes"][0, :3], ref["passes"])
            continue
        assert _rows_equal_except_flags(out["hyp"], base["hyp"]), kw
        np.testing.assert_array_equal(out["best_key"], base["best_key"])
        assert out["stats"]["passes"] == base["stats"]["passes"]
    # one start pose per plot (C4 shape), with final positions
    one = [IcpBatch(ti, plots, None, want_final_xy=True, team_warps=t, helpers=(t > 1), cta_per_icp=False) for t in (1, 4)]
    res = [b.run().results() for b in one]
    assert _rows_equal_except_flags(res[0]["hyp"], res[1]["hyp"])
    np.testing.assert_array_equal(res[0]["final_xy"], res[1]["final_xy"])
    for b in one:
        b.close()
    ti.close()


@pytest.mark.parametrize("n,dims", [(40, 2), (64, 3), (100, 3), (150, 3), (200, 2), (500, 3), (1000, 3)])
def test_cta_per_icp_is_bit_identical(gpu, n, dims):
    """CTA-per-ICP kernel (icp_team.cu: every phase of a pass cooperative across 32 e threads - the shape for batches
    smaller than the machine, e.g. ONE stand x 4096 start poses sharded over 8 GPUs) vs the one-warp-per-ICP kernel:
    identical bits in every result row, best key, final positions and pass count, for every thread-count class, on the
    window and on the global-grid path, with duplicated targets (ties) and start poses thrown off the window."""
    from coregistrationgame_b200 import IcpBatch, TargetIndex
    tgt, plots, _ = orc.synthetic_scene(40000, n, seed=121 + n, dims=dims, n_plots=3, hidden_pose=True, out_frac=0.15,
                                        dup_every=7)
    small = plots[2][: max(33, n // 3)]                     # a smaller plot in the same size class
    hyp = orc.hypothesis_table(6, flips=(0, 1), translations=[(0.0, 0.0), (60.0, -45.0)])
    ti = TargetIndex(tgt)
    outs = []
    for kw in (dict(cta_per_icp=False), dict(cta_per_icp=True), dict(cta_per_icp=True, disable_window=True),
               dict(cta_per_icp=True, ctas_per_sm=1), dict(cta_per_icp=True, fixed_frac=0.8), dict(cta_per_icp=False, fixed_frac=0.8)):
        b = IcpBatch(ti, [plots[0], plots[1], small], hyp, **kw)
        assert b.info["cta_per_icp"] == int(kw["cta_per_icp"])
        if kw["cta_per_icp"]:
            assert b.info["warps_per_cta"] in (b.info["elems_per_lane"], b.info["elems_per_lane"] // 2)   # one or two trees per thread
        outs.append(b.run().results())
        b.close()
    for o in outs[1:4]:
        assert _rows_equal_except_flags(o["hyp"], outs[0]["hyp"])
        np.testing.assert_array_equal(o["best_key"], outs[0]["best_key"])
        assert o["stats"]["passes"] == outs[0]["stats"]["passes"] and o["stats"]["queries"] == outs[0]["stats"]["queries"]
    assert _rows_equal_except_flags(outs[4]["hyp"], outs[5]["hyp"])
    np.testing.assert_array_equal(outs[4]["best_key"], outs[5]["best_key"])
    assert outs[2]["stats"]["global_path_queries"] == outs[2]["stats"]["queries"]
    # the skip test works in this shape too (first pass of each stage-1 run searches everything, later passes few)
    assert outs[1]["stats"]["searched_queries"] < 0.6 * outs[1]["stats"]["queries"]
    # and the oracle agrees (pass counts, k of plot 0)
    ref = orc.run_hypotheses(plots[0], tgt, hyp[:3], centre=plots[0][:, :2].mean(axis=0), min_k=3, closed_form=True)
    np.testing.assert_array_equal(outs[1]["hyp"]["passes"][0, :3], ref["passes"])
    np.testing.assert_array_equal(outs[1]["hyp"]["k"][0, :3], ref["k"])
    # one start pose per plot with final positions
    one = [IcpBatch(ti, plots, None, want_final_xy=True, cta_per_icp=c) for c in (False, True)]
    res = [b.run().results() for b in one]
    assert _rows_equal_except_flags(res[0]["hyp"], res[1]["hyp"])
    np.testing.assert_array_equal(res[0]["final_xy"], res[1]["final_xy"])
    for b in one:
        b.close()
    ti.close()


def test_cta_per_icp_against_oracle(gpu):
    """The CTA-per-ICP kernel against the oracle directly: C2 slice, adversarial scene with the fixed-fraction sweep,
    exact ties everywhere, reflection / single stage / iteration cap."""
    tgt, plots, _ = orc.synthetic_scene(100000, 200, seed=2, dims=3, hidden_pose=True)
    _check_batch_against_oracle(tgt, plots, orc.hypothesis_table(8, flips=(0, 1)), cta_per_icp=True)
    tgt, plots, _ = orc.synthetic_scene(20000, 120, seed=5, dims=3, out_frac=0.3, omit_frac=0.3, dup_every=10,
                                        lattice_patch=8, hidden_pose=True)
    hyp = orc.hypothesis_table(8, flips=(0, 1))
    _check_batch_against_oracle(tgt, plots, hyp, cta_per_icp=True)
    for frac in (0.5, 0.7, 0.95):
        _check_batch_against_oracle(tgt, plots, hyp[:6], fixed_frac=frac, cta_per_icp=True)
    tgt, _, _ = orc.synthetic_scene(5000, 10, seed=9, dims=2, dup_every=5, lattice_patch=10, hidden_pose=False)
    src = np.vstack([tgt[:60], tgt[:20] + np.array([0.5, 0.5])])
    src = np.vstack([src, src[:7]])
    _check_batch_against_oracle(tgt, [src], orc.hypothesis_table(4, flips=(0,), translations=[(0.0, 0.0), (1.0, 0.0)]), cta_per_icp=True)
    tgt, plots, _ = orc.synthetic_scene(20000, 90, seed=12, dims=2, hidden_pose=True)
    hyp = orc.hypothesis_table(6, flips=(0, 1))
    _check_batch_against_oracle(tgt, plots, hyp, allow_reflection=True, cta_per_icp=True)
    _check_batch_against_oracle(tgt, plots, hyp, lambda_val=1.0, max_iterations=3, cta_per_icp=True)


def test_small_batches_pick_the_cta_kernel_by_themselves(gpu):
    """Planner: a batch far smaller than the machine (one stand x a few hundred start poses) runs CTA-per-ICP, a batch
    that fills it runs warp-per-ICP; plots of <= 32 trees are one warp either way."""
    from coregistrationgame_b200 import IcpBatch, TargetIndex
    tgt, plots, _ = orc.synthetic_scene(30000, 120, seed=3, dims=3, n_plots=2, hidden_pose=True)
    ti = TargetIndex(tgt)
    small = IcpBatch(ti, [plots[0]], orc.hypothesis_table(64, flips=(0, 1)))
    big = IcpBatch(ti, plots, orc.hypothesis_table(2048, flips=(0, 1)))
    tiny = IcpBatch(ti, [plots[0][:20]], orc.hypothesis_table(8, flips=(0, 1)), cta_per_icp=True)
    assert small.info["cta_per_icp"] == 1 and big.info["cta_per_icp"] == 0 and tiny.info["cta_per_icp"] == 0
    for b in (small, big, tiny):
        b.close()
    ti.close()


@pytest.mark.parametrize("dims", [2, 3])
def test_full_size_c3_properties(gpu, dims):
    """Config 3 at full size (500 trees vs 1e6 CHM points, 4096 hypotheses): size-independent properties
    plus a strided sample of hypotheses checked against the oracle."""
    from coregistrationgame_b200 import IcpBatch, TargetIndex
    from coregistrationgame_b200.batch import compose_world_transform
    tgt, plots, poses = orc.synthetic_scene(1000000, 500, seed=3, dims=dims, hidden_pose=True)
    src = plots[0]
    hyp = orc.hypothesis_table(128, flips=(0, 1), translations=orc.translation_lattice(4, 2.5))
    assert hyp.shape[0] == 4096
    ti = TargetIndex(tgt)
    b = IcpBatch(ti, [src], hyp)
    out = b.run().results()
    rows = out["hyp"][0]
    # (1) the winner undoes the hidden pose: the hidden pose was R(th) about the centroid + d
    best = int(out["best_hyp"][0])
    A = compose_world_transform(rows[best], b.centres[0])
    ang = np.degrees(np.arctan2(A[1, 0], A[0, 0]))
    th = poses[0][0]
    assert abs(((ang + th + 180) % 360) - 180) < 0.5, (ang, th)
    assert rows["k"][best] > 0.7 * 500 and rows["rmse"][best] < 1.5
    # (2) ranking key = min over hypotheses of (fp32 score, id)
    score = np.where(rows["k"] >= 3, rows["frmsd"], np.inf).astype(np.float32)
    keys = (score.view(np.uint32).astype(np.uint64) << np.uint64(32)) | np.arange(4096, dtype=np.uint64)
    assert out["best_key"][0] == keys.min()
    # (3) every result is a proper rotation and FRMSD = (N/k)^lambda * rmse
    det = rows["m00"] * rows["m11"] - rows["m01"] * rows["m10"]
    np.testing.assert_allclose(det, np.where(hyp[:, 0] * hyp[:, 3] - hyp[:, 1] * hyp[:, 2] > 0, 1.0, -1.0), atol=1e-12)
    lam2 = orc.STAGE2_LAMBDA[dims]
    np.testing.assert_allclose(rows["frmsd"], (500.0 / rows["k"]) ** lam2 * rows["rmse"], rtol=1e-12)
    # (4) idempotence: restarting stage 2 from a converged pose stops almost at once and cannot get worse
    sub = np.arange(0, 4096, 128)
    again = np.stack([np.r_[rows["m00"][h], rows["m01"][h], rows["m10"][h], rows["m11"][h],
                            rows["cx"][h] - b.centres[0][0], rows["cy"][h] - b.centres[0][1]] for h in sub])
    b2 = IcpBatch(ti, [src], again, n_stages=1, lambda_val=orc.STAGE2_LAMBDA[dims])
    r2 = b2.run().results()["hyp"][0]
    assert np.median(r2["passes"]) <= 2 and np.mean(r2["passes"] <= 3) >= 0.9, r2["passes"]
    assert (r2["frmsd"] <= rows["frmsd"][sub] * (1 + 1e-9) + 1e-9).all()
    b2.close()
    # (5) strided sample against the oracle (kd-tree NN with the same tie rule)
    sample = np.arange(7, 4096, 512)
    ref = orc.run_hypotheses(src, tgt, hyp[sample], centre=b.centres[0], closed_form=True)
    np.testing.assert_array_equal(rows["passes"][sample], np.array(ref["passes"]))
    np.testing.assert_array_equal(rows["k"][sample], np.array(ref["k"]))
    for j, h in enumerate(sample):
        A = compose_world_transform(rows[h], b.centres[0])
        got = src[:, :2] @ A[:, :2].T + A[:, 2]
        np.testing.assert_allclose(got, ref["aligned"][j][:, :2], rtol=0, atol=1e-6)
    assert out["stats"]["passes"] == int(rows["passes"].sum())
    b.close()
    ti.close()


def test_many_plots_one_pose_each_c4_shape(gpu):
    """Config 4 shape, scaled to the test budget: 600 plots x 150 trees against one shared 2e6-point CHM, ONE ICP per
    plot (the pose the user dragged it to), all in one launch.  A strided sample is checked against the oracle."""
    from coregistrationgame_b200 import IcpBatch, TargetIndex
    tgt, plots, _ = orc.synthetic_scene(2_000_000, 150, seed=4, dims=3, n_plots=600, hidden_pose=False)
    rng = np.random.default_rng(0)
    starts = []
    for p in plots:
        row = np.r_[orc.hypothesis_matrix(rng.uniform(-6.0, 6.0), 0).ravel(), rng.uniform(-2.0, 2.0, 2)]
        starts.append(orc.pre_transform(p, row, p[:, :2].mean(axis=0)))
    ti = TargetIndex(tgt)
    b = IcpBatch(ti, starts, None, centres=np.zeros((len(starts), 2)), min_k=0, want_final_xy=True)
    assert b.info["n_hyp_local"] == 1 and b.info["warps_per_cta"] == b.info["team_warps"]
    out = b.run().results()
    rows = out["hyp"][:, 0]
    offs = b.offsets
    from scipy.spatial import cKDTree
    tree = cKDTree(tgt)
    for p in range(0, len(starts), 25):
        tr = orc.RunTrace()
        ref = starts[p]
        for lam in (3.0, orc.STAGE2_LAMBDA[3]):
            ref = orc.icp_stage(ref, tgt, 3, lam, nn="tree", tree=tree, trace=tr, closed_form=True)
        assert rows["passes"][p] == tr.passes and rows["k"][p] == tr.records[-1].k, p
        np.testing.assert_allclose(out["final_xy"][offs[p]:offs[p + 1]], ref[:, :2], rtol=0, atol=1e-6)
    # most plots snap back onto their trees (position noise 0.3 m in XY, 1 m in Z)
    assert np.mean(rows["rmse"] < 1.6) > 0.9
    assert out["stats"]["passes"] == int(rows["passes"].sum())
    b.close()
    ti.close()


def test_batch_tiny_targets_and_off_map_poses(gpu):
    """Degenerate scenes: 1-3 CHM points; and start poses thrown kilometres off the stand (the exact search must
    stay bounded and still agree with the oracle)."""
    from coregistrationgame_b200 import IcpBatch, TargetIndex
    rng = np.random.default_rng(3)
    src = rng.normal(size=(12, 2)) * 4.0
    hyp = orc.hypothesis_table(4, flips=(0,))
    for m in (1, 2, 3):
        # When every inlier maps to the SAME CHM point the reference's centred targets are +-1 ulp instead of 0 and
        # its SVD returns a rounding-noise rotation (parity unpinned); the kernel treats that H as 0 -> R = I, i.e. a
        # pure translation.  Check that, not the noise.
        tgt = rng.normal(size=(m, 2)) * 3.0
        ti = TargetIndex(tgt)
        b = IcpBatch(ti, [src], hyp, min_k=0)
        rows = b.run().results()["hyp"][0]
        assert np.isfinite(rows["frmsd"]).all() and (rows["passes"] >= 2).all() and (rows["k"] >= 1).all()
        det = rows["m00"] * rows["m11"] - rows["m01"] * rows["m10"]
        np.testing.assert_allclose(det, 1.0, atol=1e-12)
        if m == 1:   # rotation part untouched: only translations are ever fitted
            np.testing.assert_allclose(np.stack([rows["m00"], rows["m01"], rows["m10"], rows["m11"]], 1), hyp[:, :4], atol=1e-15)
        b.close()
        ti.close()
    tgt, plots, _ = orc.synthetic_scene(200000, 60, seed=6, dims=3, hidden_pose=True)
    far = orc.hypothesis_table(3, flips=(0,), translations=[(0.0, 0.0), (25000.0, -18000.0), (-400.0, 90000.0)])
    out = _check_batch_against_oracle(tgt, plots, far)
    assert out["stats"]["global_path_queries"] > 0


def test_full_size_c4_properties(gpu):
    """Config 4 at full size: 10^4 plots x 150 trees against one shared 10^7-point CHM, one ICP per plot, one
    launch.  Size-independent properties + a strided oracle sample."""
    from scipy.spatial import cKDTree
    from coregistrationgame_b200 import IcpBatch, TargetIndex
    tgt, plots, _ = orc.synthetic_scene(10_000_000, 150, seed=4, dims=3, n_plots=10000, hidden_pose=False)
    rng = np.random.default_rng(1)
    starts, rows_in = [], []
    for p in plots:
        row = np.r_[orc.hypothesis_matrix(rng.uniform(-5.0, 5.0), 0).ravel(), rng.uniform(-1.5, 1.5, 2)]
        starts.append(orc.pre_transform(p, row, p[:, :2].mean(axis=0)))
        rows_in.append(row)
    ti = TargetIndex(tgt)
    info = ti.info()
    assert info["m"] == 10_000_000
    b = IcpBatch(ti, starts, None, centres=np.zeros((len(starts), 2)), min_k=0, want_final_xy=True)
    out = b.run().results()
    rows = out["hyp"][:, 0]
    offs = b.offsets
    # proper rotations, FRMSD = (N/k)^lambda * rmse, Z untouched by construction, every plot did >= 2 passes
    np.testing.assert_allclose(rows["m00"] * rows["m11"] - rows["m01"] * rows["m10"], 1.0, atol=1e-12)
    np.testing.assert_allclose(rows["frmsd"], (150.0 / rows["k"]) ** orc.STAGE2_LAMBDA[3] * rows["rmse"], rtol=1e-12)
    assert (rows["passes"] >= 2).all() and out["stats"]["passes"] == int(rows["passes"].sum())
    # the start perturbation is undone: rotation recovered to < 1 degree for the vast majority of plots
    ang = np.degrees(np.arctan2(rows["m10"], rows["m00"]))
    want = -np.array([np.degrees(np.arctan2(r[2], r[0])) for r in rows_in])
    assert np.mean(np.abs(ang - want) < 1.0) > 0.9
    assert np.mean(rows["rmse"] < 1.6) > 0.9
    # rigid: pairwise distances of a plot are preserved by its final pose
    for p in (0, 4999, 9999):
        a, f = starts[p][:, :2], out["final_xy"][offs[p]:offs[p + 1]]
        np.testing.assert_allclose(np.linalg.norm(a[1:] - a[:-1], axis=1), np.linalg.norm(f[1:] - f[:-1], axis=1), atol=1e-9)
    # strided oracle sample (kd-tree NN with the same tie rule)
    tree = cKDTree(tgt)
    for p in range(0, 10000, 1250):
        tr = orc.RunTrace()
        ref = starts[p]
        for lam in (3.0, orc.STAGE2_LAMBDA[3]):
            ref = orc.icp_stage(ref, tgt, 3, lam, nn="tree", tree=tree, trace=tr, closed_form=True)
        assert rows["passes"][p] == tr.passes and rows["k"][p] == tr.records[-1].k, p
        np.testing.assert_allclose(out["final_xy"][offs[p]:offs[p + 1]], ref[:, :2], rtol=0, atol=1e-6)
    b.close()
    ti.close()


def test_xy_plots_against_xyz_built_index(gpu):
    """ficp.py:40: matching falls back to XY when only one side has heights.  Here the CHM index is built WITH Z
    (packed 32-byte records) and reused for 2-column plots: batch kernel, stand-alone query and match-remove must
    all read X, Y out of the packed layout."""
    from coregistrationgame_b200 import IcpBatch, TargetIndex
    tgt, plots, _ = orc.synthetic_scene(40000, 90, seed=21, dims=3, n_plots=2, hidden_pose=True, dup_every=13)
    hyp = orc.hypothesis_table(6, flips=(0, 1))
    ti = TargetIndex(tgt)                       # has_z = True
    assert ti.info()["has_z"]
    plots_xy = [p[:, :2].copy() for p in plots]
    b = IcpBatch(ti, plots_xy, hyp)
    assert b.match_dims == 2
    out = b.run().results()
    for p, src in enumerate(plots_xy):
        ref = orc.run_hypotheses(src, tgt[:, :2], hyp, centre=b.centres[p], closed_form=True)
        np.testing.assert_array_equal(out["hyp"][p]["passes"], np.array(ref["passes"]))
        np.testing.assert_array_equal(out["hyp"][p]["k"], np.array(ref["k"]))
        assert out["best_key"][p] == ref["best_key"]
    b.close()
    idx, dist = ti.query(plots_xy[0])           # XY query against the XYZ-built index
    ridx, rd2 = orc.nn_assign_bruteforce(plots_xy[0], tgt[:, :2], 2)
    np.testing.assert_array_equal(idx, ridx)
    np.testing.assert_array_equal(dist, np.sqrt(rd2))
    ti.close()


def test_randomised_parity_soak(gpu):
    """A short run of tools/fuzz_parity.py (random sizes, modes, duplicates, off-map poses, tiny targets): every
    hypothesis whose parity is pinned must agree with the oracle.  The long soak (tens of thousands of hypotheses)
    is run by hand: `python tools/fuzz_parity.py 600 <seed>`."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "fuzz_parity.py"), "15", "11"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]
    assert "0 mismatching cases" in out.stdout


def test_packed_world_translation_matches_host_composition(gpu):
    """The winners' records carry the world-frame translation b = c - M centre, evaluated on the device in the arithmetic
    of batch.compose_world_transforms: the two routes (winners only / per-hypothesis table) return identical transforms."""
    from coregistrationgame_b200 import register_batch
    tgt, plots, _ = orc.synthetic_scene(40000, 120, seed=17, dims=3, n_plots=6, hidden_pose=True)
    tgt = tgt + [420000.0, 6483000.0, 0.0]
    plots = [p + [420000.0, 6483000.0, 0.0] for p in plots]
    hyp = orc.hypothesis_table(8, flips=(0, 1), translations=orc.translation_lattice(2, 2.5))
    a = register_batch(plots, tgt, hyp, per_hypothesis=False)
    b = register_batch(plots, tgt, hyp, per_hypothesis=True)
    np.testing.assert_array_equal(a["best_key"], b["best_key"])
    np.testing.assert_array_equal(a["best_transform"], b["best_transform"])


def test_stacked_input_equals_list_of_plots(gpu):
    """register_batch((rows, offsets), ...) - the input form for thousands of plots - gives exactly what the list of per-plot
    arrays gives; malformed offsets are refused."""
    from coregistrationgame_b200 import register_batch
    tgt, plots, _ = orc.synthetic_scene(50000, 60, seed=8, dims=3, n_plots=37, hidden_pose=True)
    plots = [p[: 40 + (i % 3) * 10] for i, p in enumerate(plots)]        # three size classes
    hyp = orc.hypothesis_table(4, flips=(0, 1))
    a = register_batch(plots, tgt, hyp)
    rows = np.vstack(plots)
    offs = np.concatenate([[0], np.cumsum([len(p) for p in plots])])
    b = register_batch((rows, offs), tgt, hyp)
    np.testing.assert_array_equal(a["best_key"], b["best_key"])
    np.testing.assert_array_equal(a["best_transform"], b["best_transform"])
    assert _rows_equal_except_flags(a["hyp"], b["hyp"])
    with pytest.raises(ValueError):
        register_batch((rows, offs[:-1]), tgt, hyp)
    with pytest.raises(ValueError):
        register_batch((rows, np.r_[offs[:3], offs[2:]]), tgt, hyp)


def test_library_centres_and_sliced_upload_equal_explicit_centres(gpu):
    """ficp_batch_create with centres = NULL takes each plot's centroid itself (in-order mean, the bits of
    `rows[:, :2].mean(axis=0)`) and, from 65 536 rows on, sends the rows to the device in slices while it prepares the next
    one: the per-plot results equal those of a batch given the numpy centroids explicitly, bit for bit - for a batch below
    and one above the slicing threshold (ragged sizes, UTM offsets, Fortran-ordered plots), with 1 and 3 host threads."""
    from coregistrationgame_b200 import IcpBatch, TargetIndex
    tgt, plots, _ = orc.synthetic_scene(60000, 160, seed=21, dims=3, n_plots=520, hidden_pose=False)
    tgt = tgt + [420000.0, 6483000.0, 0.0]
    plots = [np.asfortranarray(p[: 100 + (i * 7) % 61] + [420000.0, 6483000.0, 0.0]) for i, p in enumerate(plots)]
    ti = TargetIndex(tgt)
    for ps in (plots[:40], plots):                                       # 5 k rows (one slice) / 67 k rows (four slices)
        assert (sum(len(p) for p in ps) >= 65536) == (len(ps) > 40)
        cen = np.array([np.ascontiguousarray(p)[:, :2].mean(axis=0) for p in ps])
        want = IcpBatch(ti, ps, None, centres=cen, min_k=0).run().results()
        for threads in ("1", "3"):
            os.environ["FICP_HOST_THREADS"] = threads
            try:
                b = IcpBatch(ti, ps, None, min_k=0)
                got = b.run().results()
                np.testing.assert_array_equal(b.centres, cen)
            finally:
                del os.environ["FICP_HOST_THREADS"]
            np.testing.assert_array_equal(got["best_key"], want["best_key"])
            assert _rows_equal_except_flags(got["hyp"], want["hyp"])
            assert got["stats"]["passes"] == want["stats"]["passes"]
            assert b.info["rows_direct"] == 0                            # pageable numpy memory: staged by the library
        # page-locked rows (a torch pinned tensor) go to the device as they are and are split there: same bits again;
        # FICP_HOST_STAGING=1 sends the same array through the staging route
        import torch
        stacked = np.vstack([np.ascontiguousarray(p) for p in ps])
        offs = np.concatenate([[0], np.cumsum([len(p) for p in ps])]).astype(np.int64)
        for cols, md3 in ((3, True), (2, False), (4, True), (5, True)):
            wide = np.hstack([stacked, stacked[:, :2]])[:, :cols] if cols > 3 else stacked[:, :cols]
            pin = torch.empty(wide.shape, dtype=torch.float64, pin_memory=True).numpy()
            pin[...] = wide
            ref = want if md3 else IcpBatch(ti, (np.ascontiguousarray(wide), offs), None, min_k=0).run().results()
            for staging in (False, True):
                if staging:
                    os.environ["FICP_HOST_STAGING"] = "1"
                try:
                    b = IcpBatch(ti, (pin, offs), None, min_k=0)
                    got = b.run().results()
                finally:
                    os.environ.pop("FICP_HOST_STAGING", None)
                assert b.info["rows_direct"] == (0 if (staging or cols > 4) else 1), (cols, staging, b.info)
                np.testing.assert_array_equal(got["best_key"], ref["best_key"])
                assert _rows_equal_except_flags(got["hyp"], ref["hyp"])
                b.close()
    with pytest.raises(ValueError, match="finite"):                      # the read-only host pass still refuses NaN rows
        bad = torch.empty(stacked.shape, dtype=torch.float64, pin_memory=True).numpy()
        bad[...] = stacked
        bad[len(bad) // 2, 1] = np.nan
        IcpBatch(ti, (bad, offs), None, min_k=0)
    ti.close()


def test_planner_picks_the_kernel_shape_by_batch_size(gpu):
    """Auto launch shape (capi.cu): CTA-per-ICP up to 14 ICPs per SM, warp-per-ICP above; plots of <= 32 trees are one warp
    either way."""
    from coregistrationgame_b200 import IcpBatch, TargetIndex, _lib
    sms = _lib.device_props()["sms"]
    tgt, plots, _ = orc.synthetic_scene(60000, 40, seed=5, dims=3, n_plots=15 * sms + 7, hidden_pose=False)
    ti = TargetIndex(tgt)
    ident = np.array([[1.0, 0.0, 0.0, 1.0, 0.0, 0.0]])
    hyp = orc.hypothesis_table(8, flips=(0, 1))
    cases = [(plots[:11 * sms], ident, 1), (plots, ident, 0), (plots[:sms // 4], hyp, 1), (plots[:sms], hyp, 0)]
    rows = []
    for ps, h, want in cases:
        b = IcpBatch(ti, ps, h, min_k=0)
        assert b.info["cta_per_icp"] == want, (len(ps), h.shape[0], b.info)
        rows.append(b.run().results())
        b.close()
    # the same plots give the same rows whichever kernel the planner picked
    n0 = 11 * sms
    assert _rows_equal_except_flags(rows[0]["hyp"], rows[1]["hyp"][:n0])
    small = IcpBatch(ti, [p[:30] for p in plots[:5]], hyp, min_k=0)
    assert small.info["cta_per_icp"] == 0 and small.info["elems_per_lane"] == 1
    small.close()
    ti.close()
