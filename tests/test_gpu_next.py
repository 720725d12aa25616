"""GPU parity of the steps right after the ICP: greedy match-and-remove and the transform
record, against golden vectors from the unmodified reference classes and against the oracle."""
import glob
import os

import numpy as np
import pytest

from oracle import ficp_oracle as orc

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
NEXT_REMOVE = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "next_remove_*.npz")))


@pytest.fixture(scope="module")
def gpu():
    from coregistrationgame_b200 import _lib
    _lib.require_device()
    return _lib


@pytest.mark.parametrize("case", NEXT_REMOVE)
def test_remove_matches_matches_reference_golden(gpu, case):
    from coregistrationgame_b200.matching import remove_matches
    g = np.load(os.path.join(GOLDEN, case + ".npz"))
    matched = remove_matches(g["plot"], g["chm"], float(g["pct"]))
    np.testing.assert_array_equal(matched[matched >= 0], g["removed"])       # bit-exact indices, same order
    np.testing.assert_array_equal(matched, orc.remove_matches_oracle(g["plot"], g["chm"], float(g["pct"])))


@pytest.mark.parametrize("dims3", [True, False])
def test_remove_matches_random_vs_oracle(gpu, dims3):
    from coregistrationgame_b200 import TargetIndex
    from coregistrationgame_b200.matching import remove_matches, remove_matches_batch
    rng = np.random.default_rng(9)
    tgt, plots, _ = orc.synthetic_scene(50000, 120, seed=10, dims=3, n_plots=6, hidden_pose=False, dup_every=11)
    if not dims3:
        tgt = tgt.copy()
        tgt[17, 2] = np.nan                      # one missing CHM height -> XY matching for everything
    plots = [p + np.column_stack([rng.normal(0, 0.5, (len(p), 2)), rng.normal(0, 0.5, len(p))]) for p in plots]
    plots[2] = np.vstack([plots[2], plots[2][:10]])      # duplicated trees compete for the same CHM trees
    for pct in (5, 15, 60):
        batch = remove_matches_batch(plots, tgt, pct)
        for p, m in zip(plots, batch):
            np.testing.assert_array_equal(m, orc.remove_matches_oracle(p, tgt, pct))
            got = m[m >= 0]
            assert len(np.unique(got)) == len(got)      # a CHM tree is removed at most once
    # a prebuilt index can be reused
    ti = TargetIndex(tgt if dims3 else tgt[:, :2], use_z=dims3)
    np.testing.assert_array_equal(remove_matches(plots[0], tgt, 15, index=ti), batch_first(plots, tgt))
    ti.close()
    # empty inputs
    assert remove_matches(np.empty((0, 3)), tgt).shape == (0,)
    np.testing.assert_array_equal(remove_matches(plots[0], np.empty((0, 3))), np.full(len(plots[0]), -1))


def batch_first(plots, tgt):
    return orc.remove_matches_oracle(plots[0], tgt, 15)


def test_transform_record_matches_reference_golden(gpu):
    from coregistrationgame_b200.matching import transform_record
    g = np.load(os.path.join(GOLDEN, "next_get_transform.npz"))
    for i in range(int(g["n"])):
        rec = transform_record(g[f"orig_{i}"], g[f"cur_{i}"], bool(g[f"flipped_{i}"]))
        np.testing.assert_allclose([rec["r00"], rec["r01"], rec["r10"], rec["r11"]], g[f"R_{i}"].ravel(), atol=1e-12)
        np.testing.assert_allclose([rec["tx"], rec["ty"]], g[f"t_{i}"], atol=1e-6)
        assert rec["flip"] == bool(g[f"flipped_{i}"])
        for key in ("tx", "ty", "r00", "r01", "r10", "r11"):
            assert isinstance(rec[key], float)
    with pytest.raises(ValueError):
        transform_record(np.empty((0, 2)), np.empty((0, 2)))


def test_identity_record_for_unmoved_plot(gpu):
    """tests/test_transformation_serialization.py:21-47 of the reference: an unmodified plot serialises to identity."""
    from coregistrationgame_b200.matching import transform_record
    xy = np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0]])
    rec = transform_record(xy, xy.copy())
    assert np.isclose(rec["tx"], 0.0) and np.isclose(rec["ty"], 0.0)
    assert np.isclose(rec["r00"], 1.0) and np.isclose(rec["r01"], 0.0) and np.isclose(rec["r10"], 0.0) and np.isclose(rec["r11"], 1.0)


def test_radial_crop_matches_cdist_semantics(gpu):
    """chm_plot.py:144-148: keep rows with Euclidean XY distance <= dist, original order."""
    from coregistrationgame_b200 import TargetIndex
    from coregistrationgame_b200.matching import radial_crop
    rng = np.random.default_rng(4)
    pts = np.column_stack([rng.uniform(420000, 420400, 200000), rng.uniform(6483000, 6483400, 200000)])
    pts[:50] = pts[50:100]                                    # duplicates
    pts[100] = [420200.0 + 70.0, 6483200.0]                   # exactly on the circle (3-4-5 style exactness)
    ti = TargetIndex(pts, use_z=False)
    for (x, y, d) in ((420200.0, 6483200.0, 70.0), (420000.0, 6483000.0, 35.5), (419000.0, 6483200.0, 10.0),
                      (420200.0, 6483200.0, 0.0), (420200.0, 6483200.0, 1e6)):
        got = radial_crop(ti, x, y, d)
        dx, dy = pts[:, 0] - x, pts[:, 1] - y
        want = np.flatnonzero(np.sqrt(dx * dx + dy * dy) <= d)
        np.testing.assert_array_equal(got, want)
    np.testing.assert_array_equal(radial_crop(pts[:1000], 420200.0, 6483200.0, 150.0),
                                  np.flatnonzero(np.hypot(pts[:1000, 0] - 420200.0, pts[:1000, 1] - 6483200.0) <= 150.0))
    ti.close()


def test_gui_hypotheses_follow_plot_key_semantics(gpu):
    """A start pose from the GUI-step table equals pressing the reference's keys: rotate in 5-degree steps about the
    plot centroid, flip about it, shift in 0.5 m steps (trees.py:165-222 restated by oracle.pre_transform)."""
    from coregistrationgame_b200.matching import gui_hypothesis_table
    tab = gui_hypothesis_table(rot_steps=(-2, 0, 3), trans_steps=(-1, 0, 2), flips=(0, 1))
    assert tab.shape == (3 * 3 * 3 * 2, 6)
    np.testing.assert_array_equal(np.unique(tab[:, 4]), [-0.5, 0.0, 1.0])
    row = tab[np.flatnonzero((tab[:, 4] == 1.0) & (tab[:, 5] == -0.5))[5]]     # flip=1, rotation +15 deg
    np.testing.assert_allclose(row[:4], orc.hypothesis_matrix(15.0, 1).ravel(), atol=0)
