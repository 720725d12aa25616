"""Row H (hypothesis pre-transform) and row f4 (GUI-step hypotheses + write-back) pinned against the UNMODIFIED reference
domain model: golden vectors recorded from ``trees.Plot`` driven by the GUI's keys (tests/golden/make_golden_keys.py;
trees.py:165-222, :296-314, app.py:604-628).  CPU tests; the GPU leg runs the batched kernel from such a row."""
import os
import types

import numpy as np
import pytest

from oracle import ficp_oracle as orc

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "next_plot_keys.npz")
# The reference applies every key press incrementally (46 presses in case 6: a full turn in 5-degree steps) to coordinates
# of magnitude 6.5e6, one ulp = 9.3e-10 m; ONE rotation by the net angle differs from that by accumulated rounding only.
ATOL_M = 2e-8


def _cases():
    g = np.load(GOLDEN)
    return g, range(int(g["n"]))


def test_key_hypothesis_reproduces_the_reference_plot_edits():
    from coregistrationgame_b200.matching import key_hypothesis
    from coregistrationgame_b200.synthetic import apply_pose
    g, cases = _cases()
    worst = 0.0
    for c in cases:
        rot, flip, tx, ty = (int(v) for v in g[f"net_{c}"])
        row = key_hypothesis(rot, flip, tx, ty)
        got = apply_pose(g[f"orig_{c}"], row, g[f"centre0_{c}"])
        err = np.abs(got - g[f"cur_{c}"]).max()
        worst = max(worst, err)
        assert err < ATOL_M, (c, [str(k) for k in g[f"keys_{c}"]], err)
        assert bool(g[f"flipped_{c}"]) == bool(flip)
        # the oracle's pre-transform (what every batch parity test starts from) is the same map
        np.testing.assert_array_equal(orc.pre_transform(g[f"orig_{c}"], row, g[f"centre0_{c}"]), got)
        if len(g[f"keys_{c}"]) <= 1:      # a single key press: no accumulated rounding -> a few ulp at most
            assert err < 4e-9
    assert worst > 0.0     # the UTM cases do differ in the last bits: the tolerance above is not vacuous


def test_gui_table_rows_are_key_hypotheses():
    from coregistrationgame_b200.matching import gui_hypothesis_table, key_hypothesis
    tab = gui_hypothesis_table(rot_steps=(-7, 0, 5, 36), trans_steps=(-3, 0, 4), flips=(0, 1))
    k = 0
    for ty in (-3, 0, 4):
        for tx in (-3, 0, 4):
            for f in (0, 1):
                for r in (-7, 0, 5, 36):
                    np.testing.assert_array_equal(tab[k], key_hypothesis(r, f, tx, ty))
                    k += 1
    assert k == tab.shape[0]


class _Tree:
    def __init__(self, x, y):
        self.x, self.y, self.currentx, self.currenty = x, y, x, y


def _plot_like(xy):
    p = types.SimpleNamespace(trees=[_Tree(float(x), float(y)) for x, y in xy], center=(0, 0), current_center=(0, 0))
    return p


def test_write_back_equals_update_tree_positions():
    """matching.write_back against the state the reference's Plot held after update_tree_positions (bit for bit)."""
    from coregistrationgame_b200.matching import apply_registration, write_back
    g, cases = _cases()
    for c in cases:
        p = _plot_like(g[f"cur_{c}"])
        write_back(p, g[f"new_xy_{c}"])
        got = np.array([(t.currentx, t.currenty) for t in p.trees], dtype=float)
        np.testing.assert_array_equal(got, g[f"after_{c}"])
        np.testing.assert_array_equal(np.asarray(p.current_center, dtype=float), g[f"centre_after_{c}"])
        with pytest.raises(ValueError, match="does not match number of trees"):
            write_back(p, g[f"new_xy_{c}"][:-1])
    # a registration result [A | b] moves the plot like app.py:658-661 does with icp.source[:, :2]
    p = _plot_like(g["cur_1"])
    th = np.radians(12.0)
    T = np.array([[np.cos(th), -np.sin(th), 3.0], [np.sin(th), np.cos(th), -1.5]])
    apply_registration(p, T)
    want = g["cur_1"] @ T[:, :2].T + T[:, 2]
    np.testing.assert_array_equal(np.array([(t.currentx, t.currenty) for t in p.trees]), want)


def test_write_back_on_the_reference_plot_class_itself():
    """With the reference's own trees.py staged (oracle/_ref), the helper and Plot.update_tree_positions leave the SAME
    object state."""
    import importlib.util
    path = os.path.join(os.path.dirname(os.path.dirname(__file__)), "oracle", "_ref", "trees.py")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/trees.py not staged (tools/vendor_ref.sh)")
    from coregistrationgame_b200.matching import write_back
    spec = importlib.util.spec_from_file_location("trees_reference_unmodified", path)
    trees = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(trees)
    g, _ = _cases()
    a, b = trees.Plot(plotid=1), trees.Plot(plotid=2)
    for i, (x, y) in enumerate(g["orig_3"]):
        a.append_tree(trees.Tree(i, float(x), float(y)))
        b.append_tree(trees.Tree(i, float(x), float(y)))
    a.update_tree_positions(g["new_xy_3"])
    write_back(b, g["new_xy_3"])
    np.testing.assert_array_equal(a.get_tree_current_array(), b.get_tree_current_array())
    np.testing.assert_array_equal(np.asarray(a.current_center), np.asarray(b.current_center))
    Ra, ta, _ = a.get_transform()
    Rb, tb, _ = b.get_transform()
    np.testing.assert_array_equal(Ra, Rb)
    np.testing.assert_array_equal(ta, tb)


@pytest.mark.gpu
def test_batch_from_key_hypotheses_equals_icp_from_the_edited_plot():
    """GPU: the batched kernel started from key-hypothesis rows == the ICP started from the coordinates the reference's
    Plot holds after those keys (the oracle run on the golden `cur` arrays)."""
    from coregistrationgame_b200 import IcpBatch, TargetIndex
    from coregistrationgame_b200.batch import compose_world_transform
    from coregistrationgame_b200.matching import key_hypothesis
    g, cases = _cases()
    rng = np.random.default_rng(9)
    for c in (1, 3, 4, 6):
        orig, cur = g[f"orig_{c}"], g[f"cur_{c}"]
        lo, hi = cur.min(0) - 30, cur.max(0) + 30
        tgt = np.vstack([cur + rng.normal(0, 0.4, cur.shape), rng.uniform(lo, hi, (400, 2))])
        row = key_hypothesis(*(int(v) for v in g[f"net_{c}"]))
        ti = TargetIndex(tgt)
        b = IcpBatch(ti, [orig], row[None, :], centres=g[f"centre0_{c}"][None, :], min_k=0)
        out = b.run().results()
        ref = orc.ficp_run(cur, tgt, closed_form=True)
        A = compose_world_transform(out["hyp"][0, 0], b.centres[0])
        np.testing.assert_allclose(orig @ A[:, :2].T + A[:, 2], ref[:, :2], rtol=0, atol=1e-6)
        b.close()
        ti.close()
