"""Golden vectors for the hypothesis pre-transform and the write-back (8f rank 4), produced by the
UNMODIFIED reference domain model: ``Plot.rotate_plot`` (trees.py:201-211), ``Plot.coordinate_flip`` (:213-222),
``Plot.translate_plot`` (:186-199) driven exactly like the GUI keys drive them (``App.rotate_plot`` +-5 degrees,
``App.shift_plot`` 0.5 m, ``App.flip_plot``: app.py:604-628), and ``Plot.update_tree_positions`` (:296-314).

    python tests/golden/make_golden_keys.py        # build container only (the reference does not travel)

Every case stores the load-time coordinates, the key sequence, the net (rotation steps, flip, x steps, y steps), the
coordinates the reference holds after the keys, and - for the write-back - the plot state after
``update_tree_positions`` of a given (n, 2) array."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")
from trees import Plot, Tree  # noqa: E402

STEP_DEG, STEP_M = 5, 0.5      # app.py:618-624, app.py:36


def press(plot, key):
    """One key of the reference GUI (app.py:604-628)."""
    if key == "rl": plot.rotate_plot(STEP_DEG)
    elif key == "rr": plot.rotate_plot(-STEP_DEG)
    elif key == "f": plot.coordinate_flip()
    elif key == "up": plot.translate_plot((0, -STEP_M))
    elif key == "down": plot.translate_plot((0, STEP_M))
    elif key == "left": plot.translate_plot((-STEP_M, 0))
    elif key == "right": plot.translate_plot((STEP_M, 0))
    else: raise ValueError(key)


def net_of(keys):
    rot = sum({"rl": 1, "rr": -1}.get(k, 0) for k in keys)
    flip = sum(k == "f" for k in keys) % 2
    tx = sum({"right": 1, "left": -1}.get(k, 0) for k in keys)
    ty = sum({"down": 1, "up": -1}.get(k, 0) for k in keys)
    return rot, flip, tx, ty


def main():
    rng = np.random.default_rng(20261019)
    out = {}
    seqs = [
        ["rl"] * 3,                                            # +15 degrees
        ["rr"] * 7 + ["right"] * 4 + ["up"] * 2,               # -35 degrees, shift
        ["f"],                                                 # flip only
        ["f"] + ["rl"] * 5 + ["left"] * 3,                     # canonical order: flip, rotate, translate
        ["rl"] * 4 + ["f"] + ["rl"] * 2 + ["down"] * 5,        # flip in the middle of rotations
        ["right", "rl", "up", "f", "rr", "rr", "left", "f", "rl", "down", "f"],   # everything interleaved, 3 flips
        ["rl"] * 36 + ["right"] * 10,                          # a full turn in 5-degree steps (accumulated rounding)
        [],                                                    # untouched
    ]
    for c, keys in enumerate(seqs):
        n = int(rng.integers(4, 40))
        base = np.array([420100.0, 6483100.0]) if c % 2 == 0 else np.array([0.0, 0.0])
        xy = rng.uniform(-18, 18, (n, 2)) + base
        plot = Plot(plotid=c)
        for i, (x, y) in enumerate(xy):
            plot.append_tree(Tree(i, float(x), float(y)))
        centre0 = np.asarray(plot.current_center, dtype=float).copy()
        for k in keys:
            press(plot, k)
        cur = plot.get_tree_current_array()[:, 1:3].astype(float)
        rot, flip, tx, ty = net_of(keys)
        # write-back: what app.py:658-661 does with the ICP result
        new_xy = cur + rng.normal(0, 0.7, cur.shape)
        plot.update_tree_positions(new_xy)
        after = plot.get_tree_current_array()[:, 1:3].astype(float)
        out.update({f"orig_{c}": xy, f"centre0_{c}": centre0, f"keys_{c}": np.array(keys, dtype="U8"),
                    f"net_{c}": np.array([rot, flip, tx, ty], dtype=np.int64), f"cur_{c}": cur,
                    f"flipped_{c}": np.bool_(plot.flipped), f"new_xy_{c}": new_xy, f"after_{c}": after,
                    f"centre_after_{c}": np.asarray(plot.current_center, dtype=float)})
        print(c, "n", n, "keys", len(keys), "net", (rot, flip, tx, ty), "flipped", plot.flipped)
    out["n"] = np.int64(len(seqs))
    np.savez_compressed(os.path.join(HERE, "next_plot_keys.npz"), **out)


if __name__ == "__main__":
    main()
