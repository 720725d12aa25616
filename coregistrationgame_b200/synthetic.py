"""Synthetic stands for tests and benchmarks: CHM tree tops uniform at `density` per m^2 with heights
U(5,35); each plot = the n CHM points nearest a centre, jittered, optionally with outlier trees / omitted CHM points,
then moved by a hidden rigid pose.  Pure input generation (numpy + a scipy kd-tree for picking each plot's trees) -
nothing here is on the registration path."""
from __future__ import annotations

import math

import numpy as np
from scipy.spatial import cKDTree

from .batch import hypothesis_matrix


def apply_pose(src, row, centre):
    """Start pose `row = [m00 m01 m10 m11 dx dy]` about `centre`, evaluated elementwise in the order the kernel uses:
    u = p - c ;  x' = (m00*ux + m01*uy) + (cx + dx) ;  y' = (m10*ux + m11*uy) + (cy + dy)."""
    m00, m01, m10, m11, dx, dy = [float(v) for v in row]
    out = src.copy()
    ux = src[:, 0] - centre[0]
    uy = src[:, 1] - centre[1]
    ox = centre[0] + dx
    oy = centre[1] + dy
    out[:, 0] = (m00 * ux + m01 * uy) + ox
    out[:, 1] = (m10 * ux + m11 * uy) + oy
    return out


def synthetic_scene(m, n, seed=0, density=0.05, dims=3, pos_noise=0.3, z_noise=1.0, out_frac=0.0,
                    omit_frac=0.0, n_plots=1, hidden_pose=True, dup_every=0, lattice_patch=0,
                    quantise=True):
    """Synthetic stand per SURVEY.md 8(d): CHM points uniform at `density` per m^2 with heights
    U(5,35); each plot = the n targets nearest a centre, jittered, optionally with outliers, then
    moved by a hidden rigid pose.  Coordinates are rounded to fp32-representable values
    (`quantise`) so fp32-staged and fp64 paths see identical inputs.

    Returns (target (M', dims), [plot sources (n, dims)], [hidden poses (theta_deg, dx, dy)])."""
    rng_t = np.random.default_rng(1000 + seed)
    rng_s = np.random.default_rng(2000 + seed)
    rng_p = np.random.default_rng(3000 + seed)
    side = math.sqrt(m / density)
    tgt = np.empty((m, 3))
    tgt[:, 0] = rng_t.uniform(0.0, side, m)
    tgt[:, 1] = rng_t.uniform(0.0, side, m)
    tgt[:, 2] = rng_t.uniform(5.0, 35.0, m)
    if lattice_patch:
        g = np.arange(lattice_patch, dtype=float)
        gx, gy = np.meshgrid(g, g)
        k = lattice_patch * lattice_patch
        tgt[:k, 0] = side / 2 + gx.ravel()
        tgt[:k, 1] = side / 2 + gy.ravel()
        tgt[:k, 2] = 20.0
    if quantise:
        tgt = tgt.astype(np.float32).astype(np.float64)
    tree = cKDTree(tgt[:, :2])
    plots, poses = [], []
    remove = []
    for p in range(n_plots):
        if n_plots == 1:
            c = np.array([side / 2, side / 2])
        else:
            c = rng_s.uniform(0.15 * side, 0.85 * side, 2)
        _, nbr = tree.query(c, k=n)
        nbr = np.atleast_1d(nbr)
        src = tgt[nbr].copy()
        src[:, :2] += rng_s.normal(0.0, pos_noise, (n, 2))
        src[:, 2] += rng_s.normal(0.0, z_noise, n)
        n_out = int(round(out_frac * n))
        if n_out:
            rad = np.sqrt(((tgt[nbr, :2] - c) ** 2).sum(1).max())
            who = rng_s.choice(n, n_out, replace=False)
            ang = rng_s.uniform(0, 2 * np.pi, n_out)
            rr = rad * np.sqrt(rng_s.uniform(0, 1, n_out))
            src[who, 0] = c[0] + rr * np.cos(ang)
            src[who, 1] = c[1] + rr * np.sin(ang)
            src[who, 2] = rng_s.uniform(5.0, 35.0, n_out)
        if omit_frac:
            n_om = int(round(omit_frac * n))
            remove.extend(rng_s.choice(nbr, n_om, replace=False).tolist())
        if hidden_pose:
            th = rng_p.uniform(-180.0, 180.0)
            d = rng_p.uniform(-5.0, 5.0, 2)
            cc = src[:, :2].mean(axis=0)
            mrow = np.concatenate([hypothesis_matrix(th, 0).ravel(), d])
            src = apply_pose(src, mrow, cc)
            poses.append((th, float(d[0]), float(d[1])))
        else:
            poses.append((0.0, 0.0, 0.0))
        if quantise:
            src = src.astype(np.float32).astype(np.float64)
        plots.append(src[:, :dims].copy())
    if remove:
        keep = np.ones(m, dtype=bool)
        keep[np.array(remove)] = False
        tgt = tgt[keep]
    if dup_every:
        dup = tgt[::dup_every].copy()
        tgt = np.vstack([tgt, dup])   # duplicates carry the HIGHER index
    return tgt[:, :dims].copy(), plots, poses
