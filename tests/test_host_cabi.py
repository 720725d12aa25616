"""CPU-only checks of the host side: the C-ABI library loads and exports every symbol the header
declares, argument validation mirrors the reference, and the host helpers agree with the oracle.
No compute call is made here (there is no GPU in the build container and no CPU fallback)."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from oracle import ficp_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "ficp_b200.h")).read()
    return sorted(set(re.findall(r"FICP_API\s+[\w\s\*]+?\b(ficp_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from coregistrationgame_b200 import _lib
    lib = _lib.load()
    declared = _header_symbols()
    assert len(declared) >= 20
    assert sorted(_lib.SIGNATURES) == declared, "ctypes table and include/ficp_b200.h disagree"
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in the header but not exported"
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r"\bT (ficp_\w+)", out)))
    assert exported == declared, "exported symbols differ from the header"


def test_abi_struct_sizes():
    import ctypes as C
    from coregistrationgame_b200 import _lib
    assert _lib.HYP_RESULT_DTYPE.itemsize == 80
    assert C.sizeof(_lib.BatchParams) == 64
    assert C.sizeof(_lib.BatchInfo) == 88
    assert C.sizeof(_lib.TargetInfo) == 96


def test_no_device_fails_loudly():
    from coregistrationgame_b200 import _lib, TargetIndex
    if _lib.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(_lib.FicpError):
        _lib.require_device()
    with pytest.raises(_lib.FicpError):
        TargetIndex(np.zeros((4, 2)))
    from ficp import FractionalICP
    icp = FractionalICP(np.random.default_rng(0).normal(size=(5, 3)), np.random.default_rng(1).normal(size=(6, 3)))
    with pytest.raises(_lib.FicpError):          # no silent CPU path
        icp.run()


def test_constructor_and_degenerate_inputs_without_gpu():
    """ficp.py:34-44 and the empty-input conventions (ficp.py:66-68,76-77,125-126) need no device."""
    from ficp import FractionalICP
    with pytest.raises(ValueError, match="2D arrays"):
        FractionalICP(np.zeros(3), np.zeros((3, 2)))
    icp = FractionalICP([[0, 0, 1], [1, 1, 2]], [[0.0, 0.0], [1.0, 1.0]])
    assert icp.match_dims == 2 and icp.source.dtype == np.float64
    assert (icp.lambda_val, icp.threshold, icp.max_iterations, icp.allow_reflection) == (3.0, 1e-6, 1000, False)
    e = FractionalICP(np.empty((0, 3)), np.ones((4, 3)))
    assert e.run().shape == (0, 3) and e.lambda_val == 0.95
    src = np.arange(12, dtype=float).reshape(4, 3)
    t = FractionalICP(src, np.empty((0, 3)))
    corr, dist = t.find_correspondences(src, np.empty((0, 3)))
    assert corr.shape == (0, 3) and dist.size == 0 and t.find_optimal_fraction(corr, dist) == (0.0, 0)
    np.testing.assert_array_equal(t.run(), src)
    assert t.frmsd(0.5, 0, src, src) == float("inf")


def test_oversized_plot_is_refused_before_any_work():
    """plots above the stage kernels' row limit (2^24 since round 2; 8192 before) raise up front - no NN pass
    first.  Checked on the limit constant: allocating 2^24 + 1 rows here would only test numpy."""
    from coregistrationgame_b200 import ficp as shim
    assert shim._STEPWISE_MAX_N == 1 << 24 and shim._KERNEL_MAX_N == 1024
    src = open(shim.__file__).read()
    assert src.index("if n > _STEPWISE_MAX_N") < src.index("if n > _KERNEL_MAX_N") < src.index("IcpBatch(index")


def test_plot_centres_vectorised_is_bit_identical():
    """batch.plot_centres reduces plots of equal size together; it must return the bits of the per-plot
    `rows[:, :2].mean(axis=0)` the oracle (and any reference-side caller) computes - start poses depend on them."""
    from coregistrationgame_b200.batch import plot_centres
    rng = np.random.default_rng(1)
    for ld in (2, 3, 5):
        sizes = np.r_[rng.integers(1, 40, 30), [63, 64, 65, 127, 128, 129, 150, 150, 150, 255, 256, 257, 500, 512, 513, 1000, 1024]]
        for same in (False, True):
            sz = np.full(40, 150) if same else sizes
            srcs = [rng.normal(size=(int(n), ld)) * 1000 + 6.4e6 for n in sz]
            src = np.ascontiguousarray(np.vstack(srcs))
            off = np.concatenate([[0], np.cumsum(sz)]).astype(np.int64)
            want = np.array([a[:, :2].mean(axis=0) for a in srcs])
            np.testing.assert_array_equal(plot_centres(src, off), want)


def _plot_geometry(src, off, cen, use_z):
    import ctypes as C
    from coregistrationgame_b200 import _lib
    rows, n_plots = src.shape[0], off.shape[0] - 1
    u, z = np.full((rows, 2), np.nan), np.full(rows, np.nan)
    ubar, rho = np.full((n_plots, 2), np.nan), np.full(n_plots, np.nan)
    rc = _lib.load().ficp_plot_geometry(_lib.ptr(src), src.shape[1], int(use_z), _lib.ptr(off), n_plots, _lib.ptr(cen),
                                        _lib.ptr(u), _lib.ptr(z) if use_z else None, _lib.ptr(ubar), _lib.ptr(rho))
    return rc, u, z, ubar, rho


@pytest.mark.parametrize("threads", ["1", "3", "8"])
def test_plot_geometry_pass_of_batch_create(threads, monkeypatch):
    """The host pass ficp_batch_create makes over the rows (csrc/batch_prep.h; a few threads over the plots): local
    coordinates are the oracle's single subtraction `row - centre` bit for bit (oracle.pre_transform), the shift point is
    the in-order mean of them, and the radius covers every tree - with any number of host threads, for ragged plots."""
    monkeypatch.setenv("FICP_HOST_THREADS", threads)
    from coregistrationgame_b200.batch import plot_centres
    rng = np.random.default_rng(int(threads))
    for ld, use_z, sizes in ((3, True, np.full(1250, 150)), (2, False, rng.integers(1, 300, 900)), (5, True, rng.integers(1, 1025, 400)),
                             (3, True, np.array([1])), (3, False, np.array([1, 1024, 1, 7]))):
        srcs = [rng.normal(size=(int(n), ld)) * 300 + np.r_[5.3e5, 6.48e6, np.zeros(ld - 2)] for n in sizes]
        src = np.ascontiguousarray(np.vstack(srcs))
        off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
        cen = plot_centres(src, off)
        np.testing.assert_array_equal(cen, np.array([a[:, :2].mean(axis=0) for a in srcs]))
        rc, u, z, ubar, rho = _plot_geometry(src, off, cen, use_z)
        assert rc == 0
        np.testing.assert_array_equal(u, src[:, :2] - np.repeat(cen, sizes, axis=0))
        if use_z:
            np.testing.assert_array_equal(z, src[:, 2])
        for p in (0, len(sizes) // 2, len(sizes) - 1):
            up = u[off[p]:off[p + 1]]
            sx = sy = 0.0
            for a, b in up:
                sx += a
                sy += b
            assert (ubar[p, 0], ubar[p, 1]) == (sx / len(up), sy / len(up))
        # the read-only form (u_out = NULL: rows go to the device as they are) returns the same per-plot values
        ubar2, rho2 = np.full_like(ubar, np.nan), np.full_like(rho, np.nan)
        from coregistrationgame_b200 import _lib
        assert _lib.load().ficp_plot_geometry(_lib.ptr(src), src.shape[1], int(use_z), _lib.ptr(off), len(sizes), _lib.ptr(cen),
                                              None, None, _lib.ptr(ubar2), _lib.ptr(rho2)) == 0
        np.testing.assert_array_equal(ubar2, ubar)
        np.testing.assert_array_equal(rho2, rho)
        far = np.array([np.hypot(*(u[off[p]:off[p + 1]] - ubar[p]).T).max() for p in range(len(sizes))])
        assert (rho >= far).all() and (rho <= far * (1 + 1e-14) + 2.1e-6 * (far + 1)).all()
        # a caller's centre far from the plot (FractionalICP.run turns about the origin): the radius is still tight
        zero = np.zeros_like(cen)
        rc, u0, _, ubar0, rho0 = _plot_geometry(src, off, zero, use_z)
        assert rc == 0
        np.testing.assert_array_equal(u0, src[:, :2])
        far0 = np.array([np.hypot(*(u0[off[p]:off[p + 1]] - ubar0[p]).T).max() for p in range(len(sizes))])
        assert (rho0 >= far0).all() and (rho0 <= far0 * (1 + 1e-14) + 1e-300).all()
    # finite coordinates whose SUM overflows are still finite input (the rows decide, not the sums)
    big = np.full((4, 2), 1.7e308)
    assert _plot_geometry(big, np.array([0, 4]), np.zeros((1, 2)), False)[0] == 0
    # a non-finite matched coordinate anywhere -> status -2 ('x' must be finite, ficp.py:70 via scipy); unmatched columns may hold anything
    src = rng.normal(size=(70000, 4))
    off = np.arange(0, 70001, 100).astype(np.int64)
    cen = plot_centres(src, off)
    src[:, 3] = np.nan
    assert _plot_geometry(src, off, cen, True)[0] == 0
    for (r, c), use_z, want in (((69999, 2), True, -2), ((69999, 2), False, 0), ((12345, 0), False, -2), ((40000, 1), True, -2)):
        bad = src.copy()
        bad[r, c] = np.inf if c else np.nan
        assert _plot_geometry(bad, off, cen, use_z)[0] == want


def test_stack_plots_forms_and_errors():
    """The three input forms of register_batch / IcpBatch give the same C-contiguous stack - also for Fortran-ordered,
    sliced, integer and list-of-lists plots (np.concatenate keeps Fortran order: the real Data/2014 arrays come that way)."""
    from coregistrationgame_b200.batch import stack_plots
    rng = np.random.default_rng(5)
    base = [rng.normal(size=(n, 3)) for n in (4, 1, 9)]
    want = np.vstack(base)
    for form in (base, [np.asfortranarray(a) for a in base], [np.hstack([a, a])[:, :3] for a in base], [a.tolist() for a in base],
                 (want, np.array([0, 4, 5, 14])), (np.asfortranarray(want), [0, 4, 5, 14])):
        src, off, sizes = stack_plots(form)
        assert src.flags["C_CONTIGUOUS"] and src.dtype == np.float64 and off.dtype == np.int64
        np.testing.assert_array_equal(src, want)
        np.testing.assert_array_equal(off, [0, 4, 5, 14])
        np.testing.assert_array_equal(sizes, [4, 1, 9])
    src, off, sizes = stack_plots(np.arange(6).reshape(3, 2))          # one plot, integer input
    assert src.dtype == np.float64 and off.tolist() == [0, 3] and sizes.tolist() == [3]
    for bad, msg in (([], "non-empty"), ([base[0], np.empty((0, 3))], "non-empty"), ([base[0], np.zeros(3)], "non-empty"),
                     ([base[0], np.zeros((2, 2))], "same number of columns"), ((want, np.array([0, 4, 4, 14])), "offsets"),
                     ((want, np.array([1, 14])), "offsets"), ((want, np.array([0, 13])), "offsets")):
        with pytest.raises(ValueError, match=msg):
            stack_plots(bad)


def test_host_helpers_match_oracle():
    from coregistrationgame_b200 import batch
    np.testing.assert_array_equal(batch.hypothesis_table(16, (0, 1), batch.translation_lattice(3, 2.5)),
                                  orc.hypothesis_table(16, (0, 1), orc.translation_lattice(3, 2.5)))
    for n, lam in ((1, 3.0), (17, 1.3), (500, 0.95)):
        np.testing.assert_array_equal(batch.frmsd_weights(n, lam), orc.frmsd_weights(n, lam))
    for n in (1, 9, 150, 500):
        for f in (0.5, 0.6, 0.7, 0.8, 0.9, 0.95, 1.0):
            assert batch.fixed_fraction_k(n, f) == orc.fixed_fraction_k(n, f)
    keys = np.array([orc.pack_best_key(0.25, 7), orc.pack_best_key(np.inf, 0), orc.pack_best_key(0.0, 4095)], dtype=np.uint64)
    dec = batch.decode_best_keys(keys)
    np.testing.assert_array_equal(dec["best_hyp"], [7, 0, 4095])
    np.testing.assert_array_equal(dec["best_score"], [0.25, np.inf, 0.0])
    # key order = (score, id) order
    assert orc.pack_best_key(0.25, 9) > orc.pack_best_key(0.25, 7) > orc.pack_best_key(0.2499, 4000)


def test_product_path_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "coregistrationgame_b200")):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("the oracle", "").replace("as the oracle", ""), f
    assert "oracle" not in open(os.path.join(ROOT, "ficp.py")).read()


@pytest.mark.parametrize("flags", [[], ["-DFICP_TIETEST_STREAM"], ["-DFICP_PRESCAN_OWN_CELL"],
                                   ["-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=all"]],
                         ids=["product", "tietest-stream-variant", "prescan-experiment", "product-under-asan-ubsan"])
def test_nn_search_host_check(tmp_path, flags):
    """The grid NN search (ring/termination/tie logic, window + global accessors, streamed form with arbitrary
    seeds, runner-up and lower bound of the tracked form) is host-compilable: build it with g++ and compare 19 200
    queries / 67 430 tracked searches (global and window accessor) against brute force.  Also for the experimental build flags kept in the
    source, and once under AddressSanitizer + UBSan (every window / cell-table / candidate index of the search stays in range)."""
    exe = tmp_path / "nn_check"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", *flags, "-I", os.path.join(ROOT, "coregistrationgame_b200", "csrc"),
                           "-I", "/usr/local/cuda/include", os.path.join(ROOT, "tests", "hostcheck", "nn_search_check.cpp"),
                           "-o", str(exe)])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout[-2000:]
    assert "mismatches=0" in out.stdout


@pytest.mark.parametrize("sanitizer", ["address,undefined", "thread"])
def test_batch_prep_host_check_under_sanitizers(tmp_path, sanitizer):
    """csrc/batch_prep.h - the host walk over the plot rows that ficp_batch_create makes, incl. its thread fan-out and the
    slice addressing of the staging route - compiled with g++ under ASan + UBSan and under TSan and compared with a plain
    serial restatement (bit-exact centres, u, ubar; covering radius; non-finite rows refused) for 1 / 2 / 3 / 8 threads."""
    exe = tmp_path / "batch_prep_check"
    cmd = ["g++", "-O1", "-g", "-std=c++17", f"-fsanitize={sanitizer}", "-fno-sanitize-recover=all", "-pthread",
           os.path.join(ROOT, "tests", "hostcheck", "batch_prep_check.cpp"), "-o", str(exe)]
    built = subprocess.run(cmd, capture_output=True, text=True)
    if built.returncode != 0 and "sanitize" in built.stderr:
        pytest.skip(f"this toolchain has no -fsanitize={sanitizer} runtime")
    assert built.returncode == 0, built.stderr[-2000:]
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    if sanitizer == "thread" and "FATAL: ThreadSanitizer" in out.stderr and "mmap" in out.stderr:
        pytest.skip("ThreadSanitizer cannot map its shadow memory in this container")
    assert out.returncode == 0, (out.stdout + out.stderr)[-2000:]
    assert "cases ok" in out.stdout


def test_bench_cli_parses_without_a_gpu():
    """`bench.py --help` must render (a stray '%' in a help string once broke argparse) and `--impl reference` must be
    importable on a box without CUDA (the reference arm never touches the GPU)."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--help"], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr[-800:]
    for flag in ("--gpus", "--steps", "--warmup", "--impl", "--workload"):
        assert flag in out.stdout
