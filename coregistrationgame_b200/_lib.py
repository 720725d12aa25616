"""ctypes binding of ``libficp_b200.so`` (C ABI declared in ``include/ficp_b200.h``).

There is deliberately no fallback: if the shared library is missing (not built) or no CUDA
device is present, every compute call raises.  Build with ``python __graft_entry__.py`` or
``make -C coregistrationgame_b200/csrc``.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FICP_B200_LIB") or os.path.join(_HERE, "libficp_b200.so")  # env override: A/B builds

c_i32, c_i64, c_f64, c_vp = C.c_int32, C.c_int64, C.c_double, C.c_void_p
P = C.POINTER


class TargetInfo(C.Structure):
    _fields_ = [("m", c_i64), ("has_z", c_i32), ("grid_w", c_i32), ("grid_h", c_i32), ("cell", c_f64),
                ("x0", c_f64), ("y0", c_f64), ("bbox", c_f64 * 4), ("build_ms", c_f64), ("clamped", c_i32),
                ("max_cell_pts", c_i32)]


class BatchParams(C.Structure):
    _fields_ = [("n_stages", c_i32), ("max_iterations", c_i32), ("allow_reflection", c_i32), ("min_k", c_i32),
                ("threshold", c_f64), ("window_margin", c_f64), ("warps_per_cta", c_i32), ("ctas_per_sm", c_i32),
                ("disable_window", c_i32), ("team_warps", c_i32), ("no_helpers", c_i32), ("trace_passes", c_i32),
                ("cta_per_icp", c_i32), ("reserved", c_i32)]


class BatchInfo(C.Structure):
    _fields_ = [("n_plots", c_i32), ("n_hyp", c_i32), ("n_hyp_local", c_i32), ("elems_per_lane", c_i32),
                ("match_z", c_i32), ("warps_per_cta", c_i32), ("ctas", c_i32), ("ctas_per_sm", c_i32),
                ("slices_per_plot", c_i32), ("window_pts_cap", c_i32), ("window_cells_cap", c_i32),
                ("team_warps", c_i32), ("helpers", c_i32), ("trace_passes", c_i32), ("smem_bytes", c_i64), ("rows", c_i64),
                ("trace_stride", c_i32), ("cta_per_icp", c_i32), ("rows_direct", c_i32), ("reserved", c_i32)]


# numpy view of ficp_hyp_result
HYP_RESULT_DTYPE = np.dtype([("m00", "<f8"), ("m01", "<f8"), ("m10", "<f8"), ("m11", "<f8"), ("cx", "<f8"),
                             ("cy", "<f8"), ("frmsd", "<f8"), ("rmse", "<f8"), ("k", "<i4"), ("passes", "<i4"),
                             ("flags", "<i4"), ("pad", "<i4")])

# every symbol include/ficp_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "ficp_last_error": (C.c_char_p, []),
    "ficp_device_count": (c_i32, [P(c_i32)]),
    "ficp_set_device": (c_i32, [c_i32]),
    "ficp_device_props": (c_i32, [P(c_i32), P(c_i64), P(c_i64), P(c_i32)]),
    "ficp_measure_l2_read_gbs": (c_i32, [c_i64, c_i32, P(c_f64)]),
    "ficp_target_create": (c_i32, [c_vp, c_i64, c_i32, c_i32, c_f64, c_vp, P(c_vp)]),
    "ficp_target_create_device": (c_i32, [c_vp, c_i64, c_i32, c_i32, c_f64, c_vp, P(c_vp)]),
    "ficp_target_get_info": (c_i32, [c_vp, P(TargetInfo)]),
    "ficp_target_destroy": (None, [c_vp]),
    "ficp_nn_query": (c_i32, [c_vp, c_vp, c_i64, c_i32, c_i32, c_vp, c_vp, c_vp]),
    "ficp_nn_query_device": (c_i32, [c_vp, c_vp, c_i64, c_i32, c_i32, c_vp, c_vp, c_vp]),
    "ficp_nn_query_ex": (c_i32, [c_vp, c_vp, c_i64, c_i32, c_i32, c_vp, c_vp, c_i32, c_vp, c_vp]),
    "ficp_nn_query_device_ex": (c_i32, [c_vp, c_vp, c_i64, c_i32, c_i32, c_vp, c_vp, c_i32, c_vp, c_vp]),
    "ficp_match_remove": (c_i32, [c_vp, c_vp, c_vp, c_i64, c_i32, c_i32, c_vp, c_vp, c_vp]),
    "ficp_radial_crop": (c_i32, [c_vp, c_f64, c_f64, c_f64, c_vp, c_vp]),
    "ficp_select_fraction": (c_i32, [c_vp, c_i32, c_vp, c_i32, c_vp, c_i64, c_i32, c_vp, c_i64, P(c_i64), P(c_f64), c_vp]),
    "ficp_fit_rigid2d": (c_i32, [c_vp, c_i32, c_vp, c_i32, c_i64, c_i32, c_vp]),
    "ficp_apply_xy": (c_i32, [c_vp, c_vp, c_i64, c_i32, c_vp]),
    "ficp_sumsq": (c_i32, [c_vp, c_i32, c_vp, c_i32, c_i64, c_i32, P(c_f64)]),
    "ficp_stepper_create": (c_i32, [c_vp, c_vp, c_i64, c_i32, c_i32, P(c_vp)]),
    "ficp_stepper_set_weights": (c_i32, [c_vp, c_vp]),
    "ficp_stepper_pass": (c_i32, [c_vp, c_i64, P(c_i64), P(c_f64)]),
    "ficp_stepper_fit_apply": (c_i32, [c_vp, c_i32, c_vp]),
    "ficp_stepper_read_xy": (c_i32, [c_vp, c_vp]),
    "ficp_stepper_destroy": (None, [c_vp]),
    "ficp_plot_centres": (c_i32, [c_vp, c_i32, c_vp, c_i64, c_vp]),
    "ficp_plot_geometry": (c_i32, [c_vp, c_i32, c_i32, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "ficp_batch_create": (c_i32, [c_vp, c_vp, c_i32, c_i32, c_vp, c_i64, c_vp, c_vp, c_i64, c_i32, c_i32, c_vp, c_vp,
                                  c_vp, c_i32, c_vp, P(BatchParams), c_i32, c_vp, P(c_vp)]),
    "ficp_batch_get_info": (c_i32, [c_vp, P(BatchInfo)]),
    "ficp_batch_run": (c_i32, [c_vp, c_vp]),
    "ficp_batch_results": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "ficp_batch_copy_best_keys_device": (c_i32, [c_vp, c_vp, c_vp]),
    "ficp_batch_pack_best_device": (c_i32, [c_vp, c_vp, c_vp]),
    "ficp_batch_best": (c_i32, [c_vp, c_vp, c_vp, c_vp]),
    "ficp_batch_trace": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "ficp_batch_destroy": (None, [c_vp]),
}

_lib = None


class FicpError(RuntimeError):
    pass


def _try_build():
    """The shared library is a build artefact (git-ignored).  If it is missing but the CUDA toolchain is present,
    compile it in-tree once (sm_100a, ~2 min) - this is the same `make` that `__graft_entry__.build()` runs."""
    import shutil
    import subprocess
    import sys
    nvcc = shutil.which("nvcc") or ("/usr/local/cuda/bin/nvcc" if os.path.exists("/usr/local/cuda/bin/nvcc") else None)
    if not nvcc or not shutil.which("make"):
        return
    print(f"[coregistrationgame_b200] {LIB_PATH} missing: building it with {nvcc} ...", file=sys.stderr)
    try:
        subprocess.check_call(["make", "-C", os.path.join(_HERE, "csrc"), "-j", str(min(8, os.cpu_count() or 1)), f"NVCC={nvcc}"],
                              stdout=subprocess.DEVNULL)
    except Exception as exc:  # the loader reports the missing library right after
        print(f"[coregistrationgame_b200] build failed: {exc}", file=sys.stderr)


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH) and not os.environ.get("FICP_B200_LIB"):
            _try_build()
        if not os.path.exists(LIB_PATH):
            raise FicpError(
                f"{LIB_PATH} not found: the CUDA extension is not built. Run `python __graft_entry__.py` "
                "(or `make -C coregistrationgame_b200/csrc`). There is no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def last_error():
    return load().ficp_last_error().decode("utf-8", "replace")


def check(rc, what=""):
    """Map a C status to the exception the reference's callers would see."""
    if rc == 0:
        return
    msg = last_error()
    if rc in (-1, -2):     # invalid argument / non-finite: the reference raises ValueError (numpy / scipy)
        raise ValueError(msg)
    if rc == -4:
        raise NotImplementedError(msg)
    raise FicpError(f"{what}: {msg} (status {rc})")


PACK_WORDS = 14   # FICP_PACK_WORDS: 8-byte words per plot record of ficp_batch_pack_best_device / ficp_batch_best


def ptr(a):
    """void* of a C-contiguous numpy array (or None)."""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(c_vp)


def device_count():
    n = c_i32(0)
    rc = load().ficp_device_count(C.byref(n))
    return n.value if rc == 0 else 0


def require_device():
    if device_count() < 1:
        raise FicpError("no CUDA device visible: coregistrationgame_b200 has no CPU fallback (" + last_error() + ")")


def device_props():
    sms, l2, smem, clk = c_i32(), c_i64(), c_i64(), c_i32()
    check(load().ficp_device_props(C.byref(sms), C.byref(l2), C.byref(smem), C.byref(clk)), "device_props")
    return {"sms": sms.value, "l2_bytes": l2.value, "smem_optin": smem.value, "clock_khz": clk.value}
