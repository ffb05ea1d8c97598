"""ctypes binding of the C ABI declared in include/hiddenpose_lct.h.

The shared library is built in-tree (``hiddenpose_b200/libhiddenpose_lct.so``)
by :func:`build_native` / ``__graft_entry__.build()``.  There is no fallback:
if the library is missing or fails to load, every compute entry point raises.
"""
from __future__ import annotations

import ctypes
import os
import shutil
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
INCLUDE = os.path.join(ROOT, "include")
LIB_PATH = os.environ.get("HIDDENPOSE_LCT_LIB") or os.path.join(PKG_DIR, "libhiddenpose_lct.so")   # override: A/B builds

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O3", "-lineinfo",
    "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-shared",
]

# every symbol include/hiddenpose_lct.h declares
SYMBOLS = (
    "lct_abi_version", "lct_error_string", "lct_last_error", "lct_plan_create", "lct_plan_destroy",
    "lct_plan_time_bins", "lct_plan_spatial", "lct_plan_workspace_bytes", "lct_forward", "lct_backward",
    "lct_bp_laplacian", "lct_forward_host", "lct_run_staged", "lct_forward_minmax", "lct_minmax",
    "lct_normalize_feature", "lct_normalize_feature_backward",
    "lct_skip_workspace_bytes", "lct_skip_sum", "lct_skip_sum_backward",
)


class LctDesc(ctypes.Structure):
    _fields_ = [
        ("time_bins", ctypes.c_int32), ("spatial", ctypes.c_int32), ("device", ctypes.c_int32), ("reserved", ctypes.c_int32),
        ("mtx_rowptr", ctypes.POINTER(ctypes.c_int32)), ("mtx_colidx", ctypes.POINTER(ctypes.c_int32)),
        ("mtx_vals", ctypes.POINTER(ctypes.c_float)), ("falloff", ctypes.POINTER(ctypes.c_float)),
        ("filter_half", ctypes.POINTER(ctypes.c_float)),
        ("psf_z", ctypes.POINTER(ctypes.c_int32)), ("psf_yx", ctypes.POINTER(ctypes.c_int32)),
        ("psf_count", ctypes.c_int32), ("psf_value", ctypes.c_float), ("snr", ctypes.c_float), ("method_bp", ctypes.c_int32),
    ]


def _sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh"))] + \
           [os.path.join(INCLUDE, "hiddenpose_lct.h")]


def is_stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(s) > t for s in _sources())


def build_native(force=False, verbose=False):
    """Compile csrc/lct_api.cu for sm_100a into the in-tree shared library."""
    if not force and not is_stale():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "bin", "nvcc")
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libhiddenpose_lct.so")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
          ["-I" + INCLUDE, "-I" + CSRC, "-o", LIB_PATH, os.path.join(CSRC, "lct_api.cu")]
    subprocess.check_call(cmd)
    return LIB_PATH


_lib = None


def load():
    """Load the library (once) and declare the prototypes.  Raises if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: the CUDA library has not been built "
            "(run `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    vp, i32, sz = ctypes.c_void_p, ctypes.c_int32, ctypes.c_size_t
    pi32 = ctypes.POINTER(ctypes.c_int32)
    lib.lct_abi_version.restype = ctypes.c_int
    lib.lct_error_string.restype = ctypes.c_char_p
    lib.lct_error_string.argtypes = [ctypes.c_int]
    lib.lct_last_error.restype = ctypes.c_char_p
    lib.lct_plan_create.restype = ctypes.c_int
    lib.lct_plan_create.argtypes = [ctypes.POINTER(LctDesc), ctypes.POINTER(vp)]
    lib.lct_plan_destroy.restype = None
    lib.lct_plan_destroy.argtypes = [vp]
    lib.lct_plan_time_bins.restype = i32
    lib.lct_plan_time_bins.argtypes = [vp]
    lib.lct_plan_spatial.restype = i32
    lib.lct_plan_spatial.argtypes = [vp]
    lib.lct_plan_workspace_bytes.restype = sz
    lib.lct_plan_workspace_bytes.argtypes = [vp, i32]
    for fn in (lib.lct_forward, lib.lct_backward):
        fn.restype = ctypes.c_int
        fn.argtypes = [vp, vp, pi32, pi32, i32, i32, i32, vp, vp, sz, vp]
    lib.lct_run_staged.restype = ctypes.c_int
    lib.lct_run_staged.argtypes = [vp, vp, pi32, pi32, i32, i32, i32, vp, vp, sz, vp, i32, ctypes.POINTER(vp)]
    lib.lct_forward_minmax.restype = ctypes.c_int
    lib.lct_forward_minmax.argtypes = [vp, vp, pi32, pi32, i32, i32, i32, vp, vp, vp, sz, vp]
    lib.lct_minmax.restype = ctypes.c_int
    lib.lct_minmax.argtypes = [vp, i32, ctypes.c_int64, vp, vp]
    lib.lct_normalize_feature.restype = ctypes.c_int
    lib.lct_normalize_feature.argtypes = [vp, vp, vp, i32, ctypes.c_int64, ctypes.c_float, vp]
    lib.lct_normalize_feature_backward.restype = ctypes.c_int
    lib.lct_normalize_feature_backward.argtypes = [vp, vp, vp, vp, vp, i32, ctypes.c_int64, ctypes.c_float, vp]
    lib.lct_bp_laplacian.restype = ctypes.c_int
    lib.lct_bp_laplacian.argtypes = [vp, vp, vp, i32, ctypes.POINTER(ctypes.c_float), i32, vp]
    lib.lct_skip_workspace_bytes.restype = sz
    lib.lct_skip_workspace_bytes.argtypes = [i32, i32, i32]
    lib.lct_skip_sum.restype = ctypes.c_int
    lib.lct_skip_sum.argtypes = [vp, vp, vp, i32, i32, i32, i32, vp, vp]
    lib.lct_skip_sum_backward.restype = ctypes.c_int
    lib.lct_skip_sum_backward.argtypes = [vp, vp, vp, i32, i32, i32, i32, vp, vp, vp, sz, vp]
    lib.lct_forward_host.restype = ctypes.c_int
    lib.lct_forward_host.argtypes = [vp, vp, pi32, pi32, i32, i32, i32, vp, vp]
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        lib = load()
        raise RuntimeError(f"hiddenpose_lct: {lib.lct_error_string(rc).decode()} ({lib.lct_last_error().decode()})")
