"""ctypes driver for the CPU thread emulator (tests only; see lct_emu.cpp)."""
import ctypes
import os
import subprocess

import numpy as np

from hiddenpose_b200 import operators as ops

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SO = os.environ.get("LCT_EMU_SO") or os.path.join(HERE, "liblct_emu.so")      # override + LCT_EMU_DEFS: variant builds
CSRC = os.path.join(ROOT, "hiddenpose_b200", "csrc")


def build(force=False):
    srcs = [os.path.join(HERE, "lct_emu.cpp")] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    if not force and os.path.exists(SO) and all(os.path.getmtime(SO) >= os.path.getmtime(s) for s in srcs):
        return SO
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-fPIC"] + os.environ.get("LCT_EMU_DEFS", "").split() + ["-shared", "-I" + CSRC, "-I" + cuda_inc,
                           "-o", SO, os.path.join(HERE, "lct_emu.cpp")])
    return SO


_lib = None


def lib(reverse=False):
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    _lib.lct_emu_init(1 if reverse else 0)
    return _lib


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


class EmuPlan:
    def __init__(self, N, M, bin_len, wall_size=2.0, method="lct", material="diffuse"):
        self.N, self.M = N, M
        rp, ci, v = ops.resampling_csr(M)
        self.csr = (np.ascontiguousarray(rp, np.int32), np.ascontiguousarray(ci, np.int32), np.ascontiguousarray(v, np.float32))
        self.falloff = np.ascontiguousarray(ops.falloff(M, material), np.float32)
        w = ops.inverse_filter_half(N, M, ops.slope_for(M, bin_len, wall_size), method)
        self.filt = np.ascontiguousarray((w * np.float32(1.0 / (8.0 * M * N * N))).astype(np.complex64))
        # symmetric layouts (Params::filt_sym): frequencies kh, kw <= N only (quarter), or kh <= N only for the
        # 512-point kernel (half rows)
        self.filt_quarter = np.ascontiguousarray(self.filt[:, :N + 1, :N + 1] if N < 256 else self.filt[:, :N + 1, :])
        # fused layout [kt][kw/2][plane row][kw&1], rows in the H plan's position order (what lct_plan_create builds)
        L = 2 * N
        kh = np.array([lib().lct_emu_plane_row_freq(N, r) for r in range(L)])
        self.filt_plane = (np.ascontiguousarray(self.filt[:, kh, :].reshape(M + 1, L, N, 2).transpose(0, 2, 1, 3))
                           if N <= 64 else None)

    def run(self, inp, D, Tin, be, backward=False, mask=31, reverse=False, s1=None, s2=None, fused=True, drift=0, sym=False):
        M, N = self.M, self.N
        inp = np.ascontiguousarray(inp, dtype=np.float32)
        C = inp.shape[0]
        out_T = Tin if backward else M
        out = np.full((C, out_T, N, N), np.nan, dtype=np.float32)
        if s1 is None:
            s1 = np.full((C, M + 1, N, N), np.nan, dtype=np.complex64)
        if s2 is None:
            s2 = np.full((C, M + 1, 2 * N, N), np.nan, dtype=np.complex64)
        s1 = np.ascontiguousarray(s1, dtype=np.complex64)
        s2 = np.ascontiguousarray(s2, dtype=np.complex64)
        be = np.asarray(be, dtype=np.int32)
        uniform = bool(np.all(be == be[0]))
        f, i32 = ctypes.c_float, ctypes.c_int
        lib().lct_emu_set_drift(int(drift))        # 1 / 2: one warp set runs whole runs of set-barrier phases before the other
        rc = lib(reverse).lct_emu_run(
            M, N, C, D, Tin, int(be[0]), None if uniform else _p(be, i32),
            _p(inp, f), _p(out, f), _p(s1.view(np.float32), f), _p(s2.view(np.float32), f),
            _p(self.csr[0], i32), _p(self.csr[1], i32), _p(self.csr[2], f), _p(self.falloff, f),
            _p((self.filt_quarter if sym else self.filt).view(np.float32), f),
            _p(self.filt_plane.view(np.float32), f) if (fused and self.filt_plane is not None) else None,
            int(backward), int(mask), (2 if self.N >= 256 else 1) if sym else 0)
        lib().lct_emu_set_drift(0)
        assert rc == 0, rc
        return out, s1, s2
