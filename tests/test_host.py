"""CPU-side tests: host operators, reference-facing API surface, the C-ABI library's
symbols, and the kernels' logic through the CPU thread emulator (tests/emu)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from hiddenpose_b200 import _native, operators as ops
from oracle import lct_oracle as O
from tests._golden import Case, constants

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---- operators ---------------------------------------------------------------------
@pytest.mark.parametrize("M", [16, 32, 64, 128, 256, 512])
def test_resampling_csr_is_the_reference_operator(M):
    c = constants()
    rp, ci, v = ops.resampling_csr(M)
    ref = np.zeros((M, M), np.float32)
    ref[c[f"mtx{M}_rows"], c[f"mtx{M}_cols"]] = c[f"mtx{M}_vals"]
    assert np.array_equal(ops.csr_to_dense(rp, ci, v, M), ref)
    # structure the kernels rely on: contiguous band per row, <= 3 per column
    for i in range(M):
        cols = ci[rp[i]:rp[i + 1]]
        assert np.array_equal(cols, np.arange(cols[0], cols[0] + len(cols)))
        assert cols[0] == int(np.ceil(np.sqrt(i * M + 1))) - 1
    trp, tci, tv = ops.csr_transpose(rp, ci, v, M)
    assert np.diff(trp).max() <= 3
    assert np.array_equal(ops.csr_to_dense(trp, tci, tv, M), ref.T)


@pytest.mark.parametrize("N,M", [(8, 32), (16, 64), (32, 128), (64, 256)])
def test_psf_support_is_the_reference_psf(N, M):
    c = constants()
    z, y, x, val = ops.psf_support(N, M, float(c[f"psf_n{N}m{M}_slope"]))
    got = set(map(tuple, np.stack([z, y, x], 1).tolist()))
    ref = set(map(tuple, c[f"psf_n{N}m{M}_zyx"].tolist()))
    assert got == ref
    assert np.all(c[f"psf_n{N}m{M}_vals"] == val)


@pytest.mark.parametrize("name", ["m64n16_full", "m128n32_window", "m64n16_bp"])
def test_filter_half_matches_reference_samples(name):
    """Half spectrum + Hermitian mirror reproduces the reference's invpsf_real/imag."""
    import hiddenpose_b200 as hp
    c = Case(name)
    layer = hp.lct(spatial=c.N, crop=c.M, bin_len=c.bin_len, method=c.method)
    from tests.golden.make_golden import sample_idx
    idx = sample_idx(8 * c.M * c.N * c.N)
    re, im = layer.invpsf_real.numpy().ravel()[idx], layer.invpsf_imag.numpy().ravel()[idx]
    r0, i0 = c.z["invpsf_real_sample"], c.z["invpsf_imag_sample"]
    den = np.linalg.norm(r0) + np.linalg.norm(i0)
    assert (np.linalg.norm(re - r0) + np.linalg.norm(im - i0)) / den < 1e-6


def test_laplacian_matches_reference():
    assert np.array_equal(ops.laplacian_filter(), constants()["laplacian"])


# ---- API surface -------------------------------------------------------------------
def test_module_surface_matches_reference():
    import hiddenpose_b200 as hp
    l = hp.lct(spatial=16, crop=64, bin_len=0.08, wall_size=2.0, method="lct", material="diffuse")
    assert (l.spatial_grid, l.crop, l.bin_len, l.wall_size, l.method, l.material) == (16, 64, 0.08, 2.0, "lct", "diffuse")
    assert l.snr == 0.1 and abs(l.trange - 5.12) < 1e-12 and l.c == 3e8
    assert l.mtx_MxM.shape == (64, 64) and torch.equal(l.mtx_MxM.T, l.mtxi_MxM)
    assert l.invpsf_real.shape == (1, 128, 32, 32) and l.gridz_1xMx1x1.shape == (1, 64, 1, 1)
    assert dict(l.state_dict()) == {} and list(l.parameters()) == [] and list(l.buffers()) == []
    l.to("cpu")                                   # harmless, like the reference
    with pytest.raises(AssertionError):
        hp.lct(spatial=16, crop=48)               # tflct.py:20
    fp = hp.FeaturePropagation(image_size=16, time_size=64, bin_len=0.08, dnum=1, dev="cpu")
    assert isinstance(fp.method, hp.LCT) and fp.method.time_size == 64 and fp.method.image_size == 16
    with pytest.raises(AssertionError):
        hp.FeaturePropagation(mode="bp")          # feature_propagation.py:37-38


def test_no_cpu_fallback():
    import hiddenpose_b200 as hp
    l = hp.lct(spatial=8, crop=32, bin_len=0.16)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        l(torch.zeros(1, 2, 32, 8, 8), [0], [32])


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "hiddenpose_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+(oracle|tests)\b", src, re.M), f


def test_normalize_feature_keeps_negative_minimum():
    import hiddenpose_b200 as hp
    x = torch.tensor([-2.0, 0.0, 2.0]).view(1, 1, 3, 1, 1)
    assert torch.allclose(hp.normalize_feature(x).flatten(), torch.tensor([0.0, 5.0, 10.0]))


# ---- C ABI -------------------------------------------------------------------------
def test_shared_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "hiddenpose_lct.h")).read()
    declared = set(re.findall(r"\b(lct_[a-z_]+)\s*\(", header))
    assert declared == set(_native.SYMBOLS)
    _native.build_native()
    lib = ctypes.CDLL(_native.LIB_PATH)
    for sym in declared:
        assert hasattr(lib, sym), sym
    lib.lct_abi_version.restype = ctypes.c_int
    assert lib.lct_abi_version() == 3
    lib.lct_error_string.restype = ctypes.c_char_p
    assert lib.lct_error_string(2).startswith(b"unsupported")


# ---- kernel logic through the CPU thread emulator -----------------------------------
@pytest.mark.parametrize("name", ["m32n8_full", "m64n16_full", "m64n16_window", "m64n16_specular", "m512n16_window", "m32n256_window"])
def test_emulated_kernels_match_reference(name):
    """The exact device code, stepped thread by thread on the CPU, vs the reference's outputs."""
    from tests.emu.emu import EmuPlan
    c = Case(name)
    plan = EmuPlan(c.N, c.M, c.bin_len, 2.0, c.method, c.material)
    C = c.B * c.D
    y, _, _ = plan.run(c.x.reshape(C, c.tin, c.N, c.N), c.D, c.tin, c.tbes)
    assert c.y_err(y.reshape(c.B, c.D, c.M, c.N, c.N)) <= 1e-5
    gx, _, _ = plan.run(c.g.reshape(C, c.M, c.N, c.N), c.D, c.tin, c.tbes, backward=True)
    assert c.gx_err(gx.reshape(c.B, c.D, c.tin, c.N, c.N)) <= 1e-4


@pytest.mark.parametrize("name,N,M", [("m32n256_window", 256, 32), (None, 16, 32), (None, 64, 32)])
def test_emulated_quarter_filter_matches_full(name, N, M):
    """The filter's mirror symmetry in kh and kw (Params::filt_sym): the column-filter kernels reading the stored
    quarter times a twiddle give what they give reading the whole half spectrum -- forward and backward, the
    parity-split kernel (N = 256) and the register-cached one -- and the reference's own output where a golden exists."""
    from tests.emu.emu import EmuPlan
    if name:
        c = Case(name)
        plan = EmuPlan(c.N, c.M, c.bin_len)
        x, g, tin, tbes = c.x.reshape(c.B * c.D, c.tin, N, N), c.g.reshape(c.B * c.D, M, N, N), c.tin, c.tbes
    else:
        plan = EmuPlan(N, M, 0.16)
        rs = np.random.RandomState(N)
        tin, tbes = M - 4, [3]
        x, g = rs.rand(1, tin, N, N).astype(np.float32), rs.randn(1, M, N, N).astype(np.float32)
    full = plan.run(x, 1, tin, tbes, fused=False)[0]
    quarter = plan.run(x, 1, tin, tbes, fused=False, sym=True)[0]
    assert O.rel_l2(quarter, full) <= 1e-6
    gfull = plan.run(g, 1, tin, tbes, backward=True, fused=False)[0]
    gquarter = plan.run(g, 1, tin, tbes, backward=True, fused=False, sym=True)[0]
    assert O.rel_l2(gquarter, gfull) <= 1e-6
    assert np.array_equal(quarter, plan.run(x, 1, tin, tbes, fused=False, sym=True, reverse=True)[0])
    if name:
        assert c.y_err(quarter.reshape(c.B, c.D, M, N, N)) <= 1e-5
        assert c.gx_err(gquarter.reshape(c.B, c.D, tin, N, N)) <= 1e-4


def test_emulated_kernels_have_no_intra_phase_races():
    """Running the threads of each barrier-delimited phase in reverse order must not change a bit."""
    from tests.emu.emu import EmuPlan
    plan = EmuPlan(8, 32, 0.16)
    x = np.random.RandomState(1).rand(3, 25, 8, 8).astype(np.float32)
    a, _, _ = plan.run(x, 1, 25, [0, 3, 7])
    b, _, _ = plan.run(x, 1, 25, [0, 3, 7], reverse=True)
    assert np.array_equal(a, b)
    orc = O.LctOracle(8, 32, 0.16)
    yo = orc.forward(torch.from_numpy(x).view(3, 1, 25, 8, 8), [0, 3, 7], [25, 28, 32]).numpy().reshape(3, 32, 8, 8)
    assert O.rel_l2(a, yo) <= 1e-5


def test_emulated_parity_split_row_kernels():
    """N = 256 runs the H-axis passes as two 256-point transforms by output parity (RowFwdSplit / RowInvSplit):
    each stage alone against numpy's FFT, and bit-identical with the threads of every phase run backwards."""
    from tests.emu.emu import EmuPlan
    N, M = 256, 32
    plan = EmuPlan(N, M, 0.16)
    rs = np.random.RandomState(7)
    x = np.zeros((1, M, N, N), np.float32)
    s1 = (rs.randn(1, M + 1, N, N) + 1j * rs.randn(1, M + 1, N, N)).astype(np.complex64)
    s2 = (rs.randn(1, M + 1, 2 * N, N) + 1j * rs.randn(1, M + 1, 2 * N, N)).astype(np.complex64)
    fwd = plan.run(x, 1, M, [0], mask=2, s1=s1.copy())[2]
    ref = np.fft.fft(s1.astype(np.complex128), n=2 * N, axis=2)
    assert np.linalg.norm(fwd - ref) <= 1e-6 * np.linalg.norm(ref)
    inv = plan.run(x, 1, M, [0], mask=8, s2=s2.copy())[1]
    ref = (np.fft.ifft(s2.astype(np.complex128), axis=2) * (2 * N))[:, :, :N]
    assert np.linalg.norm(inv - ref) <= 1e-6 * np.linalg.norm(ref)
    assert np.array_equal(fwd, plan.run(x, 1, M, [0], mask=2, s1=s1.copy(), reverse=True)[2])
    assert np.array_equal(inv, plan.run(x, 1, M, [0], mask=8, s2=s2.copy(), reverse=True)[1])


@pytest.mark.parametrize("N", [32, 64])
def test_emulated_set_barriers_tolerate_maximal_drift(N):
    """The plane kernel synchronises its two warp sets on named 256-thread barriers between the phases that only
    involve one set (PlaneFilter::group_phase).  The emulator's drift mode runs one set through every run of such
    phases before the other starts: the largest skew the hardware could produce must not change a bit (a wrongly
    declared set barrier shows up as NaN from the poisoned shared memory)."""
    from tests.emu.emu import EmuPlan
    M = 32
    plan = EmuPlan(N, M, 0.16)
    x = np.random.RandomState(3).rand(2, M, N, N).astype(np.float32)
    base = plan.run(x, 1, M, [0, 0])[0]
    assert not np.isnan(base).any()
    for drift in (1, 2):
        assert np.array_equal(base, plan.run(x, 1, M, [0, 0], drift=drift)[0])
        assert np.array_equal(base, plan.run(x, 1, M, [0, 0], drift=drift, reverse=True)[0])


def test_emulated_measured_and_rejected_variants_still_compute_the_same():
    """The kernel variants that were measured and left off (persistent plane kernel with the next plane
    staged by a bulk copy, K5's tile through per-row bulk copies, K5's tile-walk exit test in the kernel driver) stay
    compilable and correct: a second emulator build with them switched on, in a subprocess, must reproduce the default
    build bit for bit (any thread order, maximal set drift)."""
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    code = (
        "import sys, numpy as np\n"
        "from tests.emu.emu import EmuPlan\n"
        "out = {}\n"
        "for N, M, C in ((64, 32, 3), (16, 64, 2)):\n"
        "    plan = EmuPlan(N, M, 0.16)\n"
        "    x = np.random.RandomState(N).rand(C, M - 3, N, N).astype(np.float32)\n"
        "    y = plan.run(x, 1, M - 3, [2] * C)[0]\n"
        "    assert np.array_equal(y, plan.run(x, 1, M - 3, [2] * C, reverse=True)[0])\n"
        "    if N == 64:\n"
        "        assert np.array_equal(y, plan.run(x, 1, M - 3, [2] * C, drift=1)[0])\n"
        "    out['y%d' % N] = y\n"
        "np.savez(sys.argv[1], **out)\n")
    results = []
    for tag, defs in (("default", ""), ("variants", "-DLCT_PLANE_PERSIST=1 -DLCT_TIME_INV_BULK=1 -DLCT_TIME_INV_DRIVER_EXIT=1")):
        env = dict(os.environ, LCT_EMU_DEFS=defs)
        if defs:
            env["LCT_EMU_SO"] = os.path.join(here, "emu", "liblct_emu_variants.so")
        path = os.path.join(here, "emu", f"_variant_check_{tag}.npz")
        subprocess.run([sys.executable, "-c", code, path], check=True, env=env, cwd=os.path.dirname(here))
        results.append(dict(np.load(path)))
        os.remove(path)
    for k in results[0]:
        assert np.array_equal(results[0][k], results[1][k]), k


def test_operator_structure_is_validated():
    """The kernels look for rows longer than three taps only among the first few rows of mtx and never in
    mtx^T (lct_tables.h); an operator that breaks this must be refused, not silently truncated."""
    from tests.emu.emu import EmuPlan
    plan = EmuPlan(8, 32, 0.16)
    x = np.random.RandomState(2).rand(1, 32, 8, 8).astype(np.float32)
    plan.run(x, 1, 32, [0])                                   # the reference operator is accepted
    rp, ci, v = (a.copy() for a in plan.csr)
    # stretch the last row to five contiguous entries: a long row far outside the allowed head
    last = np.arange(27, 32, dtype=np.int32)
    ci2 = np.concatenate([ci[:rp[31]], last]).astype(np.int32)
    v2 = np.concatenate([v[:rp[31]], np.full(5, 0.2, np.float32)]).astype(np.float32)
    rp2 = rp.copy()
    rp2[32] = rp2[31] + 5
    plan.csr = (rp2, ci2, v2)
    with pytest.raises(AssertionError):
        plan.run(x, 1, 32, [0])                               # lct_emu_run returns 100 when build_tables refuses


def test_layer_state_pickles_without_the_native_plan():
    """torch.save(model) / copy.deepcopy(model) (train.py:223, EMA copies): the layer's state carries host constants
    only; the device plan is rebuilt on first use (the CUDA half of this is in test_gpu_parity.py)."""
    import copy
    import pickle
    import hiddenpose_b200 as hp
    layer = hp.lct(spatial=8, crop=32, bin_len=0.16)
    layer._plan = object()                      # stand-in for a live native handle: must not travel
    for clone in (copy.deepcopy(layer), pickle.loads(pickle.dumps(layer))):
        assert clone._plan is None and clone._M == 32 and clone._N == 8
        assert np.array_equal(clone._csr[2], layer._csr[2])
        assert dict(clone.state_dict()) == {}
    layer._plan = None
