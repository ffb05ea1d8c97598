"""Pin the CPU oracle against the reference's own outputs (tests/golden/)."""
import os

import numpy as np
import pytest
import torch

from oracle import lct_oracle as O
from tests._golden import Case, case_names, constants, numpy_lct_fixture

# the two largest fixtures (512x256x256: 46 s constructor, 14 GB; 256x256x256) are compared directly with the CUDA
# path on the GPU box; replaying them through the CPU oracle too is opt-in so that this suite stays at a few minutes
HUGE = [] if os.environ.get("LCT_TEST_HUGE") else ["m512n256_full", "m256n256_full"]
SMALL = [n for n in case_names() if n.startswith(("m32n8", "m64n16"))]
LARGE = [n for n in case_names() if n not in SMALL and n not in HUGE]


@pytest.mark.parametrize("M", [16, 32, 64, 128, 256, 512])
def test_resampling_operator_matches_reference(M):
    c = constants()
    mtx, mtxi = O.resampling_operator(M)
    ref = np.zeros((M, M), np.float32)
    ref[c[f"mtx{M}_rows"], c[f"mtx{M}_cols"]] = c[f"mtx{M}_vals"]
    assert np.array_equal(mtx, ref)            # bit-exact: same float32 tree
    assert np.array_equal(mtxi, ref.T)


@pytest.mark.parametrize("N,M", [(8, 32), (16, 64), (32, 128), (64, 256)])
def test_psf_matches_reference(N, M):
    c = constants()
    psf = O.define_psf(N, M, float(c[f"psf_n{N}m{M}_slope"]))
    ref = np.zeros((2 * M, 2 * N, 2 * N), np.float32)
    zyx = c[f"psf_n{N}m{M}_zyx"]
    ref[zyx[:, 0], zyx[:, 1], zyx[:, 2]] = c[f"psf_n{N}m{M}_vals"]
    assert np.array_equal(psf, ref)


def test_laplacian_matches_reference():
    assert np.array_equal(O.filter_laplacian().astype(np.float32), constants()["laplacian"])


@pytest.mark.parametrize("name", SMALL + LARGE)
def test_forward_and_grad_match_reference(name):
    c = Case(name)
    orc = O.LctOracle(c.N, c.M, c.bin_len, 2.0, c.method, c.material)
    x = torch.from_numpy(c.x).requires_grad_(True)
    y = orc.forward(x, c.tbes, c.tens)
    (gx,) = torch.autograd.grad(y, x, torch.from_numpy(c.g))
    # same ops, same library, same dtype: expect equality up to thread-order noise
    assert c.y_err(y.detach().numpy()) < 1e-6
    assert c.gx_err(gx.numpy()) < 1e-6


def test_numpy_path_of_the_reference_agrees():
    """/root/reference/utils/lct.py:41-59 (the reference's NumPy statement of the transform, run
    unmodified by tests/golden/make_golden_numpy_lct.py) gives the volume the oracle gives, and its display tail
    (:62-82) the same three projections."""
    z, meas = numpy_lct_fixture()
    N, M = int(z["N"]), int(z["M"])
    orc = O.LctOracle(N, M, float(z["bin_len"]), float(z["wall_size"]))
    x = torch.from_numpy(np.ascontiguousarray(np.transpose(meas, [2, 0, 1]))).view(1, 1, M, N, N)     # lct.py:41
    vol = orc.forward(x, [0], [M]).numpy().reshape(M, N, N)
    assert O.rel_l2(vol, z["volume"]) < 2e-6
    views = O.display_views(vol)
    for name in ("front", "left", "top"):
        assert views[name].shape == z[name].shape
        assert O.rel_l2(views[name], z[name]) < 1e-5
    # and the restated tail reproduces the recorded images exactly when fed the recorded volume
    for name, img in O.display_views(z["volume"]).items():
        assert np.array_equal(img, z[name])


@pytest.mark.parametrize("name", ["m64n16_full", "m64n16_window"])
def test_fp64_oracle_agrees(name):
    c = Case(name)
    orc = O.LctOracle(c.N, c.M, c.bin_len, 2.0, c.method, c.material, dtype=torch.float64)
    y = orc.forward(torch.from_numpy(c.x), c.tbes, c.tens)
    assert c.y_err(y.numpy()) < 2e-6


def test_adjoint_identity_fp64():
    """<Ax, g> == <x, A^T g>."""
    orc = O.LctOracle(8, 32, 0.16, dtype=torch.float64)
    rs = np.random.RandomState(3)
    x = torch.from_numpy(rs.rand(2, 1, 20, 8, 8))
    g = torch.from_numpy(rs.randn(2, 1, 32, 8, 8))
    y = orc.forward(x, [5, 5], [25, 25])
    gx = orc.vjp(x.shape, g, [5, 5], [25, 25])
    lhs, rhs = float((y * g).sum()), float((x * gx).sum())
    assert abs(lhs - rhs) <= 1e-12 * max(abs(lhs), abs(rhs))


def test_window_asserts():
    orc = O.LctOracle(8, 32, 0.16)
    x = torch.zeros(1, 1, 32, 8, 8)
    with pytest.raises(AssertionError):
        orc.forward(x, [-1], [31])
    with pytest.raises(AssertionError):
        orc.forward(x, [1], [33])
