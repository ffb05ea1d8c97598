"""FeatureExtraction's skip branch (the op that feeds the LCT):
oracle and drop-in module against outputs of the reference module (tests/golden/skip_*.npz)
on the CPU, the CUDA stencil through the C ABI against both on the GPU."""
import ctypes
import glob
import os

import numpy as np
import pytest
import torch

from oracle import skip_oracle as S
from oracle.lct_oracle import rel_l2

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "skip_*.npz")))
TOL = 1e-5          # relative L2, fp32 (summation order differs from the reference's conv)


class SkipCase:
    def __init__(self, path):
        z = np.load(path)
        self.B, self.D, self.T, self.N = (int(z[k]) for k in "BDTN")
        rs = np.random.RandomState(int(z["seed"]))
        self.x = rs.rand(self.B, 1, self.T, self.N, self.N).astype(np.float32)
        self.g = rs.randn(self.B, self.D, self.T, self.N, self.N).astype(np.float32)
        self.feat, self.out, self.gw = z["feat"], z["out"], z["gw"]
        self.gx_skip = z["gx_total"].astype(np.float64) - z["gx_conv1"].astype(np.float64)
        self.gx_total = z["gx_total"]
        self.state = {k[len("param:"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("param:")}
        self.w = z["param:weights"]


def test_golden_files_present():
    assert len(GOLDEN) == 3


@pytest.mark.parametrize("path", GOLDEN, ids=os.path.basename)
def test_oracle_matches_reference_module(path):
    c = SkipCase(path)
    assert rel_l2(S.skip_sum(c.feat, c.x, c.w), c.out) <= 1e-6
    gx, gw = S.skip_sum_vjp(c.g, c.x, c.w)
    assert rel_l2(gx, c.gx_skip) <= 1e-5          # the golden value is a float32 difference of two gradients
    assert rel_l2(gw, c.gw) <= 1e-6


@pytest.mark.parametrize("path", GOLDEN, ids=os.path.basename)
def test_dropin_module_loads_reference_parameters(path):
    """Same parameter names / shapes as the reference (strict load) and the same function on CPU."""
    from hiddenpose_b200.feature_extraction import FeatureExtraction
    c = SkipCase(path)
    m = FeatureExtraction(basedim=c.D, in_channels=1, stride=1)
    m.load_state_dict(c.state, strict=True)
    x = torch.from_numpy(c.x).requires_grad_(True)
    out = m(x)
    out.backward(torch.from_numpy(c.g))
    assert rel_l2(out.detach().numpy(), c.out) <= 1e-6
    assert rel_l2(x.grad.numpy(), c.gx_total) <= 1e-5
    assert rel_l2(m.weights.grad.numpy(), c.gw) <= 1e-5


def test_dropin_module_initial_kernel_and_errors():
    from hiddenpose_b200.feature_extraction import FeatureExtraction
    m = FeatureExtraction(basedim=1, in_channels=1, stride=1)
    w = m.weights.detach().numpy()[0, 0]
    assert w.sum() == pytest.approx(1.0) and np.count_nonzero(w) == 8 and w[0].sum() == 0      # feature_extraction.py:141-145
    assert m.weights.requires_grad
    with pytest.raises(AssertionError):
        FeatureExtraction(basedim=1, in_channels=2)
    # stride 2 (the constructor default) keeps the torch expression
    m2 = FeatureExtraction(basedim=1, in_channels=1)
    assert m2(torch.rand(1, 1, 8, 8, 8)).shape == (1, 1, 4, 4, 4)


# ---------------------------------------------------------------------------------------------
# GPU: the CUDA stencil through the C ABI
# ---------------------------------------------------------------------------------------------

def _abi_skip_sum(feat, x, w, out=None):
    from hiddenpose_b200 import _native
    lib = _native.load()
    B, D, T, N = feat.shape[0], feat.shape[1], feat.shape[2], feat.shape[3]
    out = torch.empty_like(feat) if out is None else out
    st = torch.cuda.current_stream().cuda_stream
    _native.check(lib.lct_skip_sum(feat.data_ptr(), x.data_ptr(), w.data_ptr(), B, D, T, N, out.data_ptr(), st))
    return out


def _abi_skip_backward(g, x, w, want_x=True, want_w=True):
    from hiddenpose_b200 import _native
    lib = _native.load()
    B, D, T, N = g.shape[0], g.shape[1], g.shape[2], g.shape[3]
    gx = torch.empty_like(x) if want_x else None
    gw = torch.empty(27, device=g.device) if want_w else None
    nbytes = lib.lct_skip_workspace_bytes(B, T, N)
    ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=g.device)
    st = torch.cuda.current_stream().cuda_stream
    _native.check(lib.lct_skip_sum_backward(g.data_ptr(), x.data_ptr(), w.data_ptr(), B, D, T, N,
                                            gx.data_ptr() if want_x else None, gw.data_ptr() if want_w else None,
                                            ws.data_ptr(), nbytes, st))
    return gx, gw


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN, ids=os.path.basename)
def test_cuda_skip_sum_vs_reference_module(path):
    c = SkipCase(path)
    feat, x, g = (torch.from_numpy(a).cuda() for a in (c.feat, c.x, c.g))
    w = torch.from_numpy(c.w).cuda().reshape(27)
    out = _abi_skip_sum(feat, x, w)
    gx, gw = _abi_skip_backward(g, x, w)
    assert rel_l2(out.cpu().numpy(), c.out) <= TOL
    assert rel_l2(gx.cpu().numpy(), c.gx_skip) <= TOL
    assert rel_l2(gw.cpu().numpy().reshape(c.gw.shape), c.gw) <= TOL


@pytest.mark.gpu
@pytest.mark.parametrize("B,D,T,N", [(1, 1, 1, 8), (2, 3, 5, 8), (1, 1, 24, 16), (3, 1, 49, 32), (1, 2, 100, 64), (1, 1, 23, 128), (1, 1, 7, 256)])
def test_cuda_skip_sum_vs_oracle(B, D, T, N):
    rs = np.random.RandomState(B * 1000 + T * 10 + N)
    feat = rs.randn(B, D, T, N, N).astype(np.float32)
    x = rs.rand(B, 1, T, N, N).astype(np.float32)
    g = rs.randn(B, D, T, N, N).astype(np.float32)
    w = rs.randn(27).astype(np.float32)
    fd, xd, gd, wd = (torch.from_numpy(a).cuda() for a in (feat, x, g, w))
    out = _abi_skip_sum(fd, xd, wd)
    gx, gw = _abi_skip_backward(gd, xd, wd)
    gx_o, gw_o = S.skip_sum_vjp(g, x, w)
    assert rel_l2(out.cpu().numpy(), S.skip_sum(feat, x, w)) <= TOL
    assert rel_l2(gx.cpu().numpy(), gx_o) <= TOL
    assert rel_l2(gw.cpu().numpy(), gw_o.ravel()) <= TOL
    # in place on feat, and each gradient on its own
    assert torch.equal(_abi_skip_sum(fd.clone(), xd, wd), out)
    inplace = fd.clone()
    _abi_skip_sum(inplace, xd, wd, out=inplace)
    assert torch.equal(inplace, out)
    assert torch.equal(_abi_skip_backward(gd, xd, wd, want_w=False)[0], gx)
    assert torch.equal(_abi_skip_backward(gd, xd, wd, want_x=False)[1], gw)      # fixed reduction order


@pytest.mark.gpu
def test_cuda_skip_sum_full_size_properties():
    """At the BASELINE shape (8 x 256 x 64 x 64): adjoint identity <S x, g> = <x, S^T g>, and the weight
    gradient as the derivative along each tap, gw[k] = <g, S_{e_k} x> (S is linear in w)."""
    torch.manual_seed(410)
    B, T, N = 8, 256, 64
    x = torch.rand(B, 1, T, N, N, device="cuda")
    g = torch.randn(B, 1, T, N, N, device="cuda")
    w = torch.randn(27, device="cuda")
    zero = torch.zeros_like(g)
    sx = _abi_skip_sum(zero, x, w)
    gx, gw = _abi_skip_backward(g, x, w)
    lhs = (sx.double() * g.double()).sum().item()
    rhs = (x.double() * gx.double()).sum().item()
    assert abs(lhs - rhs) <= 1e-6 * max(abs(lhs), abs(rhs))
    for k in (0, 13, 26, 5):
        e = torch.zeros(27, device="cuda")
        e[k] = 1.0
        d = (_abi_skip_sum(zero, x, e).double() * g.double()).sum().item()
        assert abs(d - gw[k].item()) <= 1e-4 * max(1.0, abs(d))
    # identity kernel: out = feat + x
    e = torch.zeros(27, device="cuda")
    e[13] = 1.0
    assert torch.equal(_abi_skip_sum(g, x, e), g + x)


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN, ids=os.path.basename)
def test_cuda_dropin_module_vs_reference_module(path):
    """The whole drop-in module on the GPU (learned branch on torch's conv in full fp32, skip branch and sum
    on the CUDA stencil) against the reference module's output and gradients."""
    from hiddenpose_b200.feature_extraction import FeatureExtraction
    c = SkipCase(path)
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        m = FeatureExtraction(basedim=c.D, in_channels=1, stride=1)
        m.load_state_dict(c.state, strict=True)
        m = m.cuda()
        x = torch.from_numpy(c.x).cuda().requires_grad_(True)
        out = m(x)
        out.backward(torch.from_numpy(c.g).cuda())
    finally:
        torch.backends.cudnn.allow_tf32 = old
    assert rel_l2(out.detach().cpu().numpy(), c.out) <= TOL
    assert rel_l2(x.grad.cpu().numpy(), c.gx_total) <= 1e-4
    assert rel_l2(m.weights.grad.cpu().numpy(), c.gw) <= TOL


@pytest.mark.gpu
def test_cuda_skip_sum_errors():
    from hiddenpose_b200 import _native
    lib = _native.load()
    x = torch.rand(1, 1, 4, 6, 6, device="cuda")
    w = torch.rand(27, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    assert lib.lct_skip_sum(x.data_ptr(), x.data_ptr(), w.data_ptr(), 1, 1, 4, 6, x.data_ptr(), st) != 0      # N % 4
    y = torch.rand(1, 1, 4, 8, 8, device="cuda")
    assert lib.lct_skip_sum(y.data_ptr(), y.data_ptr(), w.data_ptr(), 1, 1, 4, 8, y.data_ptr(), st) != 0      # x aliases out
    assert lib.lct_skip_sum_backward(y.data_ptr(), y.data_ptr(), w.data_ptr(), 1, 1, 4, 8, None, w.data_ptr(), None, 0, st) != 0
    assert lib.lct_skip_workspace_bytes(1, 4, 8) == 27 * 4
