"""CPU oracle for the HiddenPose light-cone-transform (LCT) layer.

TEST INFRASTRUCTURE ONLY.  Nothing under ``hiddenpose_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and the CPU-baseline /
``--impl reference`` legs of ``bench.py`` use it, and there only as the checker
or as the timed CPU baseline -- never as the thing shipped.

What it is: a restatement, op for op, of the reference layer

    /root/reference/models/tflct.py:32-79    (constants: ``parpareparam``)
    /root/reference/models/tflct.py:94-179   (``forward``)
    /root/reference/utils/helper.py:13-32    (``filterLaplacian``)
    /root/reference/utils/helper.py:35-69    (``resamplingOperator``)
    /root/reference/utils/helper.py:72-125   (``definePsf``)

(``models/feature_propagation.py:46-257`` is the same arithmetic with renamed
constructor arguments) in NumPy / CPU PyTorch.  Two departures, both forced:

* ``torch.rfft(x, 3, onesided=False)`` / ``torch.ifft(x, 3)`` (tflct.py:144,151)
  were removed in PyTorch 1.8; they are restated with ``torch.fft.fftn`` /
  ``torch.fft.ifftn`` which have the same un-normalised-forward / 1/n-inverse
  convention.
* ``tflct.lct.__init__`` pins ``self.crop = 128`` whatever ``crop`` is passed
  (tflct.py:19); the oracle honours ``crop`` like ``feature_propagation.LCT``
  does (feature_propagation.py:60).  At ``crop == 128`` there is no difference.

Parity pinning: the reference has no tests or golden vectors of its own
(SURVEY.md section 4).  The oracle is pinned instead against outputs of the
reference itself, run in the build container by
``tests/golden/make_golden.py`` (which imports ``/root/reference`` with a
two-function ``torch.rfft``/``torch.ifft`` shim) and committed under
``tests/golden/``; ``tests/test_oracle.py`` replays them.
"""
from __future__ import annotations

import numpy as np
import torch

SNR = 1e-1          # tflct.py:42
LIGHT_SPEED = 3e8   # tflct.py:34


# --------------------------------------------------------------------------
# constants
# --------------------------------------------------------------------------
def resampling_operator(M: int):
    """helper.py:35-69.  Returns dense float32 ``(mtx, mtxi)``, each M x M.

    The reference builds an M^2 x M sparse matrix with one entry per row,
    ``S[x-1, ceil(sqrt(x))-1] = 1/sqrt(x)`` for x = 1..M^2 (helper.py:43-55),
    then halves the row count log2(M) times with
    ``S <- 0.5*(S[0::2] + S[1::2])`` (helper.py:56-58), all in float32.
    Row i of the result therefore only depends on rows iM..iM+M-1 of S; the
    same pairwise float32 tree is evaluated here one output row at a time on a
    dense M x M block (adding the explicit zeros is exact).
    """
    assert 2 ** int(np.log2(M)) == M                       # helper.py:40
    K = int(np.log2(M))
    mtx = np.zeros((M, M), dtype=np.float32)
    rows = np.arange(M)
    for i in range(M):
        x = (np.arange(i * M, (i + 1) * M, dtype=np.float32) + np.float32(1))   # helper.py:43-44
        col = (np.ceil(np.sqrt(x)) - 1).astype(np.int64)                        # helper.py:48
        blk = np.zeros((M, M), dtype=np.float32)
        blk[rows, col] = (np.float32(1.0) / np.sqrt(x)).astype(np.float32)      # helper.py:53,55
        for _ in range(K):                                                      # helper.py:57-58
            blk = np.float32(0.5) * (blk[0::2, :] + blk[1::2, :])
        mtx[i] = blk[0]
    return mtx, np.ascontiguousarray(mtx.T)                                     # helper.py:61


def define_psf(N: int, M: int, slope: float):
    """helper.py:72-125.  float32 ``(2M, 2N, 2N)`` light-cone PSF."""
    x = np.arange(2 * N, dtype=np.float32)
    x = x / (2 * N - 1) * 2 - 1                            # helper.py:79-80
    y = x                                                  # helper.py:84
    z = np.arange(2 * M, dtype=np.float32)
    z = z / (2 * M - 1) * 2                                # helper.py:87-88
    gy, gx, gz = np.meshgrid(x, y, z)                      # helper.py:93
    a = (4 * slope) ** 2 * (gx ** 2 + gy ** 2) - gz        # helper.py:96
    b = np.abs(a)                                          # helper.py:97
    c = np.min(b, axis=2, keepdims=True)                   # helper.py:100
    d = (np.abs(b - c) < 1e-8).astype(np.float32)          # helper.py:103-104
    e = d / np.sqrt(np.sum(d))                             # helper.py:112
    f = np.roll(np.roll(e, shift=N, axis=0), shift=N, axis=1)   # helper.py:115-116
    return np.transpose(f, [2, 0, 1])                      # helper.py:123


def filter_laplacian():
    """helper.py:13-32.  5x5x5 zero-mean Laplacian-of-Gaussian, sigma 1."""
    lim = 2
    dims = np.arange(-lim, lim + 1, dtype=np.float32)
    y, x, z = np.meshgrid(dims, dims, dims)
    r2 = x ** 2 + y ** 2 + z ** 2
    w = np.exp(-r2 / 2.0)
    w = w / np.sum(w)
    w1 = w * (r2 - 3.0)
    return w1 - np.mean(w1)


class LctOracle:
    """CPU restatement of ``tflct.lct`` (``crop`` honoured).

    ``dtype=torch.float32`` mirrors the reference's arithmetic;
    ``dtype=torch.float64`` keeps the reference's float32 constants (cast up)
    but carries the data path in double -- the tie-breaker when two float32
    results disagree.
    """

    def __init__(self, spatial=256, crop=128, bin_len=0.01, wall_size=2.0,
                 method="lct", material="diffuse", dtype=torch.float32):
        assert 2 ** int(np.log2(crop)) == crop             # tflct.py:20
        self.spatial_grid, self.crop = int(spatial), int(crop)
        self.bin_len, self.wall_size = bin_len, wall_size
        self.method, self.material, self.dtype = method, material, dtype
        M, N = self.crop, self.spatial_grid

        # tflct.py:34-42
        self.width = wall_size / 2.0
        self.bin_resolution = bin_len / LIGHT_SPEED
        self.trange = M * LIGHT_SPEED * self.bin_resolution
        self.snr = SNR

        # tflct.py:49-52
        gridz = np.arange(M, dtype=np.float32) / (M - 1)
        self.gridz = torch.from_numpy(gridz.reshape(1, M, 1, 1).astype(np.float32))

        # tflct.py:55-65
        slope = self.width / self.trange
        psf = define_psf(N, M, slope)
        fpsf = np.fft.fftn(psf)
        if method == "lct":
            invpsf = np.conjugate(fpsf) / (1 / self.snr + np.real(fpsf) ** 2 + np.imag(fpsf) ** 2)
        elif method == "bp":
            invpsf = np.conjugate(fpsf)
        else:
            raise ValueError(method)
        self.invpsf_real = torch.from_numpy(np.real(invpsf).astype(np.float32)).unsqueeze(0)
        self.invpsf_imag = torch.from_numpy(np.imag(invpsf).astype(np.float32)).unsqueeze(0)

        # tflct.py:68-70
        mtx, mtxi = resampling_operator(M)
        self.mtx = torch.from_numpy(mtx)
        self.mtxi = torch.from_numpy(mtxi)

        # tflct.py:73-77
        if method == "bp":
            self.lapw = torch.from_numpy(filter_laplacian().astype(np.float32)).reshape(1, 1, 5, 5, 5)

    # ----------------------------------------------------------------------
    def forward(self, feat, tbes, tens):
        """tflct.py:94-179, op for op, in ``self.dtype`` on the CPU."""
        dt, dev = self.dtype, feat.device      # dev is the CPU everywhere except the large-size GPU cross-checks in tests/
        feat = feat.to(dt)
        bnum, dnum, tnum, hnum, wnum = feat.shape
        M, N = self.crop, self.spatial_grid
        for tbe, ten in zip(tbes, tens):                    # tflct.py:99-101
            assert tbe >= 0
            assert ten <= M
        padded = []
        for i in range(bnum):                               # tflct.py:104-110
            head = torch.zeros((1, dnum, tbes[i], hnum, wnum), dtype=dt, device=dev)
            tail = torch.zeros((1, dnum, M - tens[i], hnum, wnum), dtype=dt, device=dev)
            padded.append(torch.cat([head, feat[i:i + 1], tail], dim=2))
        data = torch.cat(padded, dim=0)
        assert hnum == wnum and hnum == N                   # tflct.py:113-114
        data = data.view(bnum * dnum, M, hnum, wnum)        # tflct.py:121

        gridz = self.gridz.to(dev, dt)
        if self.material == "diffuse":                      # tflct.py:124-127
            data = data * (gridz ** 4)
        elif self.material == "specular":
            data = data * (gridz ** 2)

        pad = torch.zeros((bnum * dnum, 2 * M, 2 * N, 2 * N), dtype=dt, device=dev)        # tflct.py:131-133
        tmp = torch.matmul(self.mtx.to(dev, dt), data.view(bnum * dnum, M, -1))     # tflct.py:135-138
        pad[:, :M, :N, :N] = tmp.view(bnum * dnum, M, N, N)                    # tflct.py:140

        fre = torch.fft.fftn(pad, dim=(-3, -2, -1))                            # tflct.py:144
        fr, fi = fre.real, fre.imag
        wr, wi = self.invpsf_real.to(dev, dt), self.invpsf_imag.to(dev, dt)
        re_real = fr * wr - fi * wi                                            # tflct.py:148
        re_imag = fr * wi + fi * wr                                            # tflct.py:149
        re = torch.fft.ifftn(torch.complex(re_real, re_imag), dim=(-3, -2, -1))  # tflct.py:150-151

        vol = re.real[:, :M, :N, :N]                                           # tflct.py:153
        out = torch.matmul(self.mtxi.to(dev, dt), vol.reshape(bnum * dnum, M, -1))  # tflct.py:156-159
        out = out.view(bnum * dnum, M, N, N)

        if self.method == "bp":                                                # tflct.py:164-175
            v = torch.nn.functional.pad(out.unsqueeze(1), (2,) * 6, mode="replicate")
            v = torch.nn.functional.conv3d(v, self.lapw.to(dev, dt))
            out = v.squeeze(1)
            out[:, :1] = 0
        return out.view(bnum, dnum, M, hnum, wnum)                             # tflct.py:177-179

    __call__ = forward

    def vjp(self, feat_shape, grad_out, tbes, tens):
        """Gradient of ``sum(forward(x) * grad_out)`` w.r.t. x, by autograd
        through :meth:`forward` -- what the reference's backward computes
        (the reference has no explicit backward)."""
        x = torch.zeros(feat_shape, dtype=self.dtype, device=grad_out.device, requires_grad=True)
        y = self.forward(x, tbes, tens)
        (g,) = torch.autograd.grad(y, x, grad_out.to(self.dtype))
        return g


def display_views(volume_mxnxn):
    """The display tail of the reference's NumPy path, /root/reference/utils/lct.py:62-82, on an
    un-clamped (M, N, N) volume: clamp below zero (:62), keep the first ``M * 100 // 128`` depth
    bins (:64-65), divide by the maximum (:66), then the three maximum projections, each divided
    by its own maximum as passed to ``cv2.imshow`` (:68-69 front, :78-79 left, :82-83 top)."""
    v = np.array(volume_mxnxn, dtype=np.float32, copy=True)
    v[v < 0] = 0
    v = v[: v.shape[0] * 100 // 128]
    v = v / np.max(v)
    out = {}
    for name, axis in (("front", 0), ("left", 1), ("top", 2)):
        view = np.max(v, axis=axis)
        out[name] = view / np.max(view)
    return out


def rel_l2(a, b):
    """Relative L2 error of ``a`` against ``b`` (the parity metric; outputs
    are O(1e-6) in magnitude so absolute tolerances are meaningless)."""
    a = torch.as_tensor(a).double().flatten()
    b = torch.as_tensor(b).double().flatten()
    den = torch.linalg.norm(b)
    num = torch.linalg.norm(a - b)
    if den == 0:
        return float(num)
    return float(num / den)
