#!/usr/bin/env python
"""Mint golden vectors by running the REFERENCE ITSELF (build container only).

    python tests/golden/make_golden.py            # needs /root/reference

Imports ``/root/reference/models/tflct.py`` unmodified and calls its own
``lct.forward`` (tflct.py:94-179) and autograd backward on seeded inputs.  Two
shims, both outside the reference's arithmetic:

* ``torch.rfft`` / ``torch.ifft`` (removed in PyTorch 1.8) are provided as thin
  wrappers over ``torch.fft.fftn`` / ``ifftn`` with the 1.7 conventions;
* a 6-line subclass sets ``self.crop = crop`` before calling the reference's own
  ``parpareparam()`` (``tflct.lct`` pins crop to 128, tflct.py:19).  The
  ``crop == 128`` cases below use the reference class directly, no subclass.

The reference cannot travel to the GPU box, so the outputs are committed as
small ``.npz`` fixtures next to this script; tests regenerate the inputs from
the seeds recorded in each file (NumPy ``RandomState``, stable across versions).
Large cases store a fixed sample of the output plus its L2 norm instead of the
whole volume.
"""
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def load_reference():
    if not os.path.isdir(REF):
        raise SystemExit("reference tree not present; golden vectors can only be minted in the build container")
    sys.path.insert(0, REF)
    if not hasattr(torch, "rfft"):
        def rfft(x, signal_ndim, normalized=False, onesided=True):
            assert not normalized and not onesided
            return torch.view_as_real(torch.fft.fftn(x, dim=tuple(range(-signal_ndim, 0))))

        def ifft(x, signal_ndim, normalized=False):
            assert not normalized
            return torch.view_as_real(torch.fft.ifftn(torch.view_as_complex(x.contiguous()),
                                                      dim=tuple(range(-signal_ndim, 0))))
        torch.rfft, torch.ifft = rfft, ifft
    from models.tflct import lct as ref_lct          # noqa: E402
    import utils.helper as ref_helper                # noqa: E402

    class lct_cropfix(ref_lct):
        def __init__(self, spatial=256, crop=128, **kw):
            self._crop = crop
            super().__init__(spatial=spatial, crop=crop, **kw)

        def parpareparam(self):
            self.crop = self._crop
            super().parpareparam()

    return ref_lct, lct_cropfix, ref_helper


def make_input(seed, shape):
    return np.random.RandomState(seed).rand(*shape).astype(np.float32)


def make_grad(seed, shape):
    return np.random.RandomState(seed + 1000).randn(*shape).astype(np.float32)


def sample_idx(n, k=4096, seed=7):
    k = min(k, n)
    return np.sort(np.random.RandomState(seed).choice(n, size=k, replace=False))


# (name, spatial N, crop M, B, D, tbe, ten, method, material, store_full)
CASES = [
    ("m64n16_full",      16,  64, 2, 1,  0,  64, "lct", "diffuse",  True),
    ("m64n16_window",    16,  64, 1, 2, 10,  50, "lct", "diffuse",  True),
    ("m64n16_specular",  16,  64, 1, 1,  0,  64, "lct", "specular", True),
    ("m64n16_bp",        16,  64, 1, 1,  0,  64, "bp",  "diffuse",  True),
    ("m32n8_full",        8,  32, 3, 1,  0,  32, "lct", "diffuse",  True),
    ("m128n32_window",   32, 128, 1, 2, 20, 120, "lct", "diffuse",  False),   # reference class verbatim (crop 128)
    ("m128n128_full",   128, 128, 1, 1,  0, 128, "lct", "diffuse",  False),   # reference training shape, verbatim class
    ("m256n64_full",     64, 256, 1, 1,  0, 256, "lct", "diffuse",  False),   # BASELINE config 1/2 shape
    ("m32n256_window",  256,  32, 1, 1,  2,  29, "lct", "diffuse",  False),   # 512-point spatial lines: the parity-split K2/K3/K4
    ("m512n16_window",   16, 512, 1, 2,  5, 500, "lct", "diffuse",  False),   # 512 time bins: the 32-wide time kernels
    # round 2: every remaining BASELINE shape and every compiled (M, N) the CPU oracle is too slow for in a test
    ("m512n128_full",   128, 512, 1, 1,  0, 512, "lct", "diffuse",  False),   # BASELINE config 3 unit (512x128x128)
    ("m512n256_full",   256, 512, 1, 1,  0, 512, "lct", "diffuse",  False),   # BASELINE config 5 (512x256x256; 14 GB on the CPU)
    ("m256n128_window", 128, 256, 1, 1,  3, 250, "lct", "diffuse",  False),   # mid shape, partial window
    ("m64n256_full",    256,  64, 1, 1,  0,  64, "lct", "diffuse",  False),
    ("m128n256_window", 256, 128, 1, 1,  1, 127, "lct", "diffuse",  False),   # reference class verbatim (crop 128)
    ("m256n256_full",   256, 256, 1, 1,  0, 256, "lct", "diffuse",  False),
]


def bin_len_for(M):
    # released configs keep trange = M * bin_len = 5.12 (config_noise.py:20,26; train.py:80-81)
    return 0.01 * 512 / M


def main():
    only = set(sys.argv[1:])          # case names to (re)generate; none = everything, constants included
    ref_lct, lct_cropfix, helper = load_reference()
    torch.manual_seed(0)
    torch.set_num_threads(8)
    if not only:
        write_constants(helper)
    write_cases(ref_lct, lct_cropfix, only)


def write_constants(helper):

    # ---- constants -------------------------------------------------------
    const = {}
    for M in (16, 32, 64, 128, 256, 512):
        mtx, mtxi = helper.resamplingOperator(M)
        assert np.array_equal(mtx.T, mtxi)
        r, c = np.nonzero(mtx)
        const[f"mtx{M}_rows"] = r.astype(np.int32)
        const[f"mtx{M}_cols"] = c.astype(np.int32)
        const[f"mtx{M}_vals"] = mtx[r, c].astype(np.float32)
    for (N, M) in ((8, 32), (16, 64), (32, 128), (64, 256)):
        slope = 1.0 / (M * bin_len_for(M))
        psf = helper.definePsf(N, M, slope)
        z, y, x = np.nonzero(psf)
        const[f"psf_n{N}m{M}_zyx"] = np.stack([z, y, x], 1).astype(np.int32)
        const[f"psf_n{N}m{M}_vals"] = psf[z, y, x].astype(np.float32)
        const[f"psf_n{N}m{M}_slope"] = np.float64(slope)
    const["laplacian"] = helper.filterLaplacian().astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "constants.npz"), **const)
    print("constants.npz written")



def write_cases(ref_lct, lct_cropfix, only):
    # ---- forward / backward ---------------------------------------------
    for seed, (name, N, M, B, D, tbe, ten, method, material, full) in enumerate(CASES):
        if only and name not in only:
            continue
        cls = ref_lct if M == 128 else lct_cropfix
        layer = cls(spatial=N, crop=M, bin_len=bin_len_for(M), wall_size=2.0, method=method, material=material)
        assert layer.crop == M
        layer.todev("cpu", D)
        tin = ten - tbe
        x_np = make_input(seed, (B, D, tin, N, N))
        g_np = make_grad(seed, (B, D, M, N, N))
        x = torch.from_numpy(x_np).requires_grad_(True)
        y = layer(x, [tbe] * B, [ten] * B)
        (gx,) = torch.autograd.grad(y, x, torch.from_numpy(g_np))
        y_np = y.detach().numpy().astype(np.float32)
        gx_np = gx.numpy().astype(np.float32)
        rec = dict(seed=np.int64(seed), N=np.int64(N), M=np.int64(M), B=np.int64(B), D=np.int64(D),
                   tbe=np.int64(tbe), ten=np.int64(ten), method=np.str_(method), material=np.str_(material),
                   bin_len=np.float64(bin_len_for(M)), full=np.bool_(full),
                   y_norm=np.float64(np.linalg.norm(y_np.astype(np.float64))),
                   gx_norm=np.float64(np.linalg.norm(gx_np.astype(np.float64))),
                   invpsf_real_sample=layer.invpsf_real.numpy().ravel()[sample_idx(8 * M * N * N)],
                   invpsf_imag_sample=layer.invpsf_imag.numpy().ravel()[sample_idx(8 * M * N * N)])
        if full:
            rec["y"], rec["gx"] = y_np, gx_np
        else:
            iy, ig = sample_idx(y_np.size), sample_idx(gx_np.size)
            rec["y_idx"], rec["y_sample"] = iy, y_np.ravel()[iy]
            rec["gx_idx"], rec["gx_sample"] = ig, gx_np.ravel()[ig]
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **rec)
        print(f"{name}: |y|={rec['y_norm']:.6e} |gx|={rec['gx_norm']:.6e}")


if __name__ == "__main__":
    main()
