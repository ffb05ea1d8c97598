"""Shared implementation behind the two reference-facing classes
(``tflct.lct`` and ``feature_propagation.LCT``)."""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.nn as nn

from . import operators as ops
from .lct_function import LctFunction, LctPlan, remember_minmax


def _as_device(dev):
    """The reference passes 'cpu', 'cuda', 'cuda:k', a torch.device, or a bare int
    (cfg.DEVICE = 0, config/config_noise.py:7) straight into ``Tensor.to``."""
    if isinstance(dev, torch.device):
        d = dev
    elif isinstance(dev, int):
        d = torch.device("cuda", dev)
    else:
        d = torch.device(dev)
    if d.type == "cuda" and d.index is None:
        d = torch.device("cuda", torch.cuda.current_device() if torch.cuda.is_available() else 0)
    return d


class LctLayerBase(nn.Module):
    """Constants on the host, a device plan per ``todev`` call, forward/backward in CUDA.

    Holds no parameters and no buffers, so ``state_dict()`` stays empty exactly like
    the reference layer's (released checkpoints carry no LCT keys; test.py:133-136
    loads them strictly).
    """

    def _setup(self, spatial, temporal, bin_len, wall_size, method, material):
        if 2 ** int(np.log2(temporal)) != temporal:
            raise AssertionError("time size should be a power of 2")          # tflct.py:20,39
        self._N, self._M = int(spatial), int(temporal)
        self.bin_len, self.wall_size = bin_len, wall_size
        self._method, self.material = method, material
        # tflct.py:34-42
        self.c = ops.LIGHT_SPEED
        self.width = self.wall_size / 2.0
        self.bin_resolution = self.bin_len / self.c
        self.trange = self._M * self.c * self.bin_resolution
        self.snr = ops.SNR
        if method not in ("lct", "bp"):
            raise ValueError(f"method {method!r} is not supported (lct | bp)")
        # host operators in the compact forms the kernels consume (tflct.py:49-70)
        self._csr = ops.resampling_csr(self._M)
        self._falloff = ops.falloff(self._M, material)
        # the inverse filter (tflct.py:55-65): only the PSF support is found on the host; the library
        # builds the spectrum on the GPU in todev().  HIDDENPOSE_LCT_HOST_FILTER=1 builds it on the host instead.
        self._psf = ops.psf_support(self._N, self._M, self.width / self.trange)
        self._filter_half_cache = None
        self._host_filter = os.environ.get("HIDDENPOSE_LCT_HOST_FILTER", "0") not in ("", "0")
        self._lapw = ops.laplacian_filter() if method == "bp" else None       # tflct.py:73-77
        self._plan, self._dev, self.dnum = None, torch.device("cpu"), 2
        # reduce the volume's per-channel min/max in the last kernel and remember them for
        # normalize_feature (FeaturePropagation turns this on: NlosPose.py:53-54 always normalises next)
        self.fuse_minmax = False

    # -- pickling / copying ---------------------------------------------------------------
    # The reference ends training with ``torch.save(model, path)`` (train.py:223) and EMA helpers deep-copy
    # the model.  The device plan is a native handle (not picklable, not copyable): it is dropped from the
    # state and rebuilt from the host constants on the next ``forward`` / ``todev`` of the copy.
    def __getstate__(self):
        state = self.__dict__.copy()
        state["_plan"] = None
        return state

    def __setstate__(self, state):
        super().__setstate__(state)
        self._plan = None

    def _ensure_plan(self):
        if self._plan is None and self._dev.type == "cuda":
            self.todev(self._dev, self.dnum)
        return self._plan

    @property
    def _filter_half(self):
        """Host-built half spectrum (M+1, 2N, 2N) complex64; only materialised when asked for."""
        if self._filter_half_cache is None:
            self._filter_half_cache = ops.inverse_filter_half(self._N, self._M, self.width / self.trange,
                                                              self._method, self.snr)
        return self._filter_half_cache

    # -- reference attributes, materialised on demand --------------------------------
    @property
    def gridz_1xMx1x1(self):
        g = np.arange(self._M, dtype=np.float32) / (self._M - 1)
        return torch.from_numpy(g.reshape(1, -1, 1, 1))

    @property
    def mtx_MxM(self):
        return torch.from_numpy(ops.csr_to_dense(*self._csr, self._M))

    @property
    def mtxi_MxM(self):
        return torch.from_numpy(np.ascontiguousarray(ops.csr_to_dense(*self._csr, self._M).T))

    def _full_filter(self):
        """(2M, 2N, 2N) complex64 spectrum rebuilt from the stored half by Hermitian symmetry."""
        M, N, h = self._M, self._N, self._filter_half
        full = np.empty((2 * M, 2 * N, 2 * N), dtype=np.complex64)
        full[:M + 1] = h
        idx = (-np.arange(2 * N)) % (2 * N)
        full[M + 1:] = np.conj(h[M - 1:0:-1][:, idx][:, :, idx])
        return full

    @property
    def invpsf_real(self):
        return torch.from_numpy(np.ascontiguousarray(self._full_filter().real)).unsqueeze(0)

    @property
    def invpsf_imag(self):
        return torch.from_numpy(np.ascontiguousarray(self._full_filter().imag)).unsqueeze(0)

    @property
    def gridz_1xMx1x1_todev(self):
        return self.gridz_1xMx1x1.to(self._dev)

    @property
    def mtx_MxM_todev(self):
        return self.mtx_MxM.to(self._dev)

    @property
    def mtxi_MxM_todev(self):
        return self.mtxi_MxM.to(self._dev)

    @property
    def invpsf_real_todev(self):
        return self.invpsf_real.to(self._dev)

    @property
    def invpsf_imag_todev(self):
        return self.invpsf_imag.to(self._dev)

    # -- tflct.py:81-92 -----------------------------------------------------------------
    def todev(self, dev, dnum):
        """Move the layer to ``dev`` and size it for ``dnum`` feature channels.

        On a CUDA device this builds the native plan (device copies of the CSR
        operator, the half-spectrum filter and the twiddle tables).  The
        reference's ``datapad`` scratch (tflct.py:83) has no counterpart: the
        zero padding is implicit in the kernels.
        """
        d = _as_device(dev)
        self._dev, self.dnum = d, int(dnum)
        if d.type == "cuda":
            if self._plan is None or self._plan.device != d:
                self._plan = LctPlan(self._M, self._N, self._csr, self._falloff,
                                     self._filter_half if self._host_filter else None, d, lapw=self._lapw,
                                     psf=self._psf, snr=self.snr, method_bp=(self._method == "bp"))
        else:
            self._plan = None
        return self

    # -- tflct.py:94-179 ----------------------------------------------------------------
    def forward(self, feture_bxdxtxhxw, tbes, tens):
        bnum, dnum, tnum, hnum, wnum = feture_bxdxtxhxw.shape
        for tbe, ten in zip(tbes, tens):                  # tflct.py:99-101
            assert tbe >= 0
            assert ten <= self._M
        assert hnum == wnum                               # tflct.py:113
        assert hnum == self._N                            # tflct.py:114
        tbes, tens = self._windows(tbes, tens, bnum, tnum)
        if dnum != self.dnum:
            # the reference fails here too (datapad sized by todev's dnum, tflct.py:83,133,140)
            raise RuntimeError(f"input has {dnum} channels but the layer was sized with todev(dev, dnum={self.dnum})")
        y, keys = self._run(feture_bxdxtxhxw, tbes, tens, self.fuse_minmax)
        if keys is not None:
            remember_minmax(y, keys)         # implicit hand-off to a following normalize_feature(y) call
        return y

    def forward_with_minmax(self, feture_bxdxtxhxw, tbes, tens):
        """``forward`` that also returns the per-channel min / max keys its last kernel reduced
        (``(B * D, 2)`` int64, the format ``lct_normalize_feature`` takes): the explicit form of the
        hand-off, ``normalize_feature(y, minmax=keys)``.  Keys are ``None`` for ``method='bp'``."""
        bnum, dnum, tnum, hnum, wnum = feture_bxdxtxhxw.shape
        for tbe, ten in zip(tbes, tens):
            assert tbe >= 0
            assert ten <= self._M
        assert hnum == wnum
        assert hnum == self._N
        tbes, tens = self._windows(tbes, tens, bnum, tnum)
        if dnum != self.dnum:
            raise RuntimeError(f"input has {dnum} channels but the layer was sized with todev(dev, dnum={self.dnum})")
        return self._run(feture_bxdxtxhxw, tbes, tens, True)

    def _run(self, x, tbes, tens, want_minmax):
        plan = self._ensure_plan()
        if plan is None or not x.is_cuda:
            raise RuntimeError("hiddenpose_b200 LCT runs on CUDA only (no CPU fallback): call todev('cuda', dnum) "
                               "and pass a CUDA tensor")
        if x.device != plan.device:
            raise RuntimeError(f"input is on {x.device} but the layer was moved to {plan.device}")
        if x.dtype != torch.float32:
            # the reference multiplies by float32 constants (tflct.py:127,138): any other dtype fails there too
            raise RuntimeError(f"expected a float32 input, got {x.dtype} (the layer's constants are float32, as in the reference)")
        x = x.contiguous()
        if want_minmax and self._lapw is None:
            return LctFunction.apply(x, plan, tbes, tens, True)
        return LctFunction.apply(x, plan, tbes, tens), None

    @staticmethod
    def _windows(tbes, tens, bnum, tnum):
        """The reference indexes ``tbes[i]`` for i < B (tflct.py:105-107) and NlosPose always
        passes three equal entries (NlosPose.py:53), so B > 3 raises IndexError there.
        A short list whose entries are all equal is broadcast; anything else short is an error."""
        tbes, tens = [int(t) for t in tbes], [int(t) for t in tens]
        if len(tbes) < bnum or len(tens) < bnum:
            if len(tbes) == 0 or len(tens) == 0 or len(set(tbes)) != 1 or len(set(tens)) != 1:
                raise IndexError("list index out of range: tbes/tens shorter than the batch and not uniform")
            tbes, tens = [tbes[0]] * bnum, [tens[0]] * bnum
        tbes, tens = tbes[:bnum], tens[:bnum]
        for tbe, ten in zip(tbes, tens):
            if ten - tbe != tnum:
                # torch.cat in the reference yields a time axis != crop and the .view at tflct.py:121 raises
                raise RuntimeError(f"time window [{tbe}, {ten}) does not match the input's {tnum} bins")
        return tbes, tens
