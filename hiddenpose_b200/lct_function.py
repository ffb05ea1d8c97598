"""Device plan + autograd.Function over the C ABI (include/hiddenpose_lct.h).

Replaces the body of /root/reference/models/tflct.py:94-179 and the backward
autograd derives from it.  The layer is a fixed linear operator, so nothing is
saved for backward except the integer window.
"""
from __future__ import annotations

import ctypes
import weakref

import numpy as np
import torch

from . import _native


# -- implicit min/max hand-off ------------------------------------------------------------------
# FeaturePropagation's caller applies normalize_feature to the very tensor the layer returned
# (NlosPose.py:53-54).  The layer's last kernel has already reduced that tensor's per-channel min / max;
# they are remembered here per tensor *object* (weak reference: the entry dies with the tensor, so a
# recycled address can never match) together with the tensor's version counter and address, so any
# in-place write, view, clone or copy simply misses and normalize_feature runs its own reduction.
# Callers that want no implicit state use ``forward_with_minmax`` / ``normalize_feature(y, minmax=keys)``.
_minmax_registry = {}


def remember_minmax(y, keys):
    key = id(y)

    def forget(_ref, key=key):
        _minmax_registry.pop(key, None)

    _minmax_registry[key] = (weakref.ref(y, forget), keys, y._version, y.data_ptr(), tuple(y.shape))


def recall_minmax(y):
    entry = _minmax_registry.get(id(y))
    if entry is None:
        return None
    ref, keys, version, ptr, shape = entry
    if ref() is not y or y._version != version or y.data_ptr() != ptr or tuple(y.shape) != shape or not y.is_contiguous():
        return None
    return keys


def _i32_array(values):
    arr = (ctypes.c_int32 * len(values))(*[int(v) for v in values])
    return arr


class LctPlan:
    """Owns one ``lct_plan*`` (immutable device constants) on one CUDA device."""

    def __init__(self, M, N, csr, falloff, filter_half, device, lapw=None, workspace_limit_bytes=16 << 30, flags=0,
                 psf=None, snr=0.1, method_bp=False):
        if device.type != "cuda":
            raise RuntimeError("LctPlan needs a CUDA device; there is no CPU implementation of this layer")
        self.lib = _native.load()
        self.M, self.N, self.device = int(M), int(N), device
        self.lapw = None if lapw is None else np.ascontiguousarray(lapw, dtype=np.float32).ravel()
        self.workspace_limit_bytes = int(workspace_limit_bytes)
        rowptr = np.ascontiguousarray(csr[0], dtype=np.int32)
        colidx = np.ascontiguousarray(csr[1], dtype=np.int32)
        vals = np.ascontiguousarray(csr[2], dtype=np.float32)
        # either the host-built half spectrum, or the PSF support from which the library builds it on the GPU
        filt = psf_z = psf_yx = None
        if filter_half is not None:
            filt = np.ascontiguousarray(filter_half, dtype=np.complex64)
            if filt.shape != (M + 1, 2 * N, 2 * N):
                raise ValueError(f"filter_half must be {(M + 1, 2 * N, 2 * N)}, got {filt.shape}")
        elif psf is not None:
            z, y, x, val = psf
            psf_z = np.ascontiguousarray(z, dtype=np.int32)
            psf_yx = np.ascontiguousarray(np.asarray(y, dtype=np.int64) * (2 * N) + np.asarray(x, dtype=np.int64), dtype=np.int32)
        else:
            raise ValueError("LctPlan needs filter_half or psf")
        fall = None if falloff is None else np.ascontiguousarray(falloff, dtype=np.float32)
        f32p, i32p = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int32)
        desc = _native.LctDesc(
            time_bins=M, spatial=N, device=device.index if device.index is not None else torch.cuda.current_device(),
            reserved=int(flags), mtx_rowptr=rowptr.ctypes.data_as(i32p), mtx_colidx=colidx.ctypes.data_as(i32p),
            mtx_vals=vals.ctypes.data_as(f32p), falloff=None if fall is None else fall.ctypes.data_as(f32p),
            filter_half=None if filt is None else filt.view(np.float32).ctypes.data_as(f32p),
            psf_z=None if psf_z is None else psf_z.ctypes.data_as(i32p),
            psf_yx=None if psf_yx is None else psf_yx.ctypes.data_as(i32p),
            psf_count=0 if psf_z is None else len(psf_z), psf_value=0.0 if psf is None else float(psf[3]),
            snr=float(snr), method_bp=1 if method_bp else 0)
        handle = ctypes.c_void_p()
        _native.check(self.lib.lct_plan_create(ctypes.byref(desc), ctypes.byref(handle)))
        del rowptr, colidx, vals, fall, filt, psf_z, psf_yx      # host tables were copied to the device
        self.handle = handle

    def __del__(self):
        h, self.handle = getattr(self, "handle", None), None
        if h:
            try:
                self.lib.lct_plan_destroy(h)
            except Exception:
                pass

    # ------------------------------------------------------------------
    def workspace(self, channels):
        per = self.lib.lct_plan_workspace_bytes(self.handle, 1)
        head = 2 * per - self.lib.lct_plan_workspace_bytes(self.handle, 2)       # fixed header
        per_channel = per - head
        fit = max(1, (self.workspace_limit_bytes - head) // per_channel)
        nbytes = head + per_channel * min(int(channels), int(fit))
        return torch.empty(nbytes, dtype=torch.uint8, device=self.device), nbytes

    def _run(self, fn, src, tbes, tens, B, D, Tin, dst):
        ws, nbytes = self.workspace(B * D)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _native.check(fn(self.handle, src.data_ptr(), _i32_array(tbes), _i32_array(tens), B, D, Tin,
                         dst.data_ptr(), ws.data_ptr(), nbytes, stream))

    def forward(self, x, tbes, tens, want_minmax=False):
        """Returns the volume, and -- if asked -- the per-channel {min, max} keys the last kernel reduced
        while writing it (consumed by normalize_feature; see include/hiddenpose_lct.h)."""
        B, D, Tin, H, W = x.shape
        y = torch.empty((B, D, self.M, H, W), dtype=torch.float32, device=x.device)
        with torch.cuda.device(self.device):
            if not want_minmax:
                self._run(self.lib.lct_forward, x, tbes, tens, B, D, Tin, y)
                return y
            keys = torch.empty((B * D, 2), dtype=torch.int64, device=x.device)
            ws, nbytes = self.workspace(B * D)
            stream = torch.cuda.current_stream(self.device).cuda_stream
            _native.check(self.lib.lct_forward_minmax(self.handle, x.data_ptr(), _i32_array(tbes), _i32_array(tens),
                                                      B, D, Tin, y.data_ptr(), keys.data_ptr(), ws.data_ptr(), nbytes, stream))
        return y, keys

    def backward(self, gy, tbes, tens, Tin):
        B, D, M, H, W = gy.shape
        gx = torch.empty((B, D, Tin, H, W), dtype=torch.float32, device=gy.device)
        with torch.cuda.device(self.device):
            self._run(self.lib.lct_backward, gy, tbes, tens, B, D, Tin, gx)
        return gx

    def run_staged(self, src, tbes, tens, Tin, events, backward=False):
        """Measurement hook (lct_run_staged): forward or backward with six torch CUDA events
        recorded around the five kernels.  Returns the output tensor."""
        B, D = src.shape[0], src.shape[1]
        out_t = Tin if backward else self.M
        dst = torch.empty((B, D, out_t, self.N, self.N), dtype=torch.float32, device=src.device)
        ws, nbytes = self.workspace(B * D)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        handles = (ctypes.c_void_p * 6)(*[ctypes.c_void_p(e.cuda_event) for e in events])
        with torch.cuda.device(self.device):
            _native.check(self.lib.lct_run_staged(self.handle, src.data_ptr(), _i32_array(tbes), _i32_array(tens),
                                                  B, D, Tin, dst.data_ptr(), ws.data_ptr(), nbytes, stream,
                                                  1 if backward else 0, handles))
        return dst

    def laplacian(self, vol, adjoint):
        out = torch.empty_like(vol)
        C = vol.shape[0] * vol.shape[1]
        stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            _native.check(self.lib.lct_bp_laplacian(
                self.handle, vol.data_ptr(), out.data_ptr(), C,
                self.lapw.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), 1 if adjoint else 0, stream))
        return out

    def forward_host(self, x_host, tbes, tens):
        """Host-buffer entry point (lct_forward_host): H2D, forward, D2H, sync."""
        B, D, Tin, H, W = x_host.shape
        y_host = torch.empty((B, D, self.M, H, W), dtype=torch.float32, pin_memory=x_host.is_pinned())
        stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            _native.check(self.lib.lct_forward_host(self.handle, x_host.data_ptr(), _i32_array(tbes), _i32_array(tens),
                                                    B, D, Tin, y_host.data_ptr(), stream))
        return y_host


class LctFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, plan, tbes, tens, want_minmax=False):
        ctx.plan, ctx.tbes, ctx.tens, ctx.tin = plan, tuple(tbes), tuple(tens), x.shape[2]
        if want_minmax and plan.lapw is None:
            y, keys = plan.forward(x, tbes, tens, want_minmax=True)
            ctx.mark_non_differentiable(keys)
            return y, keys
        y = plan.forward(x, tbes, tens)
        if plan.lapw is not None:                       # method == 'bp' (tflct.py:164-175)
            y = plan.laplacian(y, adjoint=False)
        return y

    @staticmethod
    def backward(ctx, gy, *unused):
        if not ctx.needs_input_grad[0]:
            return None, None, None, None, None
        plan = ctx.plan
        gy = gy.contiguous().float()
        if plan.lapw is not None:
            gy = plan.laplacian(gy, adjoint=True)
        return plan.backward(gy, ctx.tbes, ctx.tens, ctx.tin), None, None, None, None


class NormalizeFeatureFunction(torch.autograd.Function):
    """feature_propagation.py:260-286 on the CUDA library: ``(x - min) / (max(x - min) + 1e-15) * scale`` per
    (batch, channel) volume, forward and backward.  ``keys`` = the per-channel min / max keys of exactly this
    tensor when the LCT's last kernel already reduced them; ``None`` makes one reduction pass find them."""

    @staticmethod
    def forward(ctx, x, scale, keys):
        lib = _native.load()
        b, c = x.shape[0], x.shape[1]
        elems = x.numel() // (b * c)
        stream = torch.cuda.current_stream(x.device).cuda_stream
        with torch.cuda.device(x.device):
            if keys is not None:
                if keys.shape != (b * c, 2) or keys.dtype != torch.int64 or keys.device != x.device:
                    raise ValueError("minmax keys must be the (B*D, 2) int64 tensor forward_with_minmax returned")
            else:
                keys = torch.empty((b * c, 2), dtype=torch.int64, device=x.device)
                _native.check(lib.lct_minmax(x.data_ptr(), b * c, elems, keys.data_ptr(), stream))
            out = torch.empty_like(x)
            _native.check(lib.lct_normalize_feature(x.data_ptr(), keys.data_ptr(), out.data_ptr(), b * c, elems, float(scale), stream))
        ctx.save_for_backward(x, keys)
        ctx.scale = float(scale)
        return out

    @staticmethod
    def backward(ctx, gout):
        x, keys = ctx.saved_tensors
        lib = _native.load()
        b, c = x.shape[0], x.shape[1]
        elems = x.numel() // (b * c)
        gout = gout.contiguous().float()
        gx = torch.empty_like(x)
        sums = torch.empty((b * c, 2), dtype=torch.float64, device=x.device)
        stream = torch.cuda.current_stream(x.device).cuda_stream
        with torch.cuda.device(x.device):
            _native.check(lib.lct_normalize_feature_backward(x.data_ptr(), gout.data_ptr(), keys.data_ptr(), gx.data_ptr(),
                                                             sums.data_ptr(), b * c, elems, ctx.scale, stream))
        return gx, None, None


class SkipSumFunction(torch.autograd.Function):
    """``feat + F.conv3d(x, weights, stride=1, padding=1)`` (feature_extraction.py:166-171) on the CUDA library.

    ``feat`` (B, D, T, N, N), ``x`` (B, 1, T, N, N), ``weights`` (1, 1, 3, 3, 3), all CUDA float32.
    Backward: d/d feat is the incoming gradient itself; d/d x and d/d weights are one stencil pass each.
    """

    @staticmethod
    def forward(ctx, feat, x, weights):
        lib = _native.load()
        feat, x, w = feat.contiguous(), x.contiguous(), weights.detach().contiguous()
        B, D, T, N = feat.shape[0], feat.shape[1], feat.shape[2], feat.shape[3]
        out = torch.empty_like(feat)
        with torch.cuda.device(feat.device):
            stream = torch.cuda.current_stream(feat.device).cuda_stream
            _native.check(lib.lct_skip_sum(feat.data_ptr(), x.data_ptr(), w.data_ptr(), B, D, T, N, out.data_ptr(), stream))
        ctx.save_for_backward(x, w)
        ctx.dims = (B, D, T, N)
        return out

    @staticmethod
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        lib = _native.load()
        B, D, T, N = ctx.dims
        g = g.contiguous().float()
        need_x, need_w = ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        gx = torch.empty_like(x) if need_x else None
        gw = torch.empty(27, dtype=torch.float32, device=x.device) if need_w else None
        if need_x or need_w:
            with torch.cuda.device(x.device):
                nbytes = lib.lct_skip_workspace_bytes(B, T, N) if need_w else 0
                ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=x.device)
                stream = torch.cuda.current_stream(x.device).cuda_stream
                _native.check(lib.lct_skip_sum_backward(
                    g.data_ptr(), x.data_ptr(), w.data_ptr(), B, D, T, N,
                    gx.data_ptr() if need_x else None, gw.data_ptr() if need_w else None,
                    ws.data_ptr(), nbytes, stream))
        return (g if ctx.needs_input_grad[0] else None), gx, (gw.view(1, 1, 3, 3, 3) if need_w else None)
