"""Drop-in for /root/reference/models/feature_propagation.py: ``FeaturePropagation``,
``LCT``, ``normalize``, ``normalize_feature`` and ``VisibleNet`` -- the names
models/NlosPose.py:7 imports."""
import numpy as np
import torch
import torch.nn as nn

from .layer import LctLayerBase


class LCT(LctLayerBase):
    """``LCT(image_size, time_size, bin_len, wall_size, mode, material)`` -- feature_propagation.py:48-69."""

    def __init__(self, image_size=256, time_size=128, bin_len=0.01, wall_size=2.0, mode="lct", material="diffuse"):
        super().__init__()
        self.image_size = image_size
        self.time_size = time_size
        assert 2 ** int(np.log2(self.time_size)) == time_size, \
            "time size should be a power of 2"                       # feature_propagation.py:61-62
        self.bin_len = bin_len
        self.wall_size = wall_size
        self.mode = mode
        self.material = material
        self._parpareparam()

    def _parpareparam(self):
        self._setup(self.image_size, self.time_size, self.bin_len, self.wall_size, self.mode, self.material)
        # the reference ends with todev('cuda', 1) (feature_propagation.py:109); FeaturePropagation
        # immediately overrides it with its own (dev, dnum), so the plan is created there.


class FeaturePropagation(nn.Module):
    """feature_propagation.py:18-44.
    Input : Tensor (batch_size, channels, time_size, image_size[0], image_size[1])
    Output: Tensor (batch_size, channels, time_size, image_size[0], image_size[1])
    """

    def __init__(self, image_size=256, time_size=512, bin_len=0.01, wall_size=2.0, mode="lct",
                 material="diffuse", dnum=1, dev="cpu"):
        super().__init__()
        assert mode == "lct", f"{mode} is not spported. Feature propagation only support lct by now"
        self.method = LCT(int(image_size), time_size, bin_len, wall_size, mode=mode, material=material)
        self.method.todev(dev, dnum)
        self.method.fuse_minmax = True       # the caller's next op is normalize_feature (NlosPose.py:54)

    def forward(self, x, time_begin, time_end):
        return self.method(x, time_begin, time_end)

    def forward_normalized(self, x, time_begin, time_end):
        """``normalize_feature(self(x, time_begin, time_end))`` -- the two lines NlosPose.py:53-54 -- with the
        min / max passed explicitly from the LCT's last kernel to the affine pass (no state left on tensors)."""
        y, keys = self.method.forward_with_minmax(x, time_begin, time_end)
        return normalize_feature(y, minmax=keys)


def _native_normalize(x, scale, minmax=None):
    from .lct_function import NormalizeFeatureFunction, recall_minmax
    return NormalizeFeatureFunction.apply(x.contiguous(), scale, minmax if minmax is not None else recall_minmax(x))


def normalize(data_bxcxdxhxw):
    """feature_propagation.py:260-270: per (b, c) min/max normalisation to [0, 1]."""
    if data_bxcxdxhxw.is_cuda and data_bxcxdxhxw.dtype == torch.float32 and data_bxcxdxhxw.dim() == 5:
        return _native_normalize(data_bxcxdxhxw, 1.0)
    b, c, d, h, w = data_bxcxdxhxw.shape
    flat = data_bxcxdxhxw.reshape(b, c, -1)
    shifted = flat - flat.min(2, keepdim=True)[0]
    return (shifted / (shifted.max(2, keepdim=True)[0] + 1e-15)).view(b, c, d, h, w)


def normalize_feature(data_bxcxdxhxw, minmax=None):
    """feature_propagation.py:273-286: min/max normalisation times 10.

    The reference calls ``nn.ReLU()(x)`` and discards the result (line 274), so
    negative LCT values do reach the ``min``; that behaviour is kept.  CUDA float32 volumes go through
    the library (``lct_normalize_feature``).  ``minmax`` (optional, not in the reference's signature) is the
    key tensor ``LCT.forward_with_minmax`` returned for exactly this volume; without it the keys the layer
    remembered for this tensor object are used if it is untouched since, else one reduction pass finds them.
    """
    if data_bxcxdxhxw.is_cuda and data_bxcxdxhxw.dtype == torch.float32 and data_bxcxdxhxw.dim() == 5:
        return _native_normalize(data_bxcxdxhxw, 10.0, minmax)
    return normalize(data_bxcxdxhxw) * 10.0


class VisibleNet(nn.Module):
    """feature_propagation.py:289-312 (only used by the 2-D backbone)."""

    def __init__(self, basedim, layernum=0):
        super().__init__()
        self.layernum = layernum

    def forward(self, x):
        x = normalize(torch.relu(x)) * 1.0e5
        depdim = x.shape[2]
        val, dep = x.topk(4, dim=2)
        dep = (depdim - 1 - dep.float()) / (depdim - 1)
        return torch.cat([val, dep], dim=1)
