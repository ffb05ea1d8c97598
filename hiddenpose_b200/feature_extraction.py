"""Drop-in for ``models/feature_extraction.py::FeatureExtraction`` -- the module that produces the
LCT's input (NlosPose.py:19-23,51).

The reference's forward (feature_extraction.py:160-171) is::

    x_conv1 = self.conv1(x)                                   # learned: pad + conv + two residual blocks
    x_conv2 = F.conv3d(x, self.weights, stride=s, padding=1)  # learnable 3x3x3 kernel, one channel
    return x_conv1 + x_conv2

``conv1`` is an ordinary learned network and stays on torch's own layers (out of this library's
scope).  The skip branch and the sum -- the traffic right in front of the LCT -- run as one CUDA
stencil pass (``lct_skip_sum``) with an explicit backward, whenever the call is the one NlosPose
makes: CUDA float32, ``stride == 1``, one input channel, square frames with ``N % 4 == 0``.
Parameter names and shapes match the reference, so its checkpoints load with ``strict=True``.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .lct_function import SkipSumFunction


def _padded_conv(cin, cout, stride):
    return [nn.ReplicationPad3d(1), nn.Conv3d(cin, cout, kernel_size=3, padding=0, stride=stride, bias=True)]


class ResConv3D(nn.Module):
    """Residual block of the learned branch (feature_extraction.py:228-256): ``leaky(tmp(x) + x)``."""

    def __init__(self, basedim, inplace=False):
        super().__init__()
        self.inplace = inplace
        self.tmp = nn.Sequential(*_padded_conv(basedim, basedim, 1),
                                 nn.LeakyReLU(negative_slope=0.2, inplace=inplace),
                                 *_padded_conv(basedim, basedim, 1))

    def forward(self, x):
        return F.leaky_relu(self.tmp(x) + x, negative_slope=0.2, inplace=self.inplace)


def skip_sum(feat, x, weights, stride=1):
    """``feat + F.conv3d(x, weights, stride, padding=1)``; the CUDA stencil when the shapes are NlosPose's."""
    native = (feat.is_cuda and x.is_cuda and weights.is_cuda and stride == 1
              and feat.dtype == x.dtype == weights.dtype == torch.float32
              and x.dim() == 5 and x.shape[1] == 1 and tuple(weights.shape) == (1, 1, 3, 3, 3)
              and x.shape[3] == x.shape[4] and x.shape[4] % 4 == 0
              and feat.shape[0] == x.shape[0] and feat.shape[2:] == x.shape[2:])
    if native:
        return SkipSumFunction.apply(feat, x, weights)
    return feat + F.conv3d(x, weights, bias=None, stride=stride, padding=1, dilation=1, groups=1)


class FeatureExtraction(nn.Module):
    """(B, 1, T, H, W) -> (B, basedim, T/s, H/s, W/s); same constructor as feature_extraction.py:130-158."""

    def __init__(self, basedim, in_channels, stride=2, norm=nn.InstanceNorm3d):
        super().__init__()
        assert in_channels == 1, f'input channels should be 1, not {in_channels}'        # feature_extraction.py:137-138
        self.stride = stride
        box = np.zeros((1, 1, 3, 3, 3), dtype=np.float32)                                # feature_extraction.py:141-145:
        box[:, :, 1:, 1:, 1:] = 1.0                                                      # mean of the 2x2x2 forward corner
        self.weights = nn.Parameter(torch.from_numpy(box / np.sum(box)))
        self.conv1 = nn.Sequential(*_padded_conv(in_channels, basedim, stride),
                                   ResConv3D(basedim, inplace=False), ResConv3D(basedim, inplace=False))

    def forward(self, x):
        return skip_sum(self.conv1(x), x, self.weights, self.stride)
