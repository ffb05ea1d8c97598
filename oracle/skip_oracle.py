"""CPU oracle for the skip branch of the reference's FeatureExtraction.

TEST INFRASTRUCTURE ONLY (same rules as ``oracle/lct_oracle.py``): nothing under
``hiddenpose_b200/`` imports this; ``tests/`` use it as the checker.

Restates /root/reference/models/feature_extraction.py:166-171 in NumPy::

    x_conv2 = F.conv3d(x, self.weights, bias=None, stride=s, padding=1, dilation=1, groups=1)
    output  = x_conv1 + x_conv2

``F.conv3d`` is a cross-correlation with zero padding: ``x_conv2[p] = sum_k w[k] * x[p + k - 1]``;
the one-channel result is broadcast over ``x_conv1``'s channels by the ``+``.  The two
vector-Jacobian products are what autograd derives for those two lines.

Parity pinning: ``tests/golden/skip_*.npz`` hold outputs and gradients of the reference module
itself (``FeatureExtraction.forward`` run by ``tests/golden/make_golden.py``).
"""
import numpy as np


def _taps():
    return [(a, b, c) for a in range(3) for b in range(3) for c in range(3)]


def skip_conv(x, w, dtype=np.float64):
    """``F.conv3d(x, w, stride=1, padding=1)`` for x (B, 1, T, H, W), w (1, 1, 3, 3, 3)."""
    x = np.asarray(x, dtype=dtype)
    w = np.asarray(w, dtype=dtype).reshape(3, 3, 3)
    B, _, T, H, W = x.shape
    xp = np.zeros((B, 1, T + 2, H + 2, W + 2), dtype=dtype)
    xp[:, :, 1:-1, 1:-1, 1:-1] = x
    out = np.zeros_like(x)
    for a, b, c in _taps():
        out += w[a, b, c] * xp[:, :, a:a + T, b:b + H, c:c + W]
    return out


def skip_sum(feat, x, w, dtype=np.float64):
    """feature_extraction.py:170: ``x_conv1 + x_conv2`` (broadcast over channels)."""
    return np.asarray(feat, dtype=dtype) + skip_conv(x, w, dtype)


def skip_sum_vjp(g, x, w, dtype=np.float64):
    """(d/d x, d/d w) of ``sum(g * skip_sum(feat, x, w))``; d/d feat is ``g``."""
    g = np.asarray(g, dtype=dtype)
    x = np.asarray(x, dtype=dtype)
    w = np.asarray(w, dtype=dtype).reshape(3, 3, 3)
    B, _, T, H, W = x.shape
    G = g.sum(axis=1, keepdims=True)                       # the broadcast becomes a sum
    Gp = np.zeros((B, 1, T + 2, H + 2, W + 2), dtype=dtype)
    Gp[:, :, 1:-1, 1:-1, 1:-1] = G
    xp = np.zeros_like(Gp)
    xp[:, :, 1:-1, 1:-1, 1:-1] = x
    gx = np.zeros_like(x)
    gw = np.zeros((3, 3, 3), dtype=dtype)
    for a, b, c in _taps():
        gx += w[a, b, c] * Gp[:, :, 2 - a:2 - a + T, 2 - b:2 - b + H, 2 - c:2 - c + W]
        gw[a, b, c] = np.sum(G * xp[:, :, a:a + T, b:b + H, c:c + W])
    return gx, gw.reshape(1, 1, 3, 3, 3)
