#!/usr/bin/env python
"""Mint golden vectors for the skip branch by running the REFERENCE's FeatureExtraction itself.

    python tests/golden/make_golden_skip.py            # build container only: needs /root/reference

Imports ``/root/reference/models/feature_extraction.py`` unmodified, builds
``FeatureExtraction(basedim=D, in_channels=1, stride=1)`` as NlosPose does (NlosPose.py:19-23),
replaces its parameters by seeded random values (the released initial kernel has only eight
non-zero taps), and records what its own ``forward`` (feature_extraction.py:160-171) and autograd
produce:

    feat      = conv1(x), captured by a forward hook          (the learned branch's output)
    out       = forward(x) = feat + F.conv3d(x, weights, padding=1)
    gw        = weights.grad                                  (only the skip branch touches it)
    gx_total  = x.grad through both branches
    gx_conv1  = x.grad through ``module.conv1`` alone         (so gx_total - gx_conv1 is the skip branch's)

Inputs are regenerated in the tests from the recorded seeds; the parameters travel in the file.
"""
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))

# (name, B, D, T, N)
CASES = [
    ("skip_b2d1t16n8", 2, 1, 16, 8),
    ("skip_b1d2t25n16", 1, 2, 25, 16),       # broadcast over two channels; T crosses one 24-bin chunk
    ("skip_b1d1t27n32", 1, 1, 27, 32),
]


def main():
    if not os.path.isdir(REF):
        raise SystemExit("reference tree not present; golden vectors can only be minted in the build container")
    sys.path.insert(0, REF)
    from models.feature_extraction import FeatureExtraction as RefFE      # noqa: E402
    torch.set_num_threads(8)
    for seed, (name, B, D, T, N) in enumerate(CASES, start=100):
        torch.manual_seed(seed)
        m = RefFE(basedim=D, in_channels=1, stride=1)
        with torch.no_grad():
            for prm in m.parameters():
                prm.copy_(torch.randn_like(prm) * 0.2)
        rs = np.random.RandomState(seed)
        x_np = rs.rand(B, 1, T, N, N).astype(np.float32)
        g_np = rs.randn(B, D, T, N, N).astype(np.float32)
        grabbed = {}
        hook = m.conv1.register_forward_hook(lambda mod, inp, out: grabbed.__setitem__("feat", out.detach().clone()))
        x = torch.from_numpy(x_np).requires_grad_(True)
        out = m(x)
        hook.remove()
        out.backward(torch.from_numpy(g_np))
        x1 = torch.from_numpy(x_np).requires_grad_(True)
        (gx_conv1,) = torch.autograd.grad(m.conv1(x1), x1, torch.from_numpy(g_np))
        rec = dict(seed=np.int64(seed), B=np.int64(B), D=np.int64(D), T=np.int64(T), N=np.int64(N),
                   feat=grabbed["feat"].numpy(), out=out.detach().numpy(), gw=m.weights.grad.numpy(),
                   gx_total=x.grad.numpy(), gx_conv1=gx_conv1.numpy())
        for k, v in m.state_dict().items():
            rec["param:" + k] = v.numpy()
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **rec)
        print(name, "out", float(out.norm()), "gw", float(m.weights.grad.norm()), "keys", sorted(m.state_dict().keys()))


if __name__ == "__main__":
    main()
