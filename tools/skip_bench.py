#!/usr/bin/env python
"""Time the skip branch (lct_skip_sum / lct_skip_sum_backward) against the torch expression the
reference runs (feature_extraction.py:166-171) at a BASELINE shape.  GPU box only.

    python tools/skip_bench.py [B T N]      # default 8 256 64 (cfg2)
"""
import json
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from hiddenpose_b200.lct_function import SkipSumFunction   # noqa: E402


def timed(fn, reps=30, warm=5):
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(warm):
        fn()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        flush.zero_()
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return ts[len(ts) // 2] * 1e3


def main():
    B, T, N = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (8, 256, 64)
    torch.manual_seed(410)
    torch.backends.cudnn.allow_tf32 = False
    feat = torch.randn(B, 1, T, N, N, device="cuda")
    x = torch.rand(B, 1, T, N, N, device="cuda")
    g = torch.randn(B, 1, T, N, N, device="cuda")
    w = torch.randn(1, 1, 3, 3, 3, device="cuda", requires_grad=True)
    xr = x.clone().requires_grad_(True)
    V = T * N * N

    def ours_f():
        return SkipSumFunction.apply(feat, x, w)

    def torch_f():
        return feat + F.conv3d(x, w, bias=None, stride=1, padding=1)

    def ours_b():
        torch.autograd.grad(SkipSumFunction.apply(feat, xr, w), (xr, w), g)

    def torch_b():
        torch.autograd.grad(feat + F.conv3d(xr, w, bias=None, stride=1, padding=1), (xr, w), g)

    with torch.no_grad():
        t_of, t_tf = timed(ours_f), timed(torch_f)
    t_ob, t_tb = timed(ours_b), timed(torch_b)
    fwd_bytes = 12 * V * B                  # read feat, read x, write out
    res = {"shape": [B, 1, T, N, N],
           "forward_us": {"cuda_stencil": t_of, "torch_conv3d_plus_add": t_tf},
           "forward_plus_backward_us": {"cuda_stencil": t_ob, "torch_conv3d_plus_add": t_tb},
           "forward_algorithmic_bytes": fwd_bytes, "forward_gbs": fwd_bytes / (t_of * 1e-6) / 1e9}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
