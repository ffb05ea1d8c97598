#!/usr/bin/env python
"""Small forward+backward sweep for compute-sanitizer (memcheck / racecheck), one tool per GPU call:
    compute-sanitizer --tool memcheck python tools/sanitize_run.py
Covers the fused plane path (N <= 64), the five-kernel path (N = 128), the parity-split kernels (N = 256), ragged windows, D > 1 and bp.

NOT RUN SO FAR: compute-sanitizer is closed on this GPU pool (both round-1 attempts came back with the pool's refusal), so the
sweep is kept for a pool where it is open; the race evidence there is comes from the CPU thread emulator (tests/emu)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hiddenpose_b200 as hp   # noqa: E402

CASES = [(32, 8, 3, 2, "lct"), (64, 16, 2, 1, "lct"), (32, 32, 1, 1, "bp"), (64, 64, 1, 1, "lct"), (32, 128, 1, 1, "lct"), (32, 256, 1, 1, "lct")]
for M, N, B, D, method in CASES:
    layer = hp.lct(spatial=N, crop=M, bin_len=0.01 * 512 / M, method=method)
    layer.todev("cuda:0", D)
    tin = M - 5
    tbes = [i % 4 for i in range(B)]
    tens = [t + tin for t in tbes]
    x = torch.rand(B, D, tin, N, N, device="cuda", requires_grad=True)
    y = layer(x, tbes, tens)
    y.backward(torch.randn_like(y))
    torch.cuda.synchronize()
    print(M, N, B, D, method, "ok", float(y.abs().max()), float(x.grad.abs().max()))
