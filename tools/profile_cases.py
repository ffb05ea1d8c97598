#!/usr/bin/env python
"""Launch sequence for profiler captures (GPU box): `python tools/profile_cases.py <workload> [--fp] [--bwd]`.
Two warm staged runs, then ONE staged run of the whole batch on a single stream (lct_run_staged: K1, [plane | K2, K3, K4],
K5 back to back) -- the run to capture: with `-k regex:lct_kernel -s <2 x kernels per run> -c <kernels per run>`.
`--fp` runs the layer as FeaturePropagation does (min/max reduced in the last kernel) through the public call instead."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import hiddenpose_b200 as hp  # noqa: E402
from bench import WORKLOADS, bin_len_for  # noqa: E402


def main():
    name = sys.argv[1]
    fp, bwd = "--fp" in sys.argv, "--bwd" in sys.argv
    B, M, N, _ = WORKLOADS[name]
    dev = torch.device("cuda", 0)
    os.environ.setdefault("LCT_STREAM_GROUPS", "1")          # whole-batch launches, one stream
    layer = hp.lct(spatial=N, crop=M, bin_len=bin_len_for(M))
    layer.todev(dev, 1)
    layer.fuse_minmax = fp
    torch.manual_seed(410)
    x = torch.rand(B, 1, M, N, N, device=dev)
    g = torch.randn(B, 1, M, N, N, device=dev)
    tb, te = [0] * B, [M] * B
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    with torch.no_grad():
        for _ in range(3):
            flush.zero_()
            if fp:
                layer(x, tb, te)
            else:
                ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
                for e in ev:
                    e.record()
                layer._plan.run_staged(g if bwd else x, tb, te, M, ev, backward=bwd)
        torch.cuda.synchronize()
    print("done", name, "fp" if fp else "", "bwd" if bwd else "")


if __name__ == "__main__":
    main()
