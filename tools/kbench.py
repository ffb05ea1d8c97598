#!/usr/bin/env python
"""Per-kernel timing of the chain for kernel work (GPU box): `python tools/kbench.py [cfg2 cfg3 ...] [--bwd] [--reps 40]`.
Prints one compact line per workload: per-stage median microseconds from lct_run_staged (single stream, L2 flushed
before every run) and the headline step (public call, two stream groups).  Library: HIDDENPOSE_LCT_LIB or the in-tree .so."""
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import hiddenpose_b200 as hp  # noqa: E402
from bench import WORKLOADS, bin_len_for, chain_bytes, stage_bytes, measured_peak  # noqa: E402


def main():
    names = [a for a in sys.argv[1:] if a in WORKLOADS or a.startswith("custom:")] or ["cfg2"]     # custom:B,M,N
    bwd = "--bwd" in sys.argv
    reps = int(sys.argv[sys.argv.index("--reps") + 1]) if "--reps" in sys.argv else 40
    dev = torch.device("cuda", 0)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    peak = measured_peak()[0]
    tag = os.path.basename(os.environ.get("HIDDENPOSE_LCT_LIB", "in-tree"))
    for name in names:
        B, M, N = [int(v) for v in name[7:].split(",")] if name.startswith("custom:") else WORKLOADS[name][:3]
        fuse = "--fp" in sys.argv
        layer = hp.lct(spatial=N, crop=M, bin_len=bin_len_for(M))
        layer.todev(dev, 1)
        layer.fuse_minmax = fuse
        plan = layer._plan
        torch.manual_seed(410)
        x = torch.rand(B, 1, M, N, N, device=dev)
        g = torch.randn(B, 1, M, N, N, device=dev)
        tb, te = [0] * B, [M] * B
        src = g if bwd else x
        with torch.no_grad():
            for _ in range(3):
                plan.run_staged(src, tb, te, M, [torch.cuda.Event(enable_timing=True) for _ in range(6)], backward=bwd) \
                    if False else None
            evs = []
            for _ in range(reps + 3):
                e = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
                for q in e:
                    q.record()
                flush.zero_()
                plan.run_staged(src, tb, te, M, e, backward=bwd)
                evs.append(e)
            torch.cuda.synchronize()
            evs = evs[3:]
            st = [statistics.median(e[j].elapsed_time(e[j + 1]) for e in evs) * 1e3 for j in range(5)]
            pairs = []
            fn = (lambda: plan.backward(g, tb, te, M)) if bwd else (lambda: layer(x, tb, te))
            for _ in range(reps + 3):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                flush.zero_()
                a.record()
                fn()
                b.record()
                pairs.append((a, b))
            torch.cuda.synchronize()
            head = statistics.median(a.elapsed_time(b) for a, b in pairs[3:]) * 1e3
        sb = stage_bytes(M, N, B)
        fr = [sb[j] / (st[j] * 1e-6) / 1e9 / peak if st[j] > 1 else 0 for j in range(5)]
        A = chain_bytes(M, N, B)
        print(f"[{tag}] {name}{' bwd' if bwd else ''}{' fp' if fuse else ''}: K1 {st[0]:.1f} ({fr[0]:.2f})  K2 {st[1]:.1f} ({fr[1]:.2f})  K3 {st[2]:.1f} ({fr[2]:.2f})  "
              f"K4 {st[3]:.1f} ({fr[3]:.2f})  K5 {st[4]:.1f} ({fr[4]:.2f})  serial {sum(st):.1f}  step {head:.1f} us  "
              f"chain {A / (head * 1e-6) / 1e9 / peak:.3f}", flush=True)
        del layer, plan, x, g
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
