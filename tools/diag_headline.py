#!/usr/bin/env python
"""Why bench.py's 20-step mean sat 16 us above the steady-state step (GPU box): prints the per-step CUDA-event times of the
headline loop with and without the NVML clock sampler, for 20 and 200 steps.  Finding (round 2): the first step after the
barrier + synchronize takes 300-490 us because the launch queue is empty and the device waits on the host; bench.py now
queues a ~1 ms memset pre-roll after the barrier."""
import os, sys, statistics, time
sys.path.insert(0, '/root/repo')
import torch
import hiddenpose_b200 as hp
from bench import ClockSampler, bin_len_for
dev = torch.device('cuda', 0)
B, M, N = 8, 256, 64
layer = hp.lct(spatial=N, crop=M, bin_len=bin_len_for(M)); layer.todev(dev, 1)
x = torch.rand(B, 1, M, N, N, device=dev); tb, te = [0]*B, [M]*B
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
def run(K, sampler_on, W=5):
    with torch.no_grad():
        for _ in range(W): layer(x, tb, te)
        torch.cuda.synchronize()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        s = ClockSampler(0)
        if sampler_on: s.start()
        t0 = time.perf_counter()
        for a, b in evs:
            flush.zero_(); a.record(); layer(x, tb, te); b.record()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        if sampler_on: s.stop()
    ms = [a.elapsed_time(b) for a, b in evs]
    print(f"K={K} sampler={sampler_on}: mean {statistics.fmean(ms)*1e3:.1f} median {statistics.median(ms)*1e3:.1f} min {min(ms)*1e3:.1f} max {max(ms)*1e3:.1f} us; host issue {1e6*(t1-t0)/K:.0f} us/step; first5 {[round(v*1e3) for v in ms[:5]]}")
for K in (20, 20, 200):
    run(K, False); run(K, True)
