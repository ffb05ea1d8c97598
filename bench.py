#!/usr/bin/env python
"""Throughput of the LCT hot path on B200 (BASELINE.json metric: LCT transients/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload ...]

One "step" = one LCT forward over one batch of synthetic transients.  The headline workload is
BASELINE.json configs[1]'s LCT part: batch 8 x 1 x 256 x 64 x 64 per GPU (`--workload cfg3|cfg4|cfg5`
selects the other shapes).  N > 1 shards by transient: every rank runs the same per-GPU batch (weak
scaling), no data-path collective.

Prints ONE JSON line (rank 0).  `value` is device-resident throughput; `e2e` is the same metric through
the public module API with pinned-host input and host output inside the timed region; `roofline` is the
dominant kernel's algorithmic bytes over its CUDA-event duration against MEASURED_PEAKS.json;
`cpu_baseline` is the reference layer itself (baseline/_ref, else the oracle port) on the host cores.

Two more records ride on every line, because BASELINE.json names them and the driver only runs the
default command:

* `strong_cfg3`  -- configs[2]: ONE batch of 64 x 512x128x128 transients sharded over the N ranks
                    (`sharding.shard_batch`: 64 / 32 / 16 / 8 per GPU), forward and forward+backward;
* `train_cfg4`   -- configs[3]: the reference training step's LCT neighbourhood at 16 x 128^3 per GPU
                    (FeatureExtraction -> FeaturePropagation -> normalize_feature, forward + backward +
                    optimizer step) under DistributedDataParallel, with a stand-in parameter block the size
                    of NlosPose's 88 263 656 weights so that NCCL all-reduces the 353 MB of gradients the
                    real model would (utils/train_epoch.py:37-38,74-76 at the train.py:77-86 shape).

`--workload cfg4_train` runs only the training harness and prints its own line.
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np   # noqa: E402,F401
import torch         # noqa: E402

WORKLOADS = {
    # name: (per-GPU batch, M, N, description)
    "cfg1": (1, 256, 64, "LCT forward, one 1 x 1 x 256x64x64 transient (BASELINE.json configs[0], the reference's CPU-runnable case)"),
    "cfg2": (8, 256, 64, "LCT forward, batch 8 x 1 x 256x64x64 per GPU (BASELINE.json configs[1], LCT part)"),
    "cfg3": (8, 512, 128, "LCT forward, batch 8 x 1 x 512x128x128 per GPU (configs[2] at 8 GPUs)"),
    "cfg4": (16, 128, 128, "LCT forward, batch 16 x 1 x 128x128x128 per GPU (configs[3], LCT part)"),
    "cfg5": (1, 512, 256, "LCT forward, single 512x256x256 transient (configs[4])"),
    "tiny": (2, 64, 16, "LCT forward, batch 2 x 1 x 64x16x16 (harness self-test)"),
    "cfg4_train": (16, 128, 128, "training step around the LCT, batch 16 x 1 x 128x128x128 per GPU, DDP + NCCL gradient all-reduce (configs[3])"),
}
STAGES = ("time_fwd", "row_fwd", "col_filter", "row_inv", "time_inv")
POSE_NET_PARAMETERS = 88_263_656            # NlosPose's trainable weights: 353 MB of fp32 gradients
NOMINAL_HBM_GBS = 8000.0                    # the figure BASELINE.json's north_star quotes


def bin_len_for(M):
    return 0.01 * 512 / M          # trange = 5.12 as in the released configs


def stage_bytes(M, N, C):
    """Algorithmic bytes per launch of each kernel (DESIGN.md 'Kernels')."""
    V = M * N * N
    return [12 * V * C, 24 * V * C, 32 * V * C + 32 * V, 24 * V * C, 12 * V * C]


def chain_bytes(M, N, C):
    """A = 104 V C + 32 V, the contract figure of SURVEY.md 8d for one direction."""
    return sum(stage_bytes(M, N, C))


def chain_flops(M, N, C):
    """fp32 operations of the half-spectrum dataflow at 5 n log2 n per complex n-point line (half for the real
    T-axis lines), unpruned: per voxel 2 * (5 log2 2M + 10 log2 2N + 20 log2 2N) + 24 for the filter product."""
    V = M * N * N
    return C * V * (10 * np.log2(2 * M) + 60 * np.log2(2 * N) + 24)


def make_config(desc, B, M, N):
    """Identical in both arms (`--impl ours` and `--impl reference`)."""
    return {"workload": desc, "B_per_gpu": B, "M": M, "N": N, "seed": 410,
            "l2": "GPU arm: flushed between timed steps (512 MiB write outside the events); CPU reference arm: not applicable",
            "timing": "GPU arm: CUDA events per step on the launch stream, sum over K steps, max over ranks (a ~1 ms memset pre-roll after the "
                      "barrier lets the host run ahead of the device before step 0, as it does in every later step); CPU reference arm: host clock per step"}


def ncu_traffic(workload, kernel=None):
    """DRAM bytes per launch from the committed ncu --set full capture (profiles/traffic.json), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            table = json.load(f).get(workload, {})
        return table if kernel is None else table.get(kernel)
    except Exception:
        return None if kernel is not None else {}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _once(self):
        nv = self.nv
        self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
        r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
            else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
                 0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting"}
        for bit, name in names.items():
            if r & bit:
                self.reasons.add(name)

    def run(self):
        if self.nv is None:
            return
        while not self._stop_evt.is_set():
            try:
                self._once()
            except Exception:
                return
            time.sleep(0.002)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=1.0)
        if self.nv is None or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml unavailable"]}
        return {"sm_mhz": float(statistics.median(self.samples)), "sm_max_mhz": float(self.max_mhz),
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------------
# CPU arm: the reference itself (baseline/_ref) when it was staged, else the oracle port
# ------------------------------------------------------------------------------------------------------
def cpu_reference_forward(M, N):
    """Returns (forward(x, tbes, tens), kind, description) of the CPU implementation of the path."""
    from baseline import ref_runner
    if ref_runner.available():
        layer = ref_runner.reference_layer(N, M, bin_len_for(M), dnum=1)
        return layer.forward, "reference", ("the reference's own models/tflct.py::lct.forward (tflct.py:94-179, staged unmodified under "
                                            "baseline/_ref; torch.rfft/ifft shim + crop pin), torch CPU fp32")
    from oracle.lct_oracle import LctOracle
    orc = LctOracle(N, M, bin_len_for(M))
    return orc.forward, "port", "oracle port of tflct.py:94-179 (oracle/lct_oracle.py), torch CPU fp32 -- baseline/_ref was not staged"


def cpu_reference_times(M, N, B, steps, warm, threads=None, arm=None):
    """Host-clock seconds of `steps` forwards of a whole batch of B transients on the CPU arm."""
    if threads:
        torch.set_num_threads(threads)
    fwd, kind, what = arm if arm is not None else cpu_reference_forward(M, N)
    torch.manual_seed(410)
    x = torch.rand(B, 1, M, N, N)
    tbes, tens = [0] * B, [M] * B
    times = []
    with torch.no_grad():
        for i in range(warm + steps):
            t0 = time.perf_counter()
            fwd(x, tbes, tens)
            if i >= warm:
                times.append(time.perf_counter() - t0)
    return times, kind, what


def library_port_rate(M, N, B, dev, reps=10, warm=3):
    """ms per forward of the oracle port run on the GPU through torch (cuFFT + dense cuBLAS matmuls, the
    strongest pre-existing implementation of the reference's op sequence) -- a second baseline
    reported beside the CPU one, never the product path."""
    from oracle.lct_oracle import LctOracle
    orc = LctOracle(N, M, bin_len_for(M))
    torch.manual_seed(410)
    x = torch.rand(B, 1, M, N, N, device=dev)
    evs = []
    with torch.no_grad():
        for i in range(warm + reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            orc.forward(x, [0] * B, [M] * B)
            b.record()
            if i >= warm:
                evs.append((a, b))
    torch.cuda.synchronize()
    return statistics.median(a.elapsed_time(b) for a, b in evs)


def run_reference(args, rank, out):
    """`--impl reference`: the reference's own CPU implementation of the path on the host cores, the whole batch of
    the workload per step, all host threads; rank 0 only (the other ranks exit without work)."""
    if rank != 0:
        return
    name = "cfg4" if args.workload == "cfg4_train" else args.workload
    B, M, N, desc = WORKLOADS[name]
    cores = os.cpu_count() or 1
    times, kind, what = cpu_reference_times(M, N, B, args.steps, args.warmup, threads=cores)
    total = sum(times)
    value = B * len(times) / total
    line = {
        "impl": "reference", "metric": "lct_transients_per_sec_fwd", "value": value, "unit": "transients/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": make_config(desc, B, M, N),
        "cpu_baseline": {"value": value, "unit": "transients/s", "cores": torch.get_num_threads(), "kind": kind,
                         "sample": f"the whole batch {B}x1x{M}x{N}x{N} of the workload per step, forward under no_grad, {what}, "
                                   f"{torch.get_num_threads()} threads; host has {cores} logical cores"},
        "e2e": {"value": value, "unit": "transients/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), file=out, flush=True)


def claim_stdout():
    """Route everything libraries print on fd 1 (NCCL's version banner, ...) to stderr and return a
    handle on the real stdout: the contract is ONE JSON line there."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


# ------------------------------------------------------------------------------------------------------
# GPU arm helpers
# ------------------------------------------------------------------------------------------------------
class Ctx:
    """What every measurement needs: device, ranks, the L2 flush buffer, barrier and event helpers."""

    def __init__(self, dev, rank, world):
        import torch.distributed as dist
        self.dev, self.rank, self.world, self.dist = dev, rank, world, dist
        self.flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    @staticmethod
    def events(n):
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(n)]
        for e in evs:
            e.record()              # instantiates the underlying cudaEvent_t
        return evs

    def max_over_ranks(self, v):
        from hiddenpose_b200 import sharding
        return sharding.max_over_ranks(v, self.dev)

    def time_steps(self, fn, steps, warm, per_step=False):
        """W untimed calls, then K calls each bracketed by a CUDA event pair on the current stream with an L2 flush
        before it (outside the events), a barrier + synchronize on both sides.  Returns total ms, max over ranks
        (and this rank's per-step list when asked)."""
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        evs = [self.events(2) for _ in range(steps)]
        self.barrier()
        for _ in range(8):
            self.flush.zero_()      # pre-roll (see the headline loop): the host leads the device from the first timed step on
        for a, b in evs:
            self.flush.zero_()
            a.record()
            fn()
            b.record()
        self.barrier()
        ms = [a.elapsed_time(b) for a, b in evs]
        total = self.max_over_ranks(sum(ms))
        return (total, ms) if per_step else total


def strong_cfg3(ctx, steps, warm, peak):
    """BASELINE.json configs[2]: one batch of 64 x 512x128x128 transients, sharded contiguously over the ranks
    (tflct.py:121 flattens B*D and nothing mixes channels, which is what makes the split legal).  Every rank seeds
    the same batch, keeps its shard, and runs forward and forward+backward; time = max over ranks."""
    import hiddenpose_b200 as hp
    from hiddenpose_b200 import sharding
    total, M, N = 64, 512, 128
    lo, hi = sharding.shard_bounds(total, ctx.world, ctx.rank)
    C = hi - lo
    gen = torch.Generator(device=ctx.dev).manual_seed(410)
    full = torch.rand(total, 1, M, N, N, device=ctx.dev, generator=gen)            # 2 GiB, identical on every rank
    x = sharding.shard_batch(full, ctx.world, ctx.rank).clone()
    del full
    g = torch.randn(C, 1, M, N, N, device=ctx.dev, generator=gen)
    layer = hp.lct(spatial=N, crop=M, bin_len=bin_len_for(M))
    layer.todev(ctx.dev, 1)
    tbes, tens = [0] * C, [M] * C
    with torch.no_grad():
        fwd_ms = ctx.time_steps(lambda: layer(x, tbes, tens), steps, warm)
    xg = x.requires_grad_(True)

    def fwd_bwd():
        layer(xg, tbes, tens).backward(g)
        xg.grad = None
    fb_ms = ctx.time_steps(fwd_bwd, steps, warm)
    A = chain_bytes(M, N, C)                              # per GPU (this rank's shard; equal shards at 1/2/4/8)
    V = M * N * N
    rec = {
        "what": "BASELINE.json configs[2]: ONE batch of 64 x 1 x 512x128x128 transients sharded over the ranks "
                "(sharding.shard_batch), forward; total transients/s, time = max over ranks",
        "total_transients": total, "per_gpu": C, "steps": steps, "warmup": warm,
        "value": total * steps / (fwd_ms * 1e-3), "unit": "transients/s", "ms_per_step": fwd_ms / steps,
        "chain_bytes_per_gpu": A, "chain_gbs_per_gpu": A / (fwd_ms / steps * 1e-3) / 1e9,
        "chain_frac": A / (fwd_ms / steps * 1e-3) / 1e9 / peak,
        "chain_frac_of_8tbs": A / (fwd_ms / steps * 1e-3) / 1e9 / NOMINAL_HBM_GBS,
        "filter_share_of_A": 32 * V / A,
        "fwd_bwd": {"value": total * steps / (fb_ms * 1e-3), "unit": "transients/s", "ms_per_step": fb_ms / steps,
                    "chain_frac": 2 * A / (fb_ms / steps * 1e-3) / 1e9 / peak},
        "scaling": "strong",
        "note": "per-GPU launches shrink with N (64/32/16/8 channels) while the filter (32 V bytes per launch) does not: "
                "its share of A grows from 0.5 % at C = 64 to 3.7 % at C = 8 -- the only first-order deviation from linear",
    }
    del x, xg, g, layer
    torch.cuda.empty_cache()
    return rec


class _PoseNetStandIn(torch.nn.Module):
    """Stands in for everything downstream of the LCT in NlosPose (UNet3d + posenet3d_50 + heads: out of scope,
    SURVEY.md section 2): the same number of trainable fp32 weights, split into 25 MB tensors so that DDP buckets
    and all-reduces their gradients exactly as it would the real ones.  Its arithmetic is a mean over the volume;
    its backward hands every weight tensor a gradient, in reverse registration order, before the LCT's backward
    starts -- the order in which the real network's gradients become ready."""

    def __init__(self, n_params):
        super().__init__()
        chunk = 25 * (1 << 20) // 4
        sizes = [chunk] * (n_params // chunk) + ([n_params % chunk] if n_params % chunk else [])
        self.blocks = torch.nn.ParameterList([torch.nn.Parameter(torch.zeros(s)) for s in sizes])

    def forward(self, volume):
        touch = sum(p[0] for p in self.blocks)           # every block takes part in the graph
        return volume.mean() + 0.0 * touch


class _TrainPath(torch.nn.Module):
    """feature_extraction -> feature_propagation -> normalize_feature (NlosPose.py:51-54) -> stand-in for the rest."""

    def __init__(self, M, N, dev):
        super().__init__()
        import hiddenpose_b200 as hp
        self.M = M
        self.feature_extraction = hp.FeatureExtraction(basedim=1, in_channels=1, stride=1)
        self.feature_propagation = hp.FeaturePropagation(image_size=N, time_size=M, bin_len=bin_len_for(M), wall_size=2.0,
                                                         mode="lct", material="diffuse", dnum=1, dev=dev)
        self.rest = _PoseNetStandIn(POSE_NET_PARAMETERS - sum(p.numel() for p in self.feature_extraction.parameters()))

    def forward(self, x):
        import hiddenpose_b200 as hp
        f = self.feature_extraction(x)
        v = self.feature_propagation(f, [0, 0, 0], [self.M] * 3)
        return self.rest(hp.normalize_feature(v))


def train_cfg4(ctx, steps, warm):
    """BASELINE.json configs[3]: the training step around the LCT at 16 x 1 x 128^3 per GPU, DDP over NCCL."""
    import contextlib
    import hiddenpose_b200 as hp
    from torch.nn.parallel import DistributedDataParallel as DDP
    B, M, N = 16, 128, 128
    dev = ctx.dev
    torch.manual_seed(410)
    model = _TrainPath(M, N, dev).to(dev)
    n_params = sum(p.numel() for p in model.parameters())
    ddp = DDP(model, device_ids=[dev.index], gradient_as_bucket_view=True) if ctx.world > 1 else None
    net = ddp if ddp is not None else model
    opt = torch.optim.SGD(model.parameters(), lr=1e-4)
    gen = torch.Generator(device=dev).manual_seed(410 + ctx.rank)
    x = torch.rand(B, 1, M, N, N, device=dev, generator=gen)

    def step(sync):
        def run():
            opt.zero_grad(set_to_none=True)
            with (contextlib.nullcontext() if (sync or ddp is None) else ddp.no_sync()):
                net(x).backward()
            opt.step()
        return run

    rec = {"what": "training step around the LCT (utils/train_epoch.py:37-38,74-76 at the train.py:77-86 shape): "
                   "FeatureExtraction(stride 1) -> FeaturePropagation(128, 128) -> normalize_feature -> stand-in for the pose net "
                   "(same parameter count), loss.backward(), SGD step; batch 16 x 1 x 128x128x128 per GPU",
           "per_gpu_batch": B, "parameters": n_params, "gradient_bytes": 4 * n_params, "steps": steps, "warmup": warm}
    local_ms = ctx.time_steps(step(False), steps, warm) / steps
    rec["ms_per_step_no_allreduce"] = local_ms
    if ddp is not None:
        sync_ms = ctx.time_steps(step(True), steps, warm) / steps
        flat = torch.zeros(n_params, device=dev)
        alone_ms = ctx.time_steps(lambda: ctx.dist.all_reduce(flat), steps, warm) / steps
        del flat
        rec.update({
            "ms_per_step_with_allreduce": sync_ms, "allreduce_alone_ms": alone_ms,
            "allreduce_busbw_gbs": 4 * n_params * 2 * (ctx.world - 1) / ctx.world / (alone_ms * 1e-3) / 1e9,
            # what wrapping the step in DDP costs (bucketed all-reduce on NCCL's stream next to the backward kernels, bucket
            # bookkeeping) and how much of the stand-alone all-reduce time stays hidden behind the backward pass
            "ddp_cost_ms": sync_ms - local_ms,
            "overlap": max(0.0, min(1.0, 1.0 - (sync_ms - local_ms) / alone_ms)),
            "value": ctx.world * B / (sync_ms * 1e-3), "unit": "transients/s", "ms_per_step": sync_ms,
        })
    else:
        rec.update({"value": B / (local_ms * 1e-3), "unit": "transients/s", "ms_per_step": local_ms,
                    "note_n1": "one rank: no gradient exchange to time"})
    # the part of the step this library owns: skip branch -> LCT -> normalize_feature, forward + backward
    fe, fp = model.feature_extraction, model.feature_propagation
    feat = torch.randn(B, 1, M, N, N, device=dev, generator=gen).requires_grad_(True)
    xin = x.clone().requires_grad_(True)
    gout = torch.randn(B, 1, M, N, N, device=dev, generator=gen)

    def lct_part():
        v = hp.normalize_feature(fp(hp.skip_sum(feat, xin, fe.weights), [0, 0, 0], [M] * 3))
        v.backward(gout)
        feat.grad = xin.grad = fe.weights.grad = None
    part_ms = ctx.time_steps(lct_part, steps, warm) / steps
    rec["lct_part_ms"] = part_ms
    rec["lct_part_what"] = "skip_sum -> LCT (min/max fused) -> normalize_feature, forward + backward: the library's kernels in this step"
    rec["lct_part_share_of_step"] = part_ms / rec["ms_per_step"]
    A = chain_bytes(M, N, B)
    rec["lct_fwd_bwd_contract_ms_at_peak"] = 2 * A / (measured_peak()[0] * 1e9) * 1e3
    del model, net, ddp, opt, x, feat, xin, gout
    torch.cuda.empty_cache()
    return rec


def bound_label(dram_bytes, ms, peak, issue_bound_name):
    """'hbm' only where the kernel's measured DRAM traffic keeps HBM busy more than 60 % of its duration."""
    if dram_bytes is None:
        return issue_bound_name + " (no ncu traffic figure for this shape: by analogy with the profiled shapes)"
    return "hbm" if dram_bytes / (ms * 1e-3) / 1e9 / peak > 0.6 else issue_bound_name


def main():
    real_stdout = claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong_cfg3 record")
    ap.add_argument("--no-train", action="store_true", help="skip the train_cfg4 record")
    ap.add_argument("--lean", action="store_true", help="headline + per-kernel stages only (for profiler runs)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU work spent on the cpu_baseline sample")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, real_stdout)
        return

    import torch.distributed as dist
    import hiddenpose_b200 as hp

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = Ctx(dev, rank, world)
    barrier, new_events, flush = ctx.barrier, ctx.events, ctx.flush
    peak, peak_src = measured_peak()
    K, W = args.steps, args.warmup
    K_side = max(3, min(K, 10))          # steps for the big side records (strong_cfg3, train_cfg4)

    if args.workload == "cfg4_train":
        B, M, N, desc = WORKLOADS["cfg4_train"]
        sampler = ClockSampler(local_rank)
        sampler.start()
        rec = train_cfg4(ctx, K, W)
        clocks = sampler.stop()
        if rank == 0:
            line = {"metric": "lct_train_step_transients_per_sec", "value": rec["value"], "unit": "transients/s", "n_gpus": world,
                    "steps": K, "warmup": W, "ms_per_step": rec["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                    "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": make_config(desc, B, M, N),
                    "clocks": clocks, "train_cfg4": rec}
            print(json.dumps(line), file=real_stdout, flush=True)
        if world > 1:
            dist.destroy_process_group()
        return

    B, M, N, desc = WORKLOADS[args.workload]
    layer = hp.lct(spatial=N, crop=M, bin_len=bin_len_for(M), wall_size=2.0, method="lct", material="diffuse")
    layer.todev(dev, 1)
    plan = layer._plan
    torch.manual_seed(410 + rank)
    x = torch.rand(B, 1, M, N, N, device=dev)
    tbes, tens = [0] * B, [M] * B
    C, V = B, M * N * N

    # ---- forward, device-resident: the headline.  One event pair per step around the public call ----
    with torch.no_grad():
        for _ in range(W):
            y = layer(x, tbes, tens)
        torch.cuda.synchronize()
        step_events = [new_events(2) for _ in range(K)]
        sampler = ClockSampler(local_rank)
        barrier()
        sampler.start()
        for _ in range(int(os.environ.get("LCT_BENCH_PREROLL", "8"))):
            flush.zero_()          # ~1 ms of queued memsets: the host gets ahead of the device before the first timed step,
            #                        as it is in every later step (an empty queue makes the first step wait on the launches)
        for i in range(K):
            flush.zero_()
            step_events[i][0].record()
            y = layer(x, tbes, tens)
            step_events[i][1].record()
        barrier()
        clocks = sampler.stop()
    step_ms = [a.elapsed_time(b) for a, b in step_events]
    print(f"[rank {rank}] headline step times (us): " + " ".join(f"{1e3 * v:.0f}" for v in step_ms), file=sys.stderr)
    total_ms = ctx.max_over_ranks(sum(step_ms))
    value = world * B * K / (total_ms * 1e-3)

    # ---- same K steps again with one event per kernel (lct_run_staged: kernels back to back on one
    #      stream, so each kernel's own duration is visible) -> per-kernel roofline numbers ------------
    with torch.no_grad():
        stage_events = [new_events(6) for _ in range(K)]
        barrier()
        for i in range(K):
            flush.zero_()
            plan.run_staged(x, tbes, tens, M, stage_events[i], backward=False)
        barrier()
    stage_ms = [[ev[j].elapsed_time(ev[j + 1]) for ev in stage_events] for j in range(5)]
    serial_ms = statistics.fmean(ev[0].elapsed_time(ev[5]) for ev in stage_events)

    lean = args.lean
    # ---- forward + backward (autograd through the module), device-resident -----------------
    xg = x.clone().requires_grad_(True)
    g = torch.randn(B, 1, M, N, N, device=dev)

    def fwd_bwd():
        layer(xg, tbes, tens).backward(g)
        xg.grad = None
    fb_ms = ctx.time_steps(fwd_bwd, K, W)

    side_ms, model_path = {}, None
    latency_us, graph_ok = {"eager": None, "cuda_graph": None}, None
    e2e = {"value": None}
    if not lean:
        # ---- the two ops either side of the layer in the model, device-resident -----
        from hiddenpose_b200.feature_extraction import skip_sum
        from hiddenpose_b200.feature_propagation import normalize_feature
        KN = min(K, 50)
        w27 = torch.randn(1, 1, 3, 3, 3, device=dev)
        feat = torch.randn_like(x)
        with torch.no_grad():
            for name, fn in (("skip_sum", lambda: skip_sum(feat, x, w27)), ("normalize_feature", lambda: normalize_feature(y))):
                for _ in range(3):
                    fn()
                evs = [new_events(2) for _ in range(KN)]
                for a, b in evs:
                    flush.zero_()
                    a.record()
                    fn()
                    b.record()
                torch.cuda.synchronize()
                side_ms[name] = statistics.median(a.elapsed_time(b) for a, b in evs)

        # ---- the path NlosPose takes (NlosPose.py:25-32,51-54): FeatureExtraction's skip branch writes the layer's
        #      input, FeaturePropagation runs the LCT with min/max reduced in its last kernel, normalize_feature
        #      applies the affine map -- one chained number beside the bare layer ------------------------------------
        fprop = hp.FeaturePropagation(image_size=N, time_size=M, bin_len=bin_len_for(M), wall_size=2.0, mode="lct",
                                      material="diffuse", dnum=1, dev=dev)
        with torch.no_grad():
            def nlospose_path():
                return normalize_feature(fprop(skip_sum(feat, x, w27), [0, 0, 0], [M, M, M]))

            def fp_only():
                return fprop(x, [0, 0, 0], [M, M, M])
            path_ms = ctx.time_steps(nlospose_path, KN, 3) / KN
            fp_ms = ctx.time_steps(fp_only, KN, 3) / KN
        groups_now = max(1, min(int(os.environ.get("LCT_STREAM_GROUPS", "2")), 8, C)) if C >= 2 else 1
        model_path = {
            "what": "skip_sum -> FeaturePropagation.forward (LCT, min/max reduced in its last kernel) -> normalize_feature: "
                    "the ops of NlosPose.py:51-54 this library owns, chained, forward, device-resident, L2 flushed per step",
            "ms_per_step": path_ms, "value": world * B / (path_ms * 1e-3), "unit": "transients/s",
            "feature_propagation_only_ms": fp_ms, "bare_layer_ms": total_ms / K,
            "minmax_fusion_cost_ms": fp_ms - total_ms / K,
            "gpu_launches_per_step": 1 + (3 if N <= 64 else 5) * groups_now + 1 + 1,
        }
        del feat, fprop

        # ---- latency of one call as a latency-bound caller sees it: host clock around call + synchronize,
        #      issued kernel by kernel (eager) and as one CUDA-graph replay (hiddenpose_b200.LctGraph) ---------
        try:
            graphed = hp.LctGraph(layer, tuple(x.shape), tbes, tens)
            with torch.no_grad():
                for name, fn in (("eager", lambda: layer(x, tbes, tens)), ("cuda_graph", lambda: graphed(x))):
                    for _ in range(5):
                        fn()
                    torch.cuda.synchronize()
                    ts = []
                    for _ in range(KN):
                        t0 = time.perf_counter()
                        fn()
                        torch.cuda.synchronize()
                        ts.append((time.perf_counter() - t0) * 1e6)
                    latency_us[name] = statistics.median(ts)
            graph_ok = bool(torch.equal(graphed(x), y))
            del graphed
        except Exception as exc:                         # a side measurement: never let it take the bench line down
            print(f"latency section skipped: {exc!r}", file=sys.stderr)
            torch.cuda.synchronize()

        # ---- end to end: pinned host input -> H2D -> forward -> D2H of the volume, every step ---
        # through the public streaming API (hiddenpose_b200.LctStreamer): consecutive steps overlap
        # their upload / transform / download legs; every step still moves its own input and output.
        from hiddenpose_b200.streaming import bind_host_to_gpu, host_topology
        numa_bound = bind_host_to_gpu(local_rank)          # pinned buffers on the GPU's own NUMA node
        n_buf = 4
        x_hosts = [x.cpu().pin_memory() for _ in range(n_buf)]
        y_hosts = [torch.empty(B, 1, M, N, N).pin_memory() for _ in range(n_buf)]
        streamer = hp.LctStreamer(layer, tbes, tens, depth=2)
        streamer.run([x_hosts[i % n_buf] for i in range(W)], [y_hosts[i % n_buf] for i in range(W)])
        e2e_runs = []
        for _ in range(21):                     # host-side jitter (other tenants on the PCIe switch) comes in bursts of several runs: median of 21
            barrier()
            t0 = time.perf_counter()
            streamer.run([x_hosts[i % n_buf] for i in range(K)], [y_hosts[i % n_buf] for i in range(K)])
            torch.cuda.synchronize()
            e2e_runs.append(ctx.max_over_ranks((time.perf_counter() - t0) * 1e3))
        barrier()
        e2e_ms = statistics.median(e2e_runs)
        e2e_value = world * B * K / (e2e_ms * 1e-3)
        e2e_ok = bool(torch.equal(y_hosts[(K - 1) % n_buf], y.cpu()))

        # ---- what the link allows: the same bytes copied both ways at once with no compute in between, on every
        #      rank at the same time (so at N > 1 it is the ceiling the ranks' shared host uplinks leave each GPU) ----
        scratch_y = torch.empty_like(y)
        c_in, c_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        link_runs = []
        for _ in range(9):
            barrier()
            t0 = time.perf_counter()
            for i in range(K):
                with torch.cuda.stream(c_in):
                    x.copy_(x_hosts[i % n_buf], non_blocking=True)
                with torch.cuda.stream(c_out):
                    y_hosts[i % n_buf].copy_(scratch_y, non_blocking=True)
            torch.cuda.synchronize()
            link_runs.append(ctx.max_over_ranks((time.perf_counter() - t0) * 1e3))
        link_ms = statistics.median(link_runs)
        link_gbs = x.numel() * 4 * K / (link_ms * 1e-3) / 1e9          # per direction, per GPU
        del scratch_y, x_hosts, y_hosts, streamer
        e2e = {"value": e2e_value, "unit": "transients/s", "h2d_bytes_per_step": x.numel() * 4,
               "d2h_bytes_per_step": y.numel() * 4, "matches_device_path": e2e_ok,
               "ms_per_step": e2e_ms / K,
               "link_gbs_per_direction_per_gpu": link_gbs, "link_ms_per_step": link_ms / K,
               "frac_of_link": link_ms / e2e_ms,
               "link_what": "the same pinned buffers copied H2D and D2H concurrently with no compute, all ranks at once: the "
                            "host-link ceiling of this step at this N",
               "host_bound_to_gpu_numa_node": numa_bound, "host_topology": host_topology(local_rank),
               "how": "LctStreamer public API: pinned x -> H2D -> lct.forward -> D2H of the whole volume for every "
                      "step; upload/transform/download of consecutive steps overlap on three streams; host wall "
                      "clock from first upload to last byte landed; median of 21 runs of K steps",
               "runs_ms": e2e_runs}

    # ---- the BASELINE-named multi-GPU configurations, on every line (bounded: K_side steps each) ----------
    strong = train = None
    del xg, g
    torch.cuda.empty_cache()
    if not lean and not args.no_strong and args.workload == "cfg2":
        try:
            strong = strong_cfg3(ctx, K_side, 3, peak)
        except Exception as exc:
            strong = {"unavailable": repr(exc)[:300]}
            torch.cuda.synchronize()
    if not lean and not args.no_train and args.workload == "cfg2":
        try:
            train = train_cfg4(ctx, K_side, 3)
        except Exception as exc:
            train = {"unavailable": repr(exc)[:300]}
            torch.cuda.synchronize()

    if rank == 0:
        sb = stage_bytes(M, N, C)
        mean_stage = [statistics.fmean(v) for v in stage_ms]
        stages = [{"kernel": STAGES[j], "ms": mean_stage[j], "bytes": sb[j],
                   "gbs": sb[j] / (mean_stage[j] * 1e-3) / 1e9, "frac": sb[j] / (mean_stage[j] * 1e-3) / 1e9 / peak,
                   "share": mean_stage[j] / sum(mean_stage)} for j in range(5)]
        fused = N <= 64 and mean_stage[2] < 0.25 * mean_stage[1]   # K2+K3+K4 ran as the plane-fused kernel (events 2, 3 are empty)
        if fused:
            mid_ms = mean_stage[1] + mean_stage[2] + mean_stage[3]
            # algorithmic bytes = the contract's figure for the passes this kernel performs (SURVEY 8d:
            # K2 24VC + K3 32VC+32V + K4 24VC); what it has to move itself now that the plane stays in
            # shared memory (one read + one write of S1, filter once) is reported next to it
            mid_bytes = sb[1] + sb[2] + sb[3]
            own_bytes = 16 * V * C + 32 * V
            stages = [stages[0],
                      {"kernel": "plane_filter(H.W.filter.W'.H')", "ms": mid_ms, "bytes": mid_bytes,
                       "gbs": mid_bytes / (mid_ms * 1e-3) / 1e9, "frac": mid_bytes / (mid_ms * 1e-3) / 1e9 / peak,
                       "share": mid_ms / sum(mean_stage),
                       "plane_resident_model": {"bytes": own_bytes, "gbs": own_bytes / (mid_ms * 1e-3) / 1e9,
                                                "frac": own_bytes / (mid_ms * 1e-3) / 1e9 / peak}},
                      stages[4]]
        top = max(range(len(stages)), key=lambda j: stages[j]["ms"])
        traffic = ncu_traffic(args.workload)
        for s in stages:
            s["dram_bytes_ncu"] = traffic.get(s["kernel"])
            s["bound"] = bound_label(s["dram_bytes_ncu"], s["ms"], peak,
                                     "fp32-issue+smem" if s["kernel"].startswith(("plane", "col")) else "issue+smem-latency")
        groups = max(1, min(int(os.environ.get("LCT_STREAM_GROUPS", "2")), 8, C)) if C >= 2 else 1
        n_kernels = (3 if fused else 5) * groups              # kernels of ours launched per forward step
        A = sum(sb)
        step_mean_ms = statistics.fmean(step_ms)
        chain_gbs = A / (step_mean_ms * 1e-3) / 1e9
        dram_step = sum(s["dram_bytes_ncu"] for s in stages) if all(s["dram_bytes_ncu"] for s in stages) else None
        physical = {
            "what": "what the step really moves and computes: DRAM bytes per step from the committed ncu capture "
                    "(profiles/traffic.json; null when this shape has none) and the unpruned 5 n log2 n fp32 operation count, "
                    "both over the headline step time",
            "dram_bytes_per_step": dram_step,
            "dram_gbs": None if dram_step is None else dram_step / (step_mean_ms * 1e-3) / 1e9,
            "dram_frac_of_peak": None if dram_step is None else dram_step / (step_mean_ms * 1e-3) / 1e9 / peak,
            "fp32_flop_per_step": chain_flops(M, N, C),
            "fp32_tflops": chain_flops(M, N, C) / (step_mean_ms * 1e-3) / 1e12,
            "fused_minimum_bytes": (40 if fused else 104) * V * C + 32 * V,
        }
        roofline = {
            "bound": stages[top]["bound"], "bound_of_the_contract_figure": "hbm",
            "kernel": stages[top]["kernel"], "achieved": stages[top]["gbs"], "peak": peak,
            "unit": "GB/s", "frac": stages[top]["frac"],
            "traffic": stages[top]["dram_bytes_ncu"], "peak_source": peak_src,
            "algorithmic_bytes": stages[top]["bytes"],
            # flat scalars (nested records do not survive every parser)
            "chain_bytes": A, "chain_gbs": chain_gbs, "chain_frac": chain_gbs / peak,
            "chain_frac_of_8tbs": chain_gbs / NOMINAL_HBM_GBS,
            "time_fwd_frac": stages[0]["frac"], "time_inv_frac": stages[-1]["frac"],
            "time_fwd_ms": stages[0]["ms"], "time_inv_ms": stages[-1]["ms"], "dominant_ms": stages[top]["ms"],
            "physical_dram_gbs": physical["dram_gbs"], "physical_fp32_tflops": physical["fp32_tflops"],
            "serial_ms_per_step": serial_ms,
            "plane_resident_model": stages[top].get("plane_resident_model"),
            "note": "frac = contract bytes (SURVEY 8d) of the passes the kernel performs / its CUDA-event time / peak; `bound` says what "
                    "limits the kernel by measurement (ncu DRAM traffic vs duration): the fused kernels are not DRAM-bound. "
                    "Per-kernel times from lct_run_staged (single stream); the headline runs two channel groups on two streams",
        }
        if strong and "value" in strong:
            roofline.update({"strong_cfg3_transients_per_s": strong["value"], "strong_cfg3_ms_per_step": strong["ms_per_step"],
                             "strong_cfg3_chain_frac_per_gpu": strong["chain_frac"], "strong_cfg3_per_gpu": strong["per_gpu"],
                             "strong_cfg3_fwd_bwd_transients_per_s": strong["fwd_bwd"]["value"]})
        if train and "ms_per_step" in train:
            roofline.update({"train_cfg4_ms_per_step": train["ms_per_step"],
                             "train_cfg4_ms_no_allreduce": train["ms_per_step_no_allreduce"],
                             "train_cfg4_allreduce_alone_ms": train.get("allreduce_alone_ms"),
                             "train_cfg4_overlap": train.get("overlap"), "train_cfg4_lct_part_ms": train["lct_part_ms"]})
        line = {
            "metric": "lct_transients_per_sec_fwd", "value": value, "unit": "transients/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": total_ms / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": make_config(desc, B, M, N),
            "clocks": clocks,
            "e2e": e2e,
            "gpu_launches": n_kernels * K,
            "fwd_bwd": {"value": world * B * K / (fb_ms * 1e-3), "unit": "transients/s", "ms_per_step": fb_ms / K,
                        "gpu_launches": 2 * n_kernels * K, "chain_frac": 2 * A / (fb_ms / K * 1e-3) / 1e9 / peak},
            "roofline": roofline,
            "physical": physical,
            "stages": stages,
            "model_path": model_path,
            "strong_cfg3": strong,
            "train_cfg4": train,
            "latency_us": {"what": "one forward call of the whole batch + cudaDeviceSynchronize, host wall clock, median; "
                                   "inputs resident, warm L2 (back-to-back calls)",
                           "eager": latency_us["eager"], "cuda_graph": latency_us["cuda_graph"],
                           "graph_matches_eager": graph_ok},
        }
        if side_ms:
            line["neighbours"] = {
                "skip_sum": {"what": "x_conv1 + conv3d(x, w 3x3x3) -- FeatureExtraction's skip branch, writes the layer's input",
                             "ms": side_ms["skip_sum"], "bytes": 12 * V * C,
                             "gbs": 12 * V * C / (side_ms["skip_sum"] * 1e-3) / 1e9},
                "normalize_feature": {"what": "per-channel min/max + affine on the layer's output (min/max not fused here)",
                                      "ms": side_ms["normalize_feature"], "bytes": 12 * V * C,
                                      "gbs": 12 * V * C / (side_ms["normalize_feature"] * 1e-3) / 1e9}}
        if not args.no_cpu_baseline and world == 1 and not lean:      # reported at N = 1 only
            arm = cpu_reference_forward(M, N)
            probe, kind, what = cpu_reference_times(M, N, B, 1, 1, arm=arm)
            reps = max(2, min(40, int(args.cpu_seconds / max(probe[0], 1e-3))))
            times, kind, what = cpu_reference_times(M, N, B, reps, 0, arm=arm)
            line["cpu_baseline"] = {
                "value": B / statistics.median(times), "unit": "transients/s", "cores": torch.get_num_threads(),
                "kind": kind,
                "sample": f"{reps} forwards of the whole batch {B}x1x{M}x{N}x{N} ({what}), median; "
                          f"host has {os.cpu_count()} logical cores"}
            try:
                port_ms = library_port_rate(M, N, B, dev)
                line["cpu_baseline"]["same_ops_on_gpu_via_torch_transients_per_s"] = B / (port_ms * 1e-3)
                line["cpu_baseline"]["same_ops_on_gpu_via_torch_ms_per_step"] = port_ms
                line["cpu_baseline"]["same_ops_on_gpu_via_torch_what"] = (
                    "the reference's op sequence (oracle port) on this B200 through torch: cuFFT fftn/ifftn, dense cuBLAS "
                    "matmuls for the resampling, fp32, whole batch, median of 10 -- the strongest pre-existing implementation")
            except Exception as exc:                     # out of memory at the large shapes: report and move on
                line["cpu_baseline"]["same_ops_on_gpu_via_torch_what"] = "unavailable: " + repr(exc)[:200]
                torch.cuda.synchronize()
        print(json.dumps(line), file=real_stdout, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
