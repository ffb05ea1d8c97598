#!/usr/bin/env python
"""Throughput of the LCT hot path on B200 (BASELINE.json metric: LCT transients/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one LCT forward over one batch of synthetic transients.  At N = 1 the
workload is BASELINE.json configs[1]'s LCT part: batch 8 x 1 x 256 x 64 x 64 per GPU
(`--workload cfg3|cfg4|cfg5` selects the other shapes).  N > 1 shards by transient:
every rank runs the same per-GPU batch (weak scaling), no data-path collective.

Prints ONE JSON line (rank 0).  `value` is device-resident throughput; `e2e` is the
same metric through the public module API with pinned-host input and host output
inside the timed region; `roofline` is the dominant kernel's algorithmic bytes over
its CUDA-event duration against MEASURED_PEAKS.json; `cpu_baseline` is the CPU
oracle port (oracle/lct_oracle.py, the reference's op sequence) on the host cores.
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

import numpy as np   # noqa: E402
import torch         # noqa: E402

WORKLOADS = {
    # name: (per-GPU batch, M, N, description)
    "cfg1": (1, 256, 64, "LCT forward, one 1 x 1 x 256x64x64 transient (BASELINE.json configs[0], the reference's CPU-runnable case)"),
    "cfg2": (8, 256, 64, "LCT forward, batch 8 x 1 x 256x64x64 per GPU (BASELINE.json configs[1], LCT part)"),
    "cfg3": (8, 512, 128, "LCT forward, batch 8 x 1 x 512x128x128 per GPU (configs[2] at 8 GPUs)"),
    "cfg4": (16, 128, 128, "LCT forward, batch 16 x 1 x 128x128x128 per GPU (configs[3], LCT part)"),
    "cfg5": (1, 512, 256, "LCT forward, single 512x256x256 transient (configs[4])"),
    "tiny": (2, 64, 16, "LCT forward, batch 2 x 1 x 64x16x16 (harness self-test)"),
}
STAGES = ("time_fwd", "row_fwd", "col_filter", "row_inv", "time_inv")


def bin_len_for(M):
    return 0.01 * 512 / M          # trange = 5.12 as in the released configs


def stage_bytes(M, N, C):
    """Algorithmic bytes per launch of each kernel (DESIGN.md 'Kernels')."""
    V = M * N * N
    return [12 * V * C, 24 * V * C, 32 * V * C + 32 * V, 24 * V * C, 12 * V * C]


def ncu_traffic(workload, kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full capture (or None)."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(workload, {}).get(kernel)
    except Exception:
        return None


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


def cpu_oracle_rate(M, N, reps, warm, threads=None):
    """Transients/s of the CPU oracle port on one transient of the workload's shape."""
    from oracle.lct_oracle import LctOracle
    if threads:
        torch.set_num_threads(threads)
    orc = LctOracle(N, M, bin_len_for(M))
    torch.manual_seed(410)
    x = torch.rand(1, 1, M, N, N)
    times = []
    with torch.no_grad():
        for i in range(warm + reps):
            t0 = time.perf_counter()
            orc.forward(x, [0], [M])
            if i >= warm:
                times.append(time.perf_counter() - t0)
    return times


def library_port_rate(M, N, B, dev, reps=10, warm=3):
    """ms per forward of the same oracle port run on the GPU through torch (cuFFT + dense cuBLAS matmuls, the
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
    """`--impl reference`: the reference's own CPU implementation of the path (the oracle port:
    /root/reference cannot travel to the GPU box) on the host cores; rank 0 only."""
    if rank != 0:
        return
    B, M, N, desc = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    times = cpu_oracle_rate(M, N, args.steps, args.warmup, threads=cores)
    total = sum(times)
    value = len(times) / total
    line = {
        "impl": "reference", "metric": "lct_transients_per_sec_fwd", "value": value, "unit": "transients/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "B": B, "M": M, "N": N},
        "cpu_baseline": {"value": value, "unit": "transients/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"one 1x1x{M}x{N}x{N} transient of the workload per step, forward, torch CPU fp32, "
                                   f"{torch.get_num_threads()} threads"},
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


def main():
    real_stdout = claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-reps", type=int, default=60)
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
    from hiddenpose_b200 import sharding

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    B, M, N, desc = WORKLOADS[args.workload]
    K, W = args.steps, args.warmup
    layer = hp.lct(spatial=N, crop=M, bin_len=bin_len_for(M), wall_size=2.0, method="lct", material="diffuse")
    layer.todev(dev, 1)
    plan = layer._plan
    torch.manual_seed(410 + rank)
    x = torch.rand(B, 1, M, N, N, device=dev)
    tbes, tens = [0] * B, [M] * B
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    def new_events(n):
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(n)]
        for e in evs:
            e.record()              # instantiates the underlying cudaEvent_t
        return evs

    # ---- forward, device-resident: the headline.  One event pair per step around the public call ----
    with torch.no_grad():
        for _ in range(W):
            y = layer(x, tbes, tens)
        torch.cuda.synchronize()
        step_events = [new_events(2) for _ in range(K)]
        sampler = ClockSampler(local_rank)
        barrier()
        sampler.start()
        for i in range(K):
            flush.zero_()
            step_events[i][0].record()
            y = layer(x, tbes, tens)
            step_events[i][1].record()
        barrier()
        clocks = sampler.stop()
    step_ms = [a.elapsed_time(b) for a, b in step_events]
    total_ms = sharding.max_over_ranks(sum(step_ms), dev)
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

    # ---- forward + backward (autograd through the module), device-resident -----------------
    xg = x.clone().requires_grad_(True)
    g = torch.randn(B, 1, M, N, N, device=dev)
    for _ in range(W):
        layer(xg, tbes, tens).backward(g)
        xg.grad = None
    fb_events = [new_events(2) for _ in range(K)]
    barrier()
    for i in range(K):
        flush.zero_()
        fb_events[i][0].record()
        layer(xg, tbes, tens).backward(g)
        fb_events[i][1].record()
        xg.grad = None
    barrier()
    fb_ms = sharding.max_over_ranks(sum(a.elapsed_time(b) for a, b in fb_events), dev)

    # ---- the two ops either side of the layer in the model, device-resident -----
    from hiddenpose_b200.feature_extraction import skip_sum
    from hiddenpose_b200.feature_propagation import normalize_feature
    KN = min(K, 50)
    w27 = torch.randn(1, 1, 3, 3, 3, device=dev)
    feat = torch.randn_like(x)
    side_ms = {}
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
    del feat

    # ---- latency of one call as a latency-bound caller sees it: host clock around call + synchronize,
    #      issued kernel by kernel (eager) and as one CUDA-graph replay (hiddenpose_b200.LctGraph) ---------
    latency_us, graph_ok = {"eager": None, "cuda_graph": None}, None
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
    from hiddenpose_b200.streaming import bind_host_to_gpu
    numa_bound = bind_host_to_gpu(local_rank)          # pinned buffers on the GPU's own NUMA node
    n_buf = 4
    x_hosts = [x.cpu().pin_memory() for _ in range(n_buf)]
    y_hosts = [torch.empty(B, 1, M, N, N).pin_memory() for _ in range(n_buf)]
    streamer = hp.LctStreamer(layer, tbes, tens, depth=2)
    streamer.run([x_hosts[i % n_buf] for i in range(W)], [y_hosts[i % n_buf] for i in range(W)])
    e2e_runs = []
    for _ in range(11):                     # host-side jitter (other tenants on the PCIe switch) comes in bursts of several runs: median of 11
        barrier()
        t0 = time.perf_counter()
        streamer.run([x_hosts[i % n_buf] for i in range(K)], [y_hosts[i % n_buf] for i in range(K)])
        torch.cuda.synchronize()
        e2e_runs.append(sharding.max_over_ranks((time.perf_counter() - t0) * 1e3, dev))
    barrier()
    e2e_ms = statistics.median(e2e_runs)
    e2e_value = world * B * K / (e2e_ms * 1e-3)
    e2e_ok = bool(torch.equal(y_hosts[(K - 1) % n_buf], y.cpu()))

    # ---- what the link allows: the same bytes copied both ways at once with no compute in between ----
    scratch_y = torch.empty_like(y)
    c_in, c_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    link_runs = []
    for _ in range(3):
        barrier()
        t0 = time.perf_counter()
        for i in range(K):
            with torch.cuda.stream(c_in):
                x.copy_(x_hosts[i % n_buf], non_blocking=True)
            with torch.cuda.stream(c_out):
                y_hosts[i % n_buf].copy_(scratch_y, non_blocking=True)
        torch.cuda.synchronize()
        link_runs.append(sharding.max_over_ranks((time.perf_counter() - t0) * 1e3, dev))
    link_ms = statistics.median(link_runs)
    link_gbs = x.numel() * 4 * K / (link_ms * 1e-3) / 1e9          # per direction, per GPU
    del scratch_y

    if rank == 0:
        peak, peak_src = measured_peak()
        C = B
        sb = stage_bytes(M, N, C)
        mean_stage = [statistics.fmean(v) for v in stage_ms]
        stages = [{"kernel": STAGES[j], "ms": mean_stage[j], "bytes": sb[j],
                   "gbs": sb[j] / (mean_stage[j] * 1e-3) / 1e9, "frac": sb[j] / (mean_stage[j] * 1e-3) / 1e9 / peak,
                   "share": mean_stage[j] / sum(mean_stage)} for j in range(5)]
        top = max(range(5), key=lambda j: mean_stage[j])
        fused = N <= 64 and mean_stage[2] < 0.25 * mean_stage[1]   # K2+K3+K4 ran as the plane-fused kernel (events 2, 3 are empty)
        if fused:
            V = M * N * N
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
        groups = max(1, min(int(os.environ.get("LCT_STREAM_GROUPS", "2")), 8, C)) if C >= 2 else 1
        n_kernels = (3 if fused else 5) * groups              # kernels of ours launched per forward step
        chain_bytes = sum(sb)
        chain_gbs = chain_bytes / (statistics.fmean(step_ms) * 1e-3) / 1e9
        line = {
            "metric": "lct_transients_per_sec_fwd", "value": value, "unit": "transients/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": total_ms / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "B_per_gpu": B, "M": M, "N": N, "seed": 410,
                       "l2": "flushed between timed steps (512 MiB write outside the events)",
                       "timing": "CUDA events per step on the launch stream, sum over K steps, max over ranks"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "transients/s", "h2d_bytes_per_step": x.numel() * 4,
                    "d2h_bytes_per_step": y.numel() * 4,
                    "matches_device_path": e2e_ok,
                    "link": {"what": "same host buffers copied H2D and D2H concurrently, no compute: the PCIe ceiling of this step",
                             "gbs_per_direction": link_gbs, "ms_per_step": link_ms / K,
                             "e2e_fraction_of_link": link_ms / e2e_ms},
                    "how": "LctStreamer public API: pinned x -> H2D -> lct.forward -> D2H of the whole volume for every "
                           "step; upload/transform/download of consecutive steps overlap on three streams; host wall "
                           "clock from first upload to last byte landed; median of 11 runs of K steps",
                    "runs_ms": e2e_runs, "host_bound_to_gpu_numa_node": numa_bound},
            "gpu_launches": n_kernels * K,
            "fwd_bwd": {"value": world * B * K / (fb_ms * 1e-3), "unit": "transients/s", "ms_per_step": fb_ms / K,
                        "gpu_launches": 2 * n_kernels * K},
            "roofline": {"bound": "hbm", "kernel": stages[top]["kernel"], "achieved": stages[top]["gbs"], "peak": peak,
                         "unit": "GB/s", "frac": stages[top]["frac"],
                         "traffic": ncu_traffic(args.workload, stages[top]["kernel"]), "peak_source": peak_src,
                         "algorithmic_bytes": stages[top]["bytes"],
                         "plane_resident_model": stages[top].get("plane_resident_model"),
                         "chain": {"bytes": chain_bytes, "gbs": chain_gbs, "frac": chain_gbs / peak,
                                   "note": "A = 104*V*C + 32*V (SURVEY 8d contract figure) over the headline step time"},
                         "serial_ms_per_step": serial_ms,
                         "note": "per-kernel times from lct_run_staged (single stream); the headline runs two "
                                 "channel groups on two streams so consecutive kernels overlap"},
            "stages": stages,
            "latency_us": {"what": "one forward call of the whole batch + cudaDeviceSynchronize, host wall clock, median; "
                                   "inputs resident, warm L2 (back-to-back calls)",
                           "eager": latency_us["eager"], "cuda_graph": latency_us["cuda_graph"],
                           "graph_matches_eager": graph_ok},
            "neighbours": {
                "skip_sum": {"what": "x_conv1 + conv3d(x, w 3x3x3) -- FeatureExtraction's skip branch, writes the layer's input",
                             "ms": side_ms["skip_sum"], "bytes": 12 * M * N * N * C,
                             "gbs": 12 * M * N * N * C / (side_ms["skip_sum"] * 1e-3) / 1e9},
                "normalize_feature": {"what": "per-channel min/max + affine on the layer's output (min/max not fused here)",
                                      "ms": side_ms["normalize_feature"], "bytes": 12 * M * N * N * C,
                                      "gbs": 12 * M * N * N * C / (side_ms["normalize_feature"] * 1e-3) / 1e9}},
        }
        if not args.no_cpu_baseline and world == 1:      # reported at N = 1 only
            times = cpu_oracle_rate(M, N, args.cpu_reps, 2)
            line["cpu_baseline"] = {
                "value": 1.0 / statistics.median(times), "unit": "transients/s", "cores": torch.get_num_threads(),
                "kind": "port",
                "sample": f"{args.cpu_reps} forwards of one 1x1x{M}x{N}x{N} transient (oracle port of tflct.py:94-179, "
                          f"torch CPU fp32), median; host has {os.cpu_count()} logical cores"}
            try:
                port_ms = library_port_rate(M, N, B, dev)
                line["cpu_baseline"]["same_port_on_gpu_via_torch"] = {
                    "value": B / (port_ms * 1e-3), "unit": "transients/s", "ms_per_step": port_ms,
                    "what": "the same oracle port (reference op sequence) on this B200 through torch: cuFFT fftn/ifftn, "
                            "dense cuBLAS matmuls for the resampling, fp32, whole batch, median of 10"}
            except Exception as exc:                     # out of memory at the large shapes: report and move on
                line["cpu_baseline"]["same_port_on_gpu_via_torch"] = {"unavailable": repr(exc)[:200]}
                torch.cuda.synchronize()
        print(json.dumps(line), file=real_stdout, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
