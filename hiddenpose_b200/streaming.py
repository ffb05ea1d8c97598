"""Host-to-host streaming of transients through the LCT layer.

The reference moves one batch at a time: ``input.to(cfg.DEVICE)`` -> ``model(input)`` ->
``.cpu()`` (/root/reference/utils/train_epoch.py:37-38, models/test_tflct.py:93-95), each
step waiting for the previous one.  On a B200 the layer itself takes a fraction of the PCIe
time, so this helper overlaps the three legs of consecutive batches: while batch i is being
transformed, batch i+1 is uploading and batch i-1 is downloading (copy engines run in both
directions at once).  Results are identical to calling the layer batch by batch.
"""
from __future__ import annotations

import torch


def bind_host_to_gpu(device_index: int) -> bool:
    """Pins the calling process to the CPU cores NVML reports as local to the GPU, so that the pinned
    host buffers allocated afterwards (first touch) sit on the GPU's own NUMA node and its PCIe root
    port.  Returns False when NVML or the affinity call is unavailable (nothing is changed then)."""
    try:
        import math
        import os
        import pynvml
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        words = math.ceil((os.cpu_count() or 1) / 64)
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return False
        os.sched_setaffinity(0, cpus)
        return True
    except Exception:
        return False


def host_topology(device_index: int) -> dict:
    """Where this GPU hangs off the host, as far as the box tells: NUMA nodes online, the GPU's PCI address, the NUMA
    node sysfs and NVML report for it, and the CPUs NVML calls local.  Used by bench.py to explain the end-to-end
    (host-link-bound) numbers at N > 1; every field is best effort (None when the box hides it)."""
    info = {"numa_nodes_online": None, "pci_bus_id": None, "pci_numa_node": None, "nvml_cpu_affinity": None,
            "process_cpus": None, "host_cpus": None}
    try:
        import os
        info["host_cpus"] = os.cpu_count()
        info["process_cpus"] = len(os.sched_getaffinity(0))
        with open("/sys/devices/system/node/online") as f:
            info["numa_nodes_online"] = f.read().strip()
    except Exception:
        pass
    try:
        import math
        import os
        import pynvml
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        bus = pynvml.nvmlDeviceGetPciInfo(handle).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        info["pci_bus_id"] = bus
        words = math.ceil((os.cpu_count() or 1) / 64)
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, words)
        cpus = sorted(64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1)
        info["nvml_cpu_affinity"] = f"{cpus[0]}-{cpus[-1]} ({len(cpus)} cpus)" if cpus else None
        short = bus.lower()[-12:]                       # sysfs uses the 4-digit domain form
        with open(f"/sys/bus/pci/devices/{short}/numa_node") as f:
            info["pci_numa_node"] = int(f.read().strip())
    except Exception:
        pass
    return info


class LctStreamer:
    """Pipelines ``(x_host) -> layer -> (y_host)`` over pinned host buffers.

    ``layer`` is an ``lct`` / ``LCT`` / ``FeaturePropagation`` already moved to a CUDA device.
    ``depth`` device-side buffer sets are kept in flight (2 is enough to overlap everything).
    """

    def __init__(self, layer, tbes, tens, depth: int = 2):
        inner = layer.method if isinstance(getattr(layer, "method", None), torch.nn.Module) else layer   # FeaturePropagation wraps an LCT
        if getattr(inner, "_plan", None) is None:
            raise RuntimeError("LctStreamer needs a layer that was moved to a CUDA device with todev()")
        self.layer, self.tbes, self.tens = layer, list(tbes), list(tens)
        self.device = inner._plan.device
        self.depth = max(1, int(depth))
        self.s_in = torch.cuda.Stream(self.device)
        self.s_run = torch.cuda.Stream(self.device)
        self.s_out = torch.cuda.Stream(self.device)
        self._slots = None

    def _alloc(self, x_host):
        self._slots = []
        for _ in range(self.depth):
            self._slots.append({
                "x": torch.empty(x_host.shape, dtype=torch.float32, device=self.device),
                "y": None,
                "uploaded": torch.cuda.Event(), "computed": torch.cuda.Event(), "drained": torch.cuda.Event(),
            })
            self._slots[-1]["drained"].record(self.s_out)

    @torch.no_grad()
    def run(self, x_hosts, y_hosts):
        """Transforms every pinned host batch in ``x_hosts`` into the matching pinned buffer of
        ``y_hosts``.  Returns after the last result has landed on the host."""
        x_hosts, y_hosts = list(x_hosts), list(y_hosts)
        if len(x_hosts) != len(y_hosts):
            raise ValueError("x_hosts and y_hosts must have the same length")
        if not x_hosts:
            return y_hosts
        if self._slots is None or self._slots[0]["x"].shape != x_hosts[0].shape:
            self._alloc(x_hosts[0])
        for i, (xh, yh) in enumerate(zip(x_hosts, y_hosts)):
            slot = self._slots[i % self.depth]
            with torch.cuda.stream(self.s_in):
                self.s_in.wait_event(slot["computed"])          # the slot's previous input was consumed
                slot["x"].copy_(xh, non_blocking=True)
                slot["uploaded"].record(self.s_in)
            with torch.cuda.stream(self.s_run):
                self.s_run.wait_event(slot["uploaded"])
                self.s_run.wait_event(slot["drained"])           # the slot's previous output left the device
                slot["y"] = self.layer(slot["x"], self.tbes, self.tens)
                slot["computed"].record(self.s_run)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(slot["computed"])
                yh.copy_(slot["y"], non_blocking=True)
                slot["y"].record_stream(self.s_out)
                slot["drained"].record(self.s_out)
        self.s_out.synchronize()
        return y_hosts


class LctGraph:
    """One-launch replay of the layer for a fixed shape and window (CUDA graph).

    A single transient keeps a B200 busy for a few tens of microseconds -- less than the host needs to
    issue the layer's kernels one by one -- so a latency-bound caller (one transient per frame, as in
    /root/reference/models/test_tflct.py:84-95) captures the forward once and replays it:
    ``g = LctGraph(layer, (1, 1, T, N, N), tbes, tens); y = g(x)``.  The library neither allocates
    nor synchronises inside ``lct_forward``, which is what makes the capture legal.  ``y`` is a static
    buffer overwritten by the next call; inference only (no autograd through a replay).
    """

    def __init__(self, layer, shape, tbes, tens, normalize: bool = False):
        inner = layer.method if isinstance(getattr(layer, "method", None), torch.nn.Module) else layer
        if getattr(inner, "_plan", None) is None:
            raise RuntimeError("LctGraph needs a layer that was moved to a CUDA device with todev()")
        self.device = inner._plan.device
        self.x = torch.zeros(tuple(shape), dtype=torch.float32, device=self.device)
        tbes, tens = list(tbes), list(tens)

        def body():
            y = layer(self.x, tbes, tens)
            if normalize:
                from .feature_propagation import normalize_feature
                y = normalize_feature(y)
            return y

        side = torch.cuda.Stream(self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.no_grad(), torch.cuda.stream(side):
            body()                                              # warm up outside the capture (workspace, attributes)
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        self.graph = torch.cuda.CUDAGraph()
        # thread_local: CUDA calls made by other threads meanwhile (e.g. a process group's watchdog) must not break the capture
        with torch.no_grad(), torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
            self.y = body()

    def __call__(self, x):
        if x.shape != self.x.shape:
            raise ValueError(f"LctGraph was captured for shape {tuple(self.x.shape)}, got {tuple(x.shape)}")
        self.x.copy_(x, non_blocking=True)
        self.graph.replay()
        return self.y
