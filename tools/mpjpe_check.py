#!/usr/bin/env python
"""Downstream parity: do the CUDA LCT volume and the reference LCT volume give the same joints?

BASELINE.json asks that joint predictions agree within 0.5 mm MPJPE.  The reference's
conv stack (normalize_feature -> UNet3d -> posenet3d_50 -> soft-argmax, NlosPose.py:54-59,
criterion.py:100-154) cannot travel to the GPU box, and the CUDA layer cannot run in the
build container, so the check is split:

  on the GPU box:   python tools/mpjpe_check.py dump      -> gpurun_out/mpjpe_y_cuda.npy
  in the container: python tools/mpjpe_check.py compare   (needs /root/reference)

`compare` regenerates the same seeded input, computes the LCT volume with the CPU oracle
(= the reference's arithmetic), pushes BOTH volumes through the reference's own downstream
modules with identical seeded random weights (released weights are not reachable) and reports
the per-joint distance between the two predictions in millimetres.
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
M, N, B = 128, 128, 1                      # reference training shape (train.py:77-86)
BIN_LEN, WALL = 0.04, 2.0
OUT = os.path.join(ROOT, "gpurun_out", "mpjpe_y_cuda.npy")


def seeded_input():
    """A feature_extraction-like input: smooth non-negative blobs + noise, max-normalised."""
    rs = np.random.RandomState(2024)
    x = rs.rand(B, 1, M, N, N).astype(np.float32) * 0.05
    t, h, w = np.meshgrid(np.arange(M), np.arange(N), np.arange(N), indexing="ij")
    for _ in range(6):
        c = rs.rand(3) * [M * 0.5, N, N] + [M * 0.25, 0, 0]
        s = 4 + 10 * rs.rand()
        x[0, 0] += np.exp(-(((t - c[0]) / s) ** 2 + ((h - c[1]) / (2 * s)) ** 2 + ((w - c[2]) / (2 * s)) ** 2)).astype(np.float32)
    return torch.from_numpy(x / x.max())


def dump():
    import hiddenpose_b200 as hp
    fp = hp.FeaturePropagation(time_size=M, image_size=N, wall_size=WALL, bin_len=BIN_LEN, dnum=1, dev=0)
    y = fp(seeded_input().cuda(), [0, 0, 0], [M, M, M])
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    np.save(OUT, y.cpu().numpy())
    print("wrote", OUT, tuple(y.shape))


def stub_modules():
    class CfgNode(dict):
        def __getattr__(self, k):
            try:
                return self[k]
            except KeyError:
                raise AttributeError(k)

        def __setattr__(self, k, v):
            self[k] = v

        def defrost(self): pass
        def freeze(self): pass
        def clone(self): return self
    yacs = types.ModuleType("yacs"); yc = types.ModuleType("yacs.config"); yc.CfgNode = CfgNode
    yacs.config = yc
    sys.modules.update({"yacs": yacs, "yacs.config": yc})
    ts = types.ModuleType("torchsummary"); ts.summary = lambda *a, **k: None
    sys.modules["torchsummary"] = ts
    for name in ("matplotlib", "matplotlib.pyplot", "mpl_toolkits", "mpl_toolkits.mplot3d", "wandb", "plotly",
                 "plotly.graph_objects", "mat73", "timm", "timm.models", "timm.models.layers"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)


def soft_argmax(heat, num_joints):
    """criterion.py:100-154 (`softmax_integral_tensor`), restated for the CPU: softmax over the
    whole heat-map volume of each joint, then the expectation of each coordinate."""
    b = heat.shape[0]
    d, h, w = heat.shape[-3:]
    p = torch.softmax(heat.reshape(b, num_joints, -1), 2).reshape(b, num_joints, d, h, w)
    x = (p.sum((2, 3)) * torch.arange(w)).sum(2)
    y = (p.sum((2, 4)) * torch.arange(h)).sum(2)
    z = (p.sum((3, 4)) * torch.arange(d)).sum(2)
    return torch.stack([x, y, z], 2)                  # (b, joints, 3) in heat-map voxels


def compare():
    ref = "/root/reference"
    assert os.path.isdir(ref), "compare needs the reference tree"
    stub_modules()
    sys.path.insert(0, ref)
    from models.feature_propagation import normalize_feature
    from unet.unet3d import UNet3d
    from models.posenet3d_50 import get_pose_net_50
    from oracle.lct_oracle import LctOracle, rel_l2

    torch.set_num_threads(os.cpu_count() or 8)
    x = seeded_input()
    y_ref = LctOracle(N, M, BIN_LEN, WALL).forward(x, [0] * B, [M] * B)
    y_cuda = torch.from_numpy(np.load(OUT))
    print(f"volume rel-L2 (CUDA vs reference arithmetic): {rel_l2(y_cuda, y_ref):.3e}")

    torch.manual_seed(410)                                # train.py:98
    unet, pose = UNet3d(in_channels=1, n_channels=4).eval(), get_pose_net_50().eval()

    def joints(vol):
        with torch.no_grad():
            f = normalize_feature(vol.clone())            # NlosPose.py:54
            heat = pose(f + unet(f))                      # NlosPose.py:55-57
        return soft_argmax(heat, 24), heat.shape

    ja, shape = joints(y_ref)
    jb, _ = joints(y_cuda)
    mm_per_voxel = WALL * 1000.0 / shape[-1]              # wall_size / heat-map width
    dist = (ja - jb).norm(dim=2) * mm_per_voxel
    print(f"heat-map {tuple(shape)}, {mm_per_voxel:.2f} mm per voxel")
    print(f"joint distance between the two pipelines: mean {dist.mean():.5f} mm, max {dist.max():.5f} mm (bar: 0.5 mm MPJPE)")
    assert float(dist.mean()) <= 0.5


if __name__ == "__main__":
    {"dump": dump, "compare": compare}[sys.argv[1]]()
