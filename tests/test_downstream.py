"""Downstream parity (BASELINE.json north_star: joint predictions within 0.5 mm MPJPE).

BASELINE.json configs[1]: NlosPose's forward -- feature_extraction -> LCT -> normalize_feature -> UNet3d ->
posenet3d_50 -> soft-argmax (models/NlosPose.py:49-59, utils/criterion.py:100-154) -- at batch 8 x 256x64x64.
The pipeline is run twice on the same seeded input with identical seeded weights (the released checkpoint is an
unreachable download):

  reference arm : the reference's own modules, imported unmodified from baseline/_ref (staged by
                  baseline/make_ref.py), with the LCT computed by the CPU oracle (= tflct.py:94-179 op for op);
  this library  : the drop-in FeatureExtraction (CUDA skip branch), FeaturePropagation (CUDA LCT, min/max fused)
                  and normalize_feature (CUDA), followed by the same UNet3d / posenet3d_50 / soft-argmax.

The two joint sets must agree to 0.5 mm on average.  Skipped where baseline/_ref was not staged.
"""
import numpy as np
import pytest
import torch

from baseline import ref_runner
from oracle import lct_oracle as O

pytestmark = pytest.mark.gpu

B, M, N = 8, 256, 64                 # BASELINE.json configs[1]
BIN_LEN, WALL, JOINTS = 0.01 * 512 / M, 2.0, 24


def _measurement():
    """Max-normalised non-negative transients (what the loaders produce, nlos_pose_dataloader.py:85-87): a few smooth
    returns per sample over a noise floor."""
    rs = np.random.RandomState(2024)
    x = rs.rand(B, 1, M, N, N).astype(np.float32) * 0.05
    t, h, w = np.meshgrid(np.arange(M), np.arange(N), np.arange(N), indexing="ij")
    for b in range(B):
        for _ in range(4):
            c = rs.rand(3) * [M * 0.5, N, N] + [M * 0.25, 0, 0]
            s = 3 + 6 * rs.rand()
            x[b, 0] += np.exp(-(((t - c[0]) / (2 * s)) ** 2 + ((h - c[1]) / s) ** 2 + ((w - c[2]) / s) ** 2)).astype(np.float32)
        x[b] /= x[b].max()
    return torch.from_numpy(x)


def test_downstream_joints_agree_within_half_a_millimetre(parity_log):
    if not ref_runner.downstream_available():
        pytest.skip("the reference's downstream modules are not staged under baseline/_ref")
    import hiddenpose_b200 as hp
    RefFeatureExtraction, ref_normalize_feature, UNet3d, get_pose_net_50, softmax_integral_tensor = ref_runner.downstream_modules()
    dev = torch.device("cuda", 0)
    torch.manual_seed(410)                                                    # train.py:98
    ref_fe = RefFeatureExtraction(basedim=1, in_channels=1, stride=1).to(dev).eval()      # NlosPose.py:19-23
    unet = UNet3d(in_channels=1, n_channels=4).to(dev).eval()                 # NlosPose.py:37-40
    pose = get_pose_net_50().to(dev).eval()                                   # NlosPose.py:47-48
    our_fe = hp.FeatureExtraction(basedim=1, in_channels=1, stride=1).to(dev).eval()
    our_fe.load_state_dict(ref_fe.state_dict(), strict=True)
    our_fp = hp.FeaturePropagation(time_size=M, image_size=N, wall_size=WALL, bin_len=BIN_LEN, dnum=1, dev=0)   # NlosPose.py:25-32
    meas = _measurement().to(dev)

    def joints(volume):
        heat = pose(volume + unet(volume))                                    # NlosPose.py:55-57
        d, h, w = heat.shape[-3:]
        return softmax_integral_tensor(heat, JOINTS, True, w, h, d).reshape(B, JOINTS, 3), (d, h, w)

    with torch.no_grad():
        # reference arm
        feat_ref = ref_fe(meas)                                               # NlosPose.py:51
        y_ref = O.LctOracle(N, M, BIN_LEN, WALL).forward(feat_ref.cpu(), [0] * B, [M] * B).to(dev)     # NlosPose.py:53
        v_ref = ref_normalize_feature(y_ref.clone())                          # NlosPose.py:54
        j_ref, hm = joints(v_ref)
        # this library
        feat = our_fe(meas)
        y = our_fp(feat, [0, 0, 0], [M, M, M])
        v = hp.normalize_feature(y)
        j_ours, _ = joints(v)
    mm_per_voxel = WALL * 1000.0 / hm[2]                                      # wall_size over the heat-map width
    dist = (j_ours - j_ref).norm(dim=2) * mm_per_voxel
    e_feat, e_vol, e_norm = O.rel_l2(feat.cpu(), feat_ref.cpu()), O.rel_l2(y.cpu(), y_ref.cpu()), O.rel_l2(v.cpu(), v_ref.cpu())
    parity_log(B=B, M=M, N=N, feature_rel_l2=e_feat, volume_rel_l2=e_vol, normalized_rel_l2=e_norm,
               heatmap=f"{hm[0]}x{hm[1]}x{hm[2]}", mm_per_voxel=float(mm_per_voxel),
               mpjpe_mm=float(dist.mean()), max_joint_mm=float(dist.max()), bar_mm=0.5)
    assert e_vol <= 1e-5
    assert float(dist.mean()) <= 0.5
