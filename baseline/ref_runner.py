"""Runs the UNMODIFIED reference layer staged under baseline/_ref (see baseline/make_ref.py) on the CPU.

Used only by `bench.py --impl reference` and the `cpu_baseline` leg.  Every arithmetic line executed is the
reference's own (/root/reference/models/tflct.py:94-179 and its constants, :32-79); the two additions below are
outside its arithmetic:

* `torch.rfft` / `torch.ifft` were removed in PyTorch 1.8 (the reference pins 1.7.1); they are provided as thin
  wrappers over `torch.fft.fftn` / `ifftn` with the 1.7 conventions (un-normalised forward, 1/n inverse);
* `tflct.lct.__init__` pins `self.crop = 128` whatever is passed (tflct.py:19); a subclass sets `self.crop = crop`
  before calling the reference's own `parpareparam()`.  At crop == 128 the reference class is used as it is.
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


def available():
    return os.path.isfile(os.path.join(REF_DIR, "models", "tflct.py")) and os.path.isfile(os.path.join(REF_DIR, "utils", "helper.py"))


def _install_shims():
    if hasattr(torch, "rfft"):
        return

    def rfft(x, signal_ndim, normalized=False, onesided=True):
        assert not normalized and not onesided
        return torch.view_as_real(torch.fft.fftn(x, dim=tuple(range(-signal_ndim, 0))))

    def ifft(x, signal_ndim, normalized=False):
        assert not normalized
        return torch.view_as_real(torch.fft.ifftn(torch.view_as_complex(x.contiguous()), dim=tuple(range(-signal_ndim, 0))))

    torch.rfft, torch.ifft = rfft, ifft


def reference_layer(spatial, crop, bin_len, wall_size=2.0, method="lct", material="diffuse", dnum=1):
    """The reference's `lct` module on the CPU, sized for `dnum` channels."""
    if not available():
        raise RuntimeError("baseline/_ref is not staged (run python baseline/make_ref.py in the build container)")
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    _install_shims()
    from models.tflct import lct as ref_lct

    class lct_cropfix(ref_lct):
        def __init__(self, spatial=256, crop=128, **kw):
            self._crop = crop
            super().__init__(spatial=spatial, crop=crop, **kw)

        def parpareparam(self):
            self.crop = self._crop
            super().parpareparam()

    cls = ref_lct if crop == 128 else lct_cropfix
    layer = cls(spatial=spatial, crop=crop, bin_len=bin_len, wall_size=wall_size, method=method, material=material)
    assert layer.crop == crop
    layer.todev("cpu", dnum)
    return layer


# ------------------------------------------------------------------------------------------------------------------
# Downstream of the layer (tests/test_downstream.py only): the reference's own FeatureExtraction, normalize_feature,
# UNet3d, posenet3d_50 and soft-argmax, imported unmodified from baseline/_ref.
# ------------------------------------------------------------------------------------------------------------------
DOWNSTREAM_FILES = ("models/feature_extraction.py", "models/feature_propagation.py", "models/posenet3d_50.py",
                    "unet/unet3d.py", "utils/criterion.py")


def downstream_available():
    return available() and all(os.path.isfile(os.path.join(REF_DIR, f)) for f in DOWNSTREAM_FILES)


def _stub_unreachable_imports():
    """unet/unet3d.py imports its training-script dependencies at module level (torchsummary, the data loader, the yacs
    config); none is used by the network itself.  They are satisfied with empty stand-ins."""
    import types

    class CfgNode(dict):
        __getattr__ = dict.get

        def __setattr__(self, k, v):
            self[k] = v

    def module(name, **attrs):
        if name not in sys.modules:
            m = types.ModuleType(name)
            for k, v in attrs.items():
                setattr(m, k, v)
            sys.modules[name] = m
        return sys.modules[name]

    module("torchsummary", summary=lambda *a, **k: None)
    module("yacs")
    module("yacs.config", CfgNode=CfgNode)
    module("config")
    module("config.config_noise", _C=CfgNode())
    module("utils.nlos_dataloader", NlosDataset=object)


def downstream_modules():
    """(FeatureExtraction, normalize_feature, UNet3d, get_pose_net_50, softmax_integral_tensor) of the reference."""
    if not downstream_available():
        raise RuntimeError("the reference's downstream modules are not staged under baseline/_ref")
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import utils.helper  # noqa: F401  (makes `utils` the staged package before the stand-in submodule is attached)
    _stub_unreachable_imports()
    from models.feature_extraction import FeatureExtraction
    from models.feature_propagation import normalize_feature
    from models.posenet3d_50 import get_pose_net_50
    from unet.unet3d import UNet3d
    from utils.criterion import softmax_integral_tensor
    return FeatureExtraction, normalize_feature, UNet3d, get_pose_net_50, softmax_integral_tensor
