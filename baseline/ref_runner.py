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
