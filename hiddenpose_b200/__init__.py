"""hiddenpose_b200 -- the HiddenPose light-cone-transform layer, native to B200 (sm_100a).

Reference-facing classes (same constructors / ``todev`` / ``forward`` as
Hagtaril/HiddenPose): :class:`tflct.lct`, :class:`feature_propagation.LCT`,
:class:`feature_propagation.FeaturePropagation`.  Compute runs only in the
hand-written CUDA library behind ``include/hiddenpose_lct.h``.
"""
from ._native import build_native, load as load_native          # noqa: F401
from .feature_extraction import FeatureExtraction, skip_sum      # noqa: F401
from .feature_propagation import LCT, FeaturePropagation, VisibleNet, normalize, normalize_feature   # noqa: F401
from .lct_function import LctFunction, LctPlan                   # noqa: F401
from .streaming import LctGraph, LctStreamer                               # noqa: F401
from .tflct import lct                                           # noqa: F401

__all__ = ["lct", "LCT", "FeaturePropagation", "VisibleNet", "normalize", "normalize_feature",
           "LctFunction", "LctPlan", "LctStreamer", "build_native", "load_native"]
