"""Drop-in for /root/reference/models/tflct.py::lct (same constructor, ``todev`` and
``forward(feat, tbes, tens)``), running on the sm_100a CUDA library."""
from .layer import LctLayerBase


class lct(LctLayerBase):
    """``lct(spatial, crop, bin_len, wall_size, method, material)`` -- tflct.py:13-15.

    Unlike the reference, which pins ``self.crop = 128`` whatever is passed
    (tflct.py:19), ``crop`` is honoured (as feature_propagation.LCT does); at
    ``crop == 128`` the two agree.
    """

    def __init__(self, spatial=256, crop=128, bin_len=0.01, wall_size=2.0, method="lct", material="diffuse"):
        super().__init__()
        self.spatial_grid = spatial
        self.crop = crop
        self.bin_len = bin_len
        self.wall_size = wall_size
        self.method = method
        self.material = material
        self.parpareparam()

    def parpareparam(self):
        self._setup(self.spatial_grid, self.crop, self.bin_len, self.wall_size, self.method, self.material)
        self.todev("cpu", 2)                                  # tflct.py:79
