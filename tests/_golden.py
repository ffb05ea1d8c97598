"""Loader for the fixtures minted by tests/golden/make_golden.py (outputs of the
reference itself).  Inputs are regenerated from the recorded seeds."""
import glob
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def case_names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "m*.npz")))


def case_shape(name):
    """(M, N) of a fixture without regenerating its inputs."""
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    return int(z["M"]), int(z["N"])


def make_input(seed, shape):
    return np.random.RandomState(seed).rand(*shape).astype(np.float32)


def make_grad(seed, shape):
    return np.random.RandomState(seed + 1000).randn(*shape).astype(np.float32)


class Case:
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.name = name
        self.z = z
        for k in ("seed", "N", "M", "B", "D", "tbe", "ten"):
            setattr(self, k, int(z[k]))
        self.method, self.material = str(z["method"]), str(z["material"])
        self.bin_len = float(z["bin_len"])
        self.full = bool(z["full"])
        self.tin = self.ten - self.tbe
        self.x = make_input(self.seed, (self.B, self.D, self.tin, self.N, self.N))
        self.g = make_grad(self.seed, (self.B, self.D, self.M, self.N, self.N))
        self.tbes, self.tens = [self.tbe] * self.B, [self.ten] * self.B

    def _err(self, got, key):
        got = np.asarray(got, dtype=np.float64)
        if self.full:
            ref = self.z[key].astype(np.float64)
            assert got.shape == ref.shape
            return np.linalg.norm(got - ref) / np.linalg.norm(ref)
        idx = self.z[key + "_idx"]
        ref = self.z[key + "_sample"].astype(np.float64)
        e = np.linalg.norm(got.ravel()[idx] - ref) / np.linalg.norm(ref)
        n = np.linalg.norm(got) / float(self.z[key + "_norm"])
        return max(e, abs(n - 1.0))

    def y_err(self, y):
        """rel-L2 of ``y`` against the reference output (full volume, or the
        committed sample plus the whole-volume norm for the large cases)."""
        return self._err(y, "y")

    def gx_err(self, gx):
        return self._err(gx, "gx")


def constants():
    return np.load(os.path.join(GOLDEN_DIR, "constants.npz"))


def numpy_lct_fixture():
    """Outputs of /root/reference/utils/lct.py::lct itself, minted by
    tests/golden/make_golden_numpy_lct.py; the (H, W, T) measurement is regenerated from the seed."""
    z = np.load(os.path.join(GOLDEN_DIR, "numpy_lct_m128n32.npz"))
    N, M = int(z["N"]), int(z["M"])
    meas = np.random.RandomState(int(z["seed"])).rand(N, N, M).astype(np.float32)
    return z, meas
