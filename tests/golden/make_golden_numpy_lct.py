#!/usr/bin/env python
"""Mint the SURVEY row a19 fixture by running the reference's NumPy LCT itself (build container only).

    python tests/golden/make_golden_numpy_lct.py            # needs /root/reference

``/root/reference/utils/lct.py::lct`` (lines 9-81) is the reference's second, NumPy statement of
the same transform (lines 41-59 are rows a8-a15 of SURVEY.md section 8), followed by its display
tail: clamp below zero (:62), keep the first 100/128 of the depth axis (:64-65), divide by the
maximum (:66) and show three maximum projections (:68-82).  The function returns nothing and opens
windows, so this script calls it unmodified with two observers, neither of which changes a value:

* ``cv2.imshow`` / ``cv2.waitKey`` are replaced by recorders -- the three images the function
  would have displayed ("front", "left", "top") are its observable outputs;
* ``np.matmul`` is wrapped for the duration of the call to keep a copy of its results: the second
  product (``mtxi @ volume``, lct.py:57-58) is the un-clamped volume.

Both are committed in ``numpy_lct_m128n32.npz`` with the seed of the input.
"""
import os
import sys

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))

N, M, SEED = 32, 128, 4190
WALL, BIN_LEN = 2.0, 0.01 * 512 / 128


def make_measurement():
    """(H, W, T) float32 in [0, 1): what utils/lct.py takes (lct.py:9,17)."""
    return np.random.RandomState(SEED).rand(N, N, M).astype(np.float32)


def main():
    if not os.path.isdir(REF):
        raise SystemExit("reference tree not present; this fixture can only be minted in the build container")
    sys.path.insert(0, REF)
    import cv2
    import utils.lct as ref_np_lct

    shown, products = {}, []
    cv2.imshow = lambda name, img: shown.__setitem__(name, np.array(img, dtype=np.float32))
    cv2.waitKey = lambda *a: 0
    real_matmul = np.matmul

    def recording_matmul(a, b, *args, **kw):
        out = real_matmul(a, b, *args, **kw)
        products.append(np.array(out, copy=True))
        return out

    np.matmul = recording_matmul
    try:
        ref_np_lct.lct(make_measurement(), wall_size=WALL, crop=M, bin_len=BIN_LEN)
    finally:
        np.matmul = real_matmul
    assert set(shown) == {"front", "left", "top"} and len(products) == 2
    volume = products[1].reshape(M, N, N).astype(np.float32)          # lct.py:57-59, before the clamp
    np.savez_compressed(os.path.join(HERE, f"numpy_lct_m{M}n{N}.npz"),
                        seed=np.int64(SEED), N=np.int64(N), M=np.int64(M), wall_size=np.float64(WALL),
                        bin_len=np.float64(BIN_LEN), volume=volume,
                        front=shown["front"], left=shown["left"], top=shown["top"])
    print(f"numpy_lct_m{M}n{N}.npz: |volume| = {np.linalg.norm(volume.astype(np.float64)):.6e}, "
          f"views {shown['front'].shape} {shown['left'].shape} {shown['top'].shape}")


if __name__ == "__main__":
    main()
