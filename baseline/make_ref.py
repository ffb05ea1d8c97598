#!/usr/bin/env python
"""Stage the reference's own LCT layer for `bench.py --impl reference` (build container only).

    python baseline/make_ref.py            # needs /root/reference

The reference has no setup.py / pyproject.toml, so `pip install --target baseline/_ref /root/reference` is not
applicable (recorded in DESIGN.md section 6).  Its LCT path is two plain Python files with no build step:

    models/tflct.py     (class lct: the layer, :13-179)
    utils/helper.py     (definePsf, resamplingOperator, filterLaplacian)

plus, for the downstream-joints test only, the modules around the layer in NlosPose (see DOWNSTREAM below).

They are copied byte for byte, together with the licence, into baseline/_ref/ -- git-ignored, so no reference source
enters the history, but not gpurun-ignored, so the files travel to the GPU box where /root/reference does not
exist.  `baseline/ref_runner.py` (ours) imports them unmodified.
"""
import hashlib
import os
import shutil
import sys

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
FILES = ("models/tflct.py", "utils/helper.py", "LICENSE")
# The modules either side of the layer in NlosPose (models/NlosPose.py:19-59), staged the same way for the downstream
# parity test (tests/test_downstream.py: joints within 0.5 mm): the reference's own FeatureExtraction,
# normalize_feature, UNet3d, posenet3d_50 and soft-argmax -- none of them is a product path.
DOWNSTREAM = ("models/feature_extraction.py", "models/feature_propagation.py", "models/posenet3d_50.py",
              "unet/unet3d.py", "unet/__init__.py", "utils/criterion.py")


def stage(verbose=True):
    if not os.path.isdir(REF):
        return False
    for rel in FILES + DOWNSTREAM:
        src, dst = os.path.join(REF, rel), os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        if verbose:
            with open(dst, "rb") as f:
                print(f"baseline/_ref/{rel}  sha256 {hashlib.sha256(f.read()).hexdigest()[:16]}")
    return True


if __name__ == "__main__":
    if not stage():
        sys.exit("reference tree not present; baseline/_ref can only be staged in the build container")
