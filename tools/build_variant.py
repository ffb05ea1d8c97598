#!/usr/bin/env python
"""A/B builds for kernel work: `python tools/build_variant.py <tag> [--rev <git rev>] [nvcc flags ...]` compiles the library
(from the working tree, or from csrc/include as of a git revision) into build/lib_<tag>.so; run it with
HIDDENPOSE_LCT_LIB=build/lib_<tag>.so.  build/ is git-ignored but travels to the GPU box."""
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hiddenpose_b200 import _native  # noqa: E402


def main():
    tag, args = sys.argv[1], sys.argv[2:]
    csrc, inc = _native.CSRC, _native.INCLUDE
    if args and args[0] == "--rev":
        rev, args = args[1], args[2:]
        tmp = tempfile.mkdtemp(prefix="lct_rev_")
        for sub in ("hiddenpose_b200/csrc", "include"):
            os.makedirs(os.path.join(tmp, sub))
            names = subprocess.check_output(["git", "ls-tree", "--name-only", rev, sub + "/"], cwd=ROOT, text=True).split()
            for n in names:
                with open(os.path.join(tmp, n), "wb") as f:
                    f.write(subprocess.check_output(["git", "show", f"{rev}:{n}"], cwd=ROOT))
        csrc, inc = os.path.join(tmp, "hiddenpose_b200/csrc"), os.path.join(tmp, "include")
    os.makedirs(os.path.join(ROOT, "build"), exist_ok=True)
    out = os.path.join(ROOT, "build", f"lib_{tag}.so")
    subprocess.check_call(["nvcc"] + _native.NVCC_FLAGS + args + ["-I" + inc, "-I" + csrc, "-o", out, os.path.join(csrc, "lct_api.cu")])
    print(out)


if __name__ == "__main__":
    main()
