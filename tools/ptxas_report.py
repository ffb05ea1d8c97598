#!/usr/bin/env python
"""Registers / spills / stack of every kernel in the library: `python tools/ptxas_report.py [filter]`
(rebuilds with -Xptxas -v into a scratch file; the shipped .so is not touched)."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hiddenpose_b200 import _native  # noqa: E402


def main():
    flt = sys.argv[1] if len(sys.argv) > 1 else ""
    extra = sys.argv[2:]
    out = "/tmp/lct_ptxas_report.so"
    cmd = ["nvcc"] + _native.NVCC_FLAGS + ["-Xptxas", "-v", "-I" + _native.INCLUDE, "-I" + _native.CSRC, "-o", out,
                                           os.path.join(_native.CSRC, "lct_api.cu")] + extra
    txt = subprocess.run(cmd, capture_output=True, text=True).stderr
    name = None
    rows = []
    for line in txt.splitlines():
        m = re.search(r"Compiling entry function '(\S+)'", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"^void lct::lct_kernel<lct::", "", name)
            name = re.sub(r">\(lct::Params, int\)$", "", name).replace("lct::", "")
            continue
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
        if m and name:
            stack, sst, sld = map(int, m.groups())
            continue_ = (stack, sst, sld)
            rows.append([name, None, stack, sst, sld])
            continue
        m = re.search(r"Used (\d+) registers", line)
        if m and rows and rows[-1][1] is None:
            rows[-1][1] = int(m.group(1))
    for name, regs, stack, sst, sld in rows:
        if flt in name:
            print(f"{regs:4d} regs  stack {stack:4d}  spill st/ld {sst:4d}/{sld:4d}  {name}")


if __name__ == "__main__":
    main()
