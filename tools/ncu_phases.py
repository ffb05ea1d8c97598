#!/usr/bin/env python
"""Per-phase view of one kernel launch in an ncu report: the SASS is cut at every BAR.SYNC, and for each segment the
static instruction count, warp-level samples, stall reasons and opcode mix are printed.
    python tools/ncu_phases.py report.ncu-rep <launch index> [--dump]"""
import collections
import csv
import re
import subprocess
import sys

rep, skip = sys.argv[1], int(sys.argv[2])
dump = "--dump" in sys.argv
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--launch-skip", str(skip),
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
print(rows[0][1][:150])
h = rows[1]
ia, isrc, ismp, iex = h.index("Address"), h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
stall_cols = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
seen, data = set(), []
for r in rows[2:]:
    if len(r) <= iex or not r[iex].isdigit() or r[ia] in seen:
        continue
    seen.add(r[ia])
    data.append(r)
segs, cur = [], []
for r in data:
    cur.append(r)
    if re.search(r"\bBAR\b", r[isrc]):
        segs.append(cur)
        cur = []
if cur:
    segs.append(cur)
tot_smp = sum(int(r[ismp]) for r in data)
tot_ex = sum(int(r[iex]) for r in data)
print(f"static {len(data)} instr, executed {tot_ex} warp-instr, samples {tot_smp}")
for k, seg in enumerate(segs):
    smp = sum(int(r[ismp]) for r in seg)
    ex = sum(int(r[iex]) for r in seg)
    st = collections.Counter()
    ops = collections.Counter()
    for r in seg:
        for i in stall_cols:
            if r[i].isdigit():
                st[h[i][6:]] += int(r[i])
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[isrc])
        ops[m.group(2) if m else "?"] += int(r[iex])
    ts = sum(st.values()) or 1
    print(f"-- segment {k}: {len(seg)} static, {100 * ex / tot_ex:.1f}% of executed, {100 * smp / tot_smp:.1f}% of samples")
    print("   stalls: " + ", ".join(f"{o}:{100 * v / ts:.0f}%" for o, v in st.most_common(6)))
    print("   ops: " + ", ".join(f"{o}:{100 * v / max(ex, 1):.0f}%" for o, v in ops.most_common(10)))
    if dump:
        for r in seg:
            top = sorted(((int(r[i]), h[i][6:]) for i in stall_cols if r[i].isdigit() and int(r[i]) > 0), reverse=True)[:2]
            print(f"     {r[ismp]:>5} {r[iex]:>8} {r[isrc][:70]:70s} {top}")
