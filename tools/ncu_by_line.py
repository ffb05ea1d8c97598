"""Join ncu per-SASS-address instruction counts with nvdisasm line info: dynamic warp-instructions per source line."""
import csv, re, subprocess, sys, collections
rep, skip, kern_pat = sys.argv[1], int(sys.argv[2]), sys.argv[3]
src = subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','sass','--launch-skip',str(skip),'--launch-count','1'],capture_output=True,text=True).stdout
rows = list(csv.reader(src.splitlines()))
h = rows[1]; ia = h.index('Address'); ie = h.index('Instructions Executed'); ismp = h.index('# Samples'); isrc = h.index('Source')
dyn = [(int(r[ie]), int(r[ismp]), r[isrc]) for r in rows[2:] if len(r) > ie and r[ie].isdigit()]
if len(dyn) >= 2 and dyn[0][2] == dyn[1][2] and dyn[2][2] == dyn[3][2]: dyn = dyn[::2]
# static line map in program order for the matching kernel
txt = open('/tmp/cub/disasm.txt').read()
sec = [s for s in re.split(r'\n//-+ \.text\.', txt) if re.match(kern_pat, s)]
assert sec, "kernel not found"
lines = []; cur = None
for l in sec[0].split('\n'):
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if re.search(r'/\*[0-9a-f]{4,}\*/', l): lines.append(cur)
print("static", len(lines), "dynamic rows", len(dyn))
n = min(len(lines), len(dyn))
agg = collections.Counter(); smp = collections.Counter()
for (cnt, s, _), ln in zip(dyn[:n], lines[:n]):
    agg[ln] += cnt; smp[ln] += s
tot = sum(agg.values()); ts = sum(smp.values())
print("total", tot)
for ln, c in agg.most_common(int(sys.argv[4]) if len(sys.argv) > 4 else 25):
    text = ""
    try:
        import glob
        f = glob.glob('/root/repo/hiddenpose_b200/csrc/' + ln[0]) if ln else []
        if f: text = open(f[0]).read().split('\n')[ln[1]-1].strip()[:90]
    except Exception: pass
    print(f"{100*c/tot:5.1f}% inst {100*smp[ln]/ts:5.1f}% smp  {ln[0] if ln else None}:{ln[1] if ln else 0}  {text}")
