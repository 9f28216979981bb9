"""Joins an ncu SASS source page (csv) with nvdisasm -g line info to attribute executed
instructions and stall samples to CUDA source lines.
usage: python tools/ncu_lines.py report.ncu-rep engine.cubin kernel_substring [top]"""
import csv
import re
import subprocess
import sys
from collections import defaultdict

rep, cubin, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
dis = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout.splitlines()
# locate function
start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and kname in l)
line_of = {}
cur = None
for l in dis[start + 1:]:
    if l.startswith(".text.") or l.startswith(".section"):
        if line_of:
            break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", l)
    if m:
        line_of[int(m.group(1), 16)] = cur
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True,
                     text=True).stdout.splitlines()
rows = list(csv.reader(out))
h = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
hdr = rows[h]
ia, ie, iss = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
ite = hdr.index("Thread Instructions Executed")
base = None
agg = defaultdict(lambda: [0, 0, 0])
for r in rows[h + 1:]:
    if len(r) <= ie:
        continue
    try:
        addr = int(r[ia], 16) if r[ia].startswith("0x") else int(r[ia])
    except ValueError:
        continue
    if base is None:
        base = addr
    key = line_of.get(addr - base)
    a = agg[key]
    a[0] += int(r[ie] or 0)
    a[1] += int(r[iss] or 0)
    a[2] += int(r[ite] or 0)
ti = sum(a[0] for a in agg.values()) or 1
ts = sum(a[1] for a in agg.values()) or 1
print(f"total warp-instructions {ti}, stall samples {ts}")
srcs = {}
for (key, a) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    txt = ""
    if key:
        f, ln = key
        if f not in srcs:
            import glob
            c = glob.glob(f"/root/repo/**/{f}", recursive=True)
            srcs[f] = open(c[0]).read().splitlines() if c else []
        if 0 < ln <= len(srcs[f]):
            txt = srcs[f][ln - 1].strip()[:90]
    lanes = a[2] / a[0] if a[0] else 0
    print(f"{a[0] / ti * 100:5.1f}% inst {a[1] / ts * 100:5.1f}% stall lanes {lanes:4.1f}  {key}: {txt}")
