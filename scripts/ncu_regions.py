"""Attribute ncu SASS samples to code regions (outermost inlined-at line in a given file).
Usage: ncu_regions.py <rep> <kernel-substr> <nvdisasm -gi output> <file.cu> name:lo-hi ..."""
import csv
import re
import subprocess
import sys
from collections import defaultdict

rep, kern, dis, fname = sys.argv[1:5]
regions = []
for a in sys.argv[5:]:
    n, r = a.split(":")
    lo, hi = r.split("-")
    regions.append((n, int(lo), int(hi)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi_ = next(i for i, r in enumerate(rows) if "Address" in r and "Source" in r)
hdr = rows[hi_]
ix = {h: i for i, h in enumerate(hdr)}
sass = []
for r in rows[hi_ + 1:]:
    if len(r) < len(hdr):
        continue
    try:
        sass.append((int(r[ix["Address"]], 16), float(r[ix["# Samples"]] or 0), float(r[ix["Instructions Executed"]] or 0),
                     float(r[ix["Thread Instructions Executed"]] or 0)))
    except ValueError:
        continue
base = sass[0][0]
line_of = {}
infn = False
cur = None
prev_was_inst = True
for l in open(dis):
    if l.startswith(".text.") and kern in l:
        infn = True
        continue
    if infn and l.startswith(".text.") and kern not in l:
        break
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', l)
    if m:
        # the last annotation line of a block (no "inlined at") is the outermost location
        if m.group(3) is None:
            cur = int(m.group(2)) if m.group(1).endswith(fname) else None
        prev_was_inst = False
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", l)
    if m:
        line_of[int(m.group(1), 16)] = cur
        prev_was_inst = True
agg = defaultdict(lambda: [0.0, 0.0, 0.0])
ts = ti = 0
for addr, samp, inst, tinst in sass:
    ln = line_of.get(addr - base)
    name = "other"
    if ln is not None:
        for n, lo, hi in regions:
            if lo <= ln <= hi:
                name = n
                break
    a = agg[name]
    a[0] += samp; a[1] += inst; a[2] += tinst
    ts += samp; ti += inst
for n, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{n:12s} {a[0] / ts * 100:5.1f}% samples  {a[1] / ti * 100:5.1f}% warp-inst  lanes {a[2] / max(a[1], 1):5.1f}")
