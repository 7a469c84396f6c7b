"""Per-source-line table of an ncu --set full --import-source capture: samples, warp instructions and the dominant
stall reasons, attributed to the OUTERMOST inlined-at line in the given file (so helper code counts for its call site).
With `inner` as an extra argument the INNERMOST location is used instead (file:line of the code itself).
Usage: ncu_lines.py <rep> <kernel-substr> <nvdisasm -gi output> <file.cu> [top=40] [inner] [name:lo-hi ...]"""
import csv
import re
import subprocess
import sys
from collections import defaultdict

rep, kern, dis, fname = sys.argv[1:5]
top = 40
regions = []
inner = False
for a in sys.argv[5:]:
    if a == "inner":
        inner = True
    elif ":" in a:
        n, r = a.split(":")
        lo, hi = r.split("-")
        regions.append((n, int(lo), int(hi)))
    else:
        top = int(a)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi_ = next(i for i, r in enumerate(rows) if "Address" in r and "Source" in r)
hdr = rows[hi_]
ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
sass = []
for r in rows[hi_ + 1:]:
    if len(r) < len(hdr):
        continue
    try:
        sass.append((int(r[ix["Address"]], 16), float(r[ix["# Samples"]] or 0), float(r[ix["Instructions Executed"]] or 0),
                     float(r[ix["Thread Instructions Executed"]] or 0), [float(r[ix[c]] or 0) for c in stall_cols]))
    except ValueError:
        continue
base = sass[0][0]
line_of = {}
infn = False
cur = None
in_block = False
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
        if inner:
            if not in_block:  # the first annotation line of a block is the innermost location
                cur = (m.group(1).split("/")[-1] + ":" + m.group(2))
            in_block = True
        elif m.group(3) is None:
            cur = int(m.group(2)) if m.group(1).endswith(fname) else None
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", l)
    if m:
        line_of[int(m.group(1), 16)] = cur
        in_block = False
agg = defaultdict(lambda: [0.0, 0.0, 0.0, [0.0] * len(stall_cols)])
ts = ti = 0.0
tot_st = [0.0] * len(stall_cols)
for addr, samp, inst, tinst, st in sass:
    ln = line_of.get(addr - base)
    a = agg[ln]
    a[0] += samp; a[1] += inst; a[2] += tinst
    for k, v in enumerate(st):
        a[3][k] += v
        tot_st[k] += v
    ts += samp; ti += inst


def fmt(st, tot):
    o = sorted(((v, c) for v, c in zip(st, stall_cols) if v > 0), reverse=True)[:4]
    return " ".join(f"{c[6:]}={100 * v / max(tot, 1):.0f}%" for v, c in o)


print(f"total samples {ts:.0f}, warp instructions {ti:.3e}; stalls: {fmt(tot_st, sum(tot_st))}")
if regions:
    print("-- regions")
    for n, lo, hi in regions:
        s = [0.0, 0.0, 0.0, [0.0] * len(stall_cols)]
        for ln, a in agg.items():
            if isinstance(ln, int) and lo <= ln <= hi:
                s[0] += a[0]; s[1] += a[1]; s[2] += a[2]
                for k in range(len(stall_cols)):
                    s[3][k] += a[3][k]
        print(f"{n:14s} {lo:5d}-{hi:<5d} samples {100 * s[0] / ts:5.1f}%  instr {100 * s[1] / ti:5.1f}%  lanes {s[2] / max(s[1], 1):5.1f}  {fmt(s[3], sum(s[3]))}")
print("-- lines")
for ln, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{str(ln):>6s} samples {100 * a[0] / ts:5.1f}%  instr {100 * a[1] / ti:5.1f}%  lanes {a[2] / max(a[1], 1):5.1f}  {fmt(a[3], sum(a[3]))}")
