"""Join an ncu SASS source page (csv) with nvdisasm -g line info: per CUDA source line, stall samples,
warp instructions and average active lanes. Usage: ncu_lines.py <rep> <kernel-substr> <cubin-disasm.txt> [top]"""
import csv
import re
import subprocess
import sys
from collections import defaultdict

rep, kern, dis = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if "Address" in r and "Source" in r)
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
sass = []
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    try:
        sass.append((int(r[ix["Address"]], 16) if r[ix["Address"]].startswith("0x") else int(r[ix["Address"]]),
                     r[ix["Source"]], float(r[ix["# Samples"]] or 0), float(r[ix["Instructions Executed"]] or 0),
                     float(r[ix["Thread Instructions Executed"]] or 0)))
    except ValueError:
        continue
base = sass[0][0]
# nvdisasm: "//## File "...", line N" lines precede instructions "/*0010*/ ..."
line_of = {}
cur = None
infn = False
for l in open(dis):
    if l.startswith(".text.") and kern in l:
        infn = True
        continue
    if infn and l.startswith(".text.") and kern not in l:
        break
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", l)
    if m:
        line_of[int(m.group(1), 16)] = cur
agg = defaultdict(lambda: [0.0, 0.0, 0.0])
tot_s = tot_i = 0
for addr, src, samp, inst, tinst in sass:
    key = line_of.get(addr - base, ("?", 0))
    a = agg[key]
    a[0] += samp; a[1] += inst; a[2] += tinst
    tot_s += samp; tot_i += inst
srcs = {}
def src_line(f, n):
    import os
    for root in ("limap_b200/csrc",):
        p = os.path.join(root, f)
        if os.path.exists(p):
            if p not in srcs:
                srcs[p] = open(p).read().splitlines()
            return srcs[p][n - 1].strip()[:90] if 0 < n <= len(srcs[p]) else ""
    return ""
print(f"total samples {tot_s:.0f}  warp-instructions {tot_i:.3g}")
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    f, n = key if key else ("?", 0)
    print(f"{a[0] / tot_s * 100:5.1f}% samp {a[1] / tot_i * 100:5.1f}% inst lanes {a[2] / max(a[1], 1):5.1f} | {f}:{n}  {src_line(f, n)}")
