"""Per-CUDA-source-line instruction / stall-sample shares from `ncu -i rep --page source --print-source cuda,sass --csv`.
  python profiles/ncu_lines.py srcsass.csv [top]"""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
cur, hdr = None, None
agg = defaultdict(lambda: [0, 0])
for r in rows:
    if r and r[0] == "File Path":
        cur = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No":
        hdr = r; ci = hdr.index("Instructions Executed"); cs = hdr.index("# Samples"); continue
    if hdr is None or len(r) <= max(ci, cs) or not r[0].isdigit() or r[2] == "" or not r[ci].isdigit():
        continue
    k = (cur, int(r[0]), r[1].strip()[:100])
    agg[k][0] += int(r[ci]); agg[k][1] += int(r[cs]) if r[cs].isdigit() else 0
tot = sum(v[0] for v in agg.values()); tots = sum(v[1] for v in agg.values())
print(f"instructions {tot}  samples {tots}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{v[0] / tot * 100:5.1f}% inst {v[1] / max(tots, 1) * 100:5.1f}% smp  {k[0]}:{k[1]}  {k[2]}")
