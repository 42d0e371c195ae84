"""Turns an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel share table.
  python profiles/summarize.py gpurun_out/launches.csv > profiles/rNN_launches.md"""
import csv
import re
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0] != "ID"]
tot = defaultdict(lambda: [0, 0.0])
order = []
for r in rows:
    name = re.sub(r"\(.*", "", r[4]).replace("ndt::<unnamed>::", "").replace("void ", "").strip()
    if name not in tot:
        order.append(name)
    tot[name][0] += 1
    tot[name][1] += float(r[14]) / 1e6
total = sum(v[1] for v in tot.values())
print("| kernel | launches | total ms | share | avg ms |")
print("|---|---:|---:|---:|---:|")
for k in sorted(tot, key=lambda k: -tot[k][1]):
    n, ms = tot[k]
    print(f"| `{k[:70]}` | {n} | {ms:.4f} | {ms / total * 100:.1f}% | {ms / n:.4f} |")
print(f"\ntotal {total:.3f} ms over {len(rows)} launches (cold-cache, serialised under ncu: compare shares, not absolutes)")
