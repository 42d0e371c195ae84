"""Static SASS facts of the shipped CUDA library -> profiles/<round>_sass_notes.md (runs on the CPU box):
    python profiles/make_sass_notes.py r2"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
rnd = sys.argv[1] if len(sys.argv) > 1 else "r2"
sass = subprocess.run(["cuobjdump", "-sass", str(ROOT / "ndt_slam_b200" / "libndt_b200.so")], capture_output=True, text=True).stdout
fn, data = None, collections.defaultdict(collections.Counter)
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1)
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and fn:
        op = m.group(2)
        data[fn][op.split(".")[0]] += 1
        if op.startswith("F2I") and "FLOOR" in op:
            data[fn]["F2I.FLOOR"] += 1
KEEP = ("k_align_warp<true, true>", "k_align_team<2>", "k_align_pairs", "k_eval_warp", "k_inc_update", "k_finalize", "k_count", "k_align_cluster", "k_align_grid")
rows = []
for f, c in data.items():
    d = subprocess.run(["c++filt", f], capture_output=True, text=True).stdout.strip()
    short = re.sub(r"ndt::\(anonymous namespace\)::", "", d).split("(")[0].replace("void ", "")
    if any(k in short for k in KEEP):
        rows.append((short, sum(v for k, v in c.items() if "." not in k), c))
cols = ["FADD2", "FFMA2", "FFMA", "F2I.FLOOR", "DFMA", "DMUL", "DADD", "MUFU", "LDS", "STL", "LDL", "BAR"]
out = [f"# SASS facts of the shipped library (round {rnd[1:]})\n",
       "`cuobjdump -sass ndt_slam_b200/libndt_b200.so`, static instruction counts per kernel; written by `profiles/make_sass_notes.py`.",
       "What to look for: `FADD2` = the packed float32 additions of the transform and of the squared distance (DESIGN.md 4.2); no `FFMA2` in any matcher",
       "(a packed multiply feeding a packed add would be contracted and break the bit-exact transform); `k_eval_warp` -- transform, cell index, radius test and",
       "hit path without optimiser or fitness code -- and `k_count` (cell index of the grid build) hold no `FFMA` at all: the float32 arithmetic that decides",
       "discrete outcomes is unfused (the `FFMA` of the matchers sit in the fp64 library routines and the 1-NN certificate); `F2I.FLOOR` = the one-conversion",
       "cell index; `BAR` > 2 = the named barriers of warp teams / CTA-per-pair; `DFMA` / `DMUL` / `DADD` = fp64 hit path and optimiser; `STL` / `LDL` =",
       "per-thread optimiser state in local memory.\n",
       "| kernel | instructions | " + " | ".join(cols) + " |", "|---|---:|" + "---:|" * len(cols)]
for short, tot, c in sorted(rows):
    out.append(f"| `{short}` | {tot} | " + " | ".join(str(c[k]) for k in cols) + " |")
(ROOT / "profiles" / f"{rnd}_sass_notes.md").write_text("\n".join(out) + "\n")
print("\n".join(out))
