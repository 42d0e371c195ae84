"""Helpers to read ncu CSV exports (run on the CPU box): python profiles/ncu_tools.py raw.csv sass.csv"""
import csv
import sys
from collections import Counter

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum", "sm__cycles_elapsed.max"]


def raw(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("-----", r[hdr.index("Kernel Name")][:100])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"  {k:72s} {units[i]:16s} {r[i]}")


def sass(path, top=30):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
    h = rows[hi]
    data = [r for r in rows[hi + 1:] if len(r) == len(h)]
    cs, ci, csrc = h.index("# Samples"), h.index("Instructions Executed"), h.index("Source")
    tot = sum(int(r[ci]) for r in data if r[ci].isdigit())
    tots = sum(int(r[cs]) for r in data if r[cs].isdigit())
    print("sass instructions", len(data), "executed", tot, "samples", tots)
    op, ops = Counter(), Counter()
    for r in data:
        if not r[ci].isdigit():
            continue
        t = r[csrc].split()
        o = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        op[o] += int(r[ci]); ops[o] += int(r[cs]) if r[cs].isdigit() else 0
    for o, c in op.most_common(18):
        print(f"  {o:10s} inst {c / tot * 100:5.1f}%  samples {ops[o] / max(tots, 1) * 100:5.1f}%")
    st = {h[i]: sum(int(r[i]) for r in data if r[i].isdigit()) for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x}
    for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:9]:
        print(f"  {k:28s} {v / max(tots, 1) * 100:5.1f}%")
    for r in sorted([r for r in data if r[cs].isdigit()], key=lambda r: -int(r[cs]))[:top]:
        print("  ", r[cs].rjust(7), r[ci].rjust(10), r[csrc][:110])


if __name__ == "__main__":
    raw(sys.argv[1])
    if len(sys.argv) > 2:
        sass(sys.argv[2])
