"""Builds the committed round summaries from the ncu exports in gpurun_out/ (run on the CPU box after a capture):
    python profiles/make_summary.py r1
Writes profiles/<round>_summary.md, profiles/traffic.json and copies the small CSVs next to them."""
import csv
import io
import json
import shutil
import subprocess
import sys
from contextlib import redirect_stdout
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
G, P = ROOT / "gpurun_out", ROOT / "profiles"
rnd = sys.argv[1] if len(sys.argv) > 1 else "r1"

KEYS = [("gpu__time_duration.sum", "duration"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__registers_per_thread", "regs/thread"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "active lanes / inst"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64 pipe %"),
        ("smsp__inst_executed.sum", "warp instructions"),
        ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput %"),
        ("lts__t_bytes.sum", "L2 bytes"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
        ("l1tex__t_bytes.sum", "L1 bytes"), ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
        ("smsp__sass_inst_executed_op_local_ld.sum", "local loads (inst)"), ("smsp__sass_inst_executed_op_local_st.sum", "local stores (inst)")]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}


def raw_rows(path):
    rows = list(csv.reader(open(path)))
    return rows[0], rows[1], rows[2:]


def short(name):
    return name.replace("ndt::<unnamed>::", "").replace("void ", "").split("(")[0][:40]


def metrics_table(path, out):
    h, u, rows = raw_rows(path)
    names = [short(r[h.index("Kernel Name")]) for r in rows]
    out.write("| metric | " + " | ".join(f"`{n}`" for n in names) + " |\n|---|" + "---:|" * len(names) + "\n")
    for k, label in KEYS:
        if k not in h:
            continue
        i = h.index(k)
        vals = []
        for r in rows:
            try:
                v = float(r[i]); vals.append(f"{v:,.2f} {u[i]}".strip())
            except ValueError:
                vals.append(r[i])
        out.write(f"| {label} | " + " | ".join(vals) + " |\n")
    out.write("\n")


def bytes_of(h, u, r, key):
    i = h.index(key)
    return float(r[i]) * UNIT.get(u[i], 1.0)


def point_evals_per_launch(executed=False):
    """Point-evals of the captured launch = point_evals_per_step of the plain bench line of the same command
    (executed=True: the passes that ran on the device, the line's `executed` block)."""
    p = G / f"{rnd}_bench_plain.json"
    try:
        line = json.loads(p.read_text().strip().splitlines()[-1])
        return float(line["executed"]["point_evals_per_step"] if executed else line["point_evals_per_step"])
    except Exception:
        return None


def tool(script, *args):
    return subprocess.run([sys.executable, str(P / script), *map(str, args)], capture_output=True, text=True).stdout


out = io.StringIO()
out.write(f"# Round {rnd[1:]} profile summary (B200, sm_100a)\n\nProduced by `profiles/capture_{rnd}.sh` under gpurun; exported with "
          "`ncu -i ... --page raw --csv` / `--page source --print-source cuda,sass --csv`; this file by `profiles/make_summary.py`.\n"
          "Per-launch times under ncu are cold-cache and serialised: compare shares, not absolutes.\n\n")
lb = G / f"{rnd}_launches_bench.csv"
if lb.exists():
    out.write("## Launch list of the bench command (`python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline --c5-pairs 2048`)\n\n")
    out.write(tool("summarize.py", lb) + "\n")
    shutil.copy(lb, P / lb.name)
traffic = {}
fw = G / f"{rnd}_full_k_align_warp.raw.csv"
if fw.exists():
    out.write("## `k_align_warp` (C4: 65,536 hypotheses x 1,218 source points, one launch of the bench command), `ncu --set full`\n\n")
    metrics_table(fw, out)
    h, u, rows = raw_rows(fw)
    r = rows[0]
    traffic = {"k_align_warp_C4_bytes_per_launch": bytes_of(h, u, r, "dram__bytes_read.sum") + bytes_of(h, u, r, "dram__bytes_write.sum"),
               "dram_read_bytes": bytes_of(h, u, r, "dram__bytes_read.sum"), "dram_write_bytes": bytes_of(h, u, r, "dram__bytes_write.sum"),
               "l2_bytes": (bytes_of(h, u, r, "lts__t_bytes.sum") if "lts__t_bytes.sum" in h else
                            (float(r[h.index("lts__t_sectors.sum")]) * 32.0 if "lts__t_sectors.sum" in h else None)),
               "l1_to_l2_write_bytes": bytes_of(h, u, r, "l1tex__m_l1tex2xbar_write_bytes.sum") if "l1tex__m_l1tex2xbar_write_bytes.sum" in h else None,
               "l2_to_l1_read_bytes": bytes_of(h, u, r, "l1tex__m_xbar2l1tex_read_bytes.sum") if "l1tex__m_xbar2l1tex_read_bytes.sum" in h else None,
               "point_evals_per_launch": point_evals_per_launch(),
               "point_evals_run_per_launch": point_evals_per_launch(executed=True),
               "issue_slots_busy_pct": float(r[h.index("smsp__issue_active.avg.pct_of_peak_sustained_active")]),
               "fp64_pipe_pct": float(r[h.index("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active")]),
               "l1_hit_pct": float(r[h.index("l1tex__t_sector_hit_rate.pct")]), "l2_hit_pct": float(r[h.index("lts__t_sector_hit_rate.pct")]),
               "warp_instructions": float(r[h.index("smsp__inst_executed.sum")]),
               "source": f"profiles/{rnd}_full_k_align_warp.raw.csv (ncu --set full --clock-control none, one launch inside bench.py)"}
    shutil.copy(fw, P / fw.name)
    sass = G / f"{rnd}_full_k_align_warp.sass.csv"
    if sass.exists():
        t = tool("ncu_tools.py", fw, sass)
        out.write("Opcode mix, stall reasons (warp-state samples) and hottest SASS lines:\n\n```\n" + "\n".join(t.splitlines()[20:75]) + "\n```\n\n")
    ss = G / f"{rnd}_full_k_align_warp.srcsass.csv"
    if ss.exists():
        out.write("Instructions / stall samples per CUDA source line (top 30):\n\n```\n" + tool("ncu_lines.py", ss, 30) + "```\n\n")
fp = G / f"{rnd}_full_k_align_pairs.raw.csv"
if fp.exists():
    out.write("## `k_align_pairs` (C5: 8,192 scan pairs, one warp per pair), `ncu --set full`\n\n")
    metrics_table(fp, out)
    shutil.copy(fp, P / fp.name)
ft = G / f"{rnd}_full_k_align_team.raw.csv"
if ft.exists():
    out.write("## `k_align_team<4>` (4,096 hypotheses of C4 in one call: a team of four warps per match), `ncu --set full`\n\n")
    metrics_table(ft, out)
    shutil.copy(ft, P / ft.name)
fg = G / f"{rnd}_full_grid_c3.raw.csv"
if fg.exists():
    out.write("## Grid-build kernels at C3 size (4.0 M target points, 0.1 m cells, 3930 x 3952 cells), `ncu --set full`\n\n")
    metrics_table(fg, out)
    shutil.copy(fg, P / fg.name)
lc = G / f"{rnd}_launches_c2c3.csv"
if lc.exists():
    out.write("## Launch list: C2-like (300 k-point local map, 0.5 m cells) and C3 grid build + single match (`profiles/prof_c2c3.py`)\n\n")
    out.write(tool("summarize.py", lc) + "\n")
    shutil.copy(lc, P / lc.name)
(P / f"{rnd}_summary.md").write_text(out.getvalue())
if traffic:
    (P / "traffic.json").write_text(json.dumps(traffic, indent=1) + "\n")
print(out.getvalue()[:3000])
