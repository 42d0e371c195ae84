"""Print the key figures of bench.py JSON lines:  python profiles/show_bench.py gpurun_out/r2/bench_n2.json ..."""
import json
import sys

for path in sys.argv[1:]:
    l = json.loads(open(path).read().strip().splitlines()[-1])
    c5 = l.get("c5") or {}
    w = l.get("weak") or {}
    print(f"{path}: N={l['n_gpus']} value {l['value'] / 1e9:.1f} G pe/s  {l['ms_per_step']:.3f} ms/step  {l['matches_per_sec'] / 1e6:.2f} M matches/s  "
          f"e2e {l['e2e']['value'] / 1e9:.1f} G  executed {l['executed']['value'] / 1e9:.1f} G  bcast {l.get('grid_broadcast_ms')} ms "
          f"({l.get('grid_blob_bytes')} B)  roofline frac {l['roofline'].get('frac')}")
    if w:
        print(f"    weak: {w['value'] / 1e9:.1f} G pe/s  {w['ms_per_step']:.3f} ms/step  {w['matches_per_sec'] / 1e6:.2f} M matches/s ({w['hypotheses_total']} hypotheses)")
    if c5 and "error" not in c5:
        print(f"    c5: {c5['matches_per_sec'] / 1e6:.2f} M pairs/s  {c5['ms_per_step']:.3f} ms  e2e {c5['e2e']['matches_per_sec'] / 1e6:.2f} M  "
              f"e2e_xy {c5['e2e_xy']['matches_per_sec'] / 1e6:.2f} M")
    elif c5:
        print("    c5 error:", c5["error"])
    print("    clocks", l.get("clocks"), " reloc err", l.get("reloc_best_error_m"), " parity", (l.get("parity") or {}).get("within_bar"))
