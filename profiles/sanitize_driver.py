"""Small end-to-end driver for compute-sanitizer (one tool per run):
    compute-sanitizer --tool memcheck|racecheck|initcheck|synccheck python profiles/sanitize_driver.py
C1 ndt_align (cluster + block kernels, fitness), a 256-hypothesis ndt_align_batch on the C4 grid (k_align_warp),
a 3,000-pose ndt_eval_batch (k_eval_warp), a 64-pair ndt_match_pairs (both schedules) and the voxel filter."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import ndt_common as common  # noqa: E402
from ndt_slam_b200 import capi, synth  # noqa: E402
from oracle import oracle_api as oa  # noqa: E402  (data preparation only)

which = set(sys.argv[1:]) or {"c1", "c4", "pairs"}
prm = common.params(resolution=0.5)
if "c1" in which:
    pb = common.c1_problem()
    g = capi.Ndt(prm)
    g.set_target(pb["tgt"]); g.set_source(pb["src"])
    r = g.align(pb["guess"])
    print("c1 align", list(r.pose), r.iters, r.evals, r.fitness)
    small = np.ascontiguousarray(pb["src"][::3])
    g.set_source(small)
    print("c1 block", list(g.align(pb["guess"]).pose))
    print("voxel", g.approx_voxel_filter(pb["src"], 0.1).shape)
    e = g.eval(pb["guess"]); print("eval", e.score, e.n_pairs)
if "c4" in which:
    d = synth.c4_reloc(seed=4)
    src = oa.approx_voxel_filter(synth.to_xyzw(common.prep_scan(d["scan"])), 0.05)
    g = capi.Ndt(prm)
    g.set_target(synth.to_xyzw(d["map_pts"])); g.set_source(src)
    hyp = np.ascontiguousarray(d["hypotheses"][::256][:256])
    res = g.align_batch(hyp)
    print("c4 batch", int(res["iters"].sum()), int(res["evals"].sum()))
    out = g.eval_batch(np.ascontiguousarray(d["hypotheses"][::21][:3000]))
    print("c4 sweep", float(out[:, 0].sum()))
if "pairs" in which:
    srcs, tgts = [], []
    for i in range(64):
        dd = synth.c5_pair(i)
        tgts.append(synth.to_xyzw(common.prep_scan(dd["scan_a"]))); srcs.append(synth.to_xyzw(common.prep_scan(dd["scan_b"])))
    def pack(cl):
        off = np.zeros(len(cl) + 1, np.int64); off[1:] = np.cumsum([c.shape[0] for c in cl])
        return np.ascontiguousarray(np.concatenate(cl, axis=0), dtype=np.float32), off
    s, so = pack(srcs); t, to = pack(tgts)
    g = capi.Ndt(prm)
    res = g.match_pairs(s, so, t, to, np.zeros((64, 3)), 64, source_leaf=0.05)
    print("pairs", int(res["iters"].sum()), float(res["fitness"].sum()))
print("sanitize_driver done")
