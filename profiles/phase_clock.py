"""Per-phase cycle counts of the single-match kernels (diagnostic build -DNDT_PHASE_CLOCK=1):
   NDT_B200_LIB=ndt_slam_b200/build/libndt_b200_phase.so python profiles/phase_clock.py"""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import ndt_common as common
import bench
from ndt_slam_b200 import capi
pb = common.c1_problem()
g = capi.Ndt(capi.default_params(resolution=0.5))
g.set_target(pb["tgt"]); g.set_source(pb["src"])
for _ in range(2):
    r = g.align(pb["guess"])
print("C1 cluster: evals", r.evals, "kernel ms", g.last_kernel_ms(), flush=True)
small = np.ascontiguousarray(pb["src"][::2])
g.set_source(small)
for _ in range(2):
    r = g.align(pb["guess"])
print("C1 block<tile> (427 pts): evals", r.evals, "kernel ms", g.last_kernel_ms(), flush=True)
c5 = bench.build_c5(0, 64)
for sched in (capi.PAIRS_CTA, capi.PAIRS_WARP):
    g5 = capi.Ndt(capi.default_params(resolution=0.5, pairs_schedule=sched))
    for _ in range(2):
        r5 = g5.match_pairs(c5["src"], c5["so"], c5["tgt"], c5["to"], np.zeros((64, 3)), 64, source_leaf=0.05)
    print("pairs sched", sched, ": evals[0..2]", r5["evals"][:3], "kernel ms", g5.last_kernel_ms(), flush=True)
    g5.synchronize()
