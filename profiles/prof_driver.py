"""Small fixed workload for ncu captures (run plainly first, then under ncu with the same command line).
  python profiles/prof_driver.py [n_hyp]   -> one C4 batch match of n_hyp hypotheses + 3 C1 single matches"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from ndt_slam_b200 import capi, synth  # noqa: E402
from oracle import oracle_api as oa  # noqa: E402  (data preparation only)

n_hyp = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
wl = bench.build_c4(65536)
prm = capi.default_params(resolution=0.5)
g = capi.Ndt(prm)
g.set_target(wl["tgt"]); g.set_source(wl["src"])
hyp = np.ascontiguousarray(wl["hyp"][:n_hyp])
for _ in range(2):
    res = g.align_batch(hyp)
print("C4 batch", n_hyp, "kernel_ms", g.last_kernel_ms(), "point_evals", int(res["point_evals"].sum()))

d = synth.c1_pair(1)
ra = oa.resample(d["scan_a"], 0.05, 0.25); rb = oa.resample(d["scan_b"], 0.05, 0.25)
tgt = synth.to_xyzw(synth.transform(ra, d["pose_a"]))
src = oa.approx_voxel_filter(synth.to_xyzw(rb), 0.05)
g1 = capi.Ndt(prm)
for _ in range(3):
    g1.set_target(tgt); g1.set_source(src)
    r = g1.align(np.array(d["pose_a"]))
print("C1 align kernel_ms", g1.last_kernel_ms(), "evals", r.evals)
e = g1.eval(np.array(d["pose_a"])); e = g1.eval(np.array(d["pose_a"]))
print("C1 single eval (2 kernels) ms", g1.last_kernel_ms())
torch.cuda.synchronize()
