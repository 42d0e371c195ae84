"""Per-kernel launch list workload for C2-sized and C3-sized grid builds + matches (run under
`ncu --metrics gpu__time_duration.sum`):  python profiles/prof_c2c3.py"""
import sys
from pathlib import Path
import numpy as np
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from ndt_slam_b200 import capi, synth  # noqa: E402
from oracle import oracle_api as oa  # noqa: E402  (data preparation only)

rng = synth.rng_for(5)
# C2-like: dense local map (300 k points in ~80 x 50 m), one filtered scan
segs = synth.office(2, 40.0, 25.0, 24)
tgt = synth.to_xyzw(np.concatenate([synth.sample_walls(segs, 0.004, 0.012, rng) for _ in range(4)], axis=0)[:300_000])
scan = synth.raycast(segs, (8.0, 6.0, 0.2), rng)
src = oa.approx_voxel_filter(synth.to_xyzw(oa.resample(scan, 0.05, 0.25)), 0.05)
g = capi.Ndt(capi.default_params(resolution=0.5))
for _ in range(2):
    g.set_target(tgt); g.set_source(src); r = g.align([8.02, 5.98, 0.21])
print("C2-like: target", tgt.shape[0], "grid ms", "match ms", g.last_kernel_ms(), "evals", r.evals, "leaves", g.grid_info().n_leaves)
# C3
d3 = synth.c3_dense(seed=3)
t3, s3 = synth.to_xyzw(d3["target"]), synth.to_xyzw(d3["source"])
g3 = capi.Ndt(capi.default_params(resolution=0.1))
dt = torch.from_numpy(t3).cuda()
for _ in range(2):
    g3.set_target(dt.data_ptr(), n=t3.shape[0], space=capi.MEM_DEVICE); gb = g3.last_kernel_ms()
    g3.set_source(s3); r3 = g3.align(list(d3["guess"]))
print("C3: grid ms", gb, "match ms", g3.last_kernel_ms(), "evals", r3.evals)
torch.cuda.synchronize()
