"""C5 workload for ncu launch lists / captures: python profiles/prof_c5.py [n_pairs]  (device-resident inputs)"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from ndt_slam_b200 import capi  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
c5 = bench.build_c5(0, n)
g = capi.Ndt(capi.default_params(resolution=0.5))
d_src, d_tgt = torch.from_numpy(c5["src"]).cuda(), torch.from_numpy(c5["tgt"]).cuda()
d_g = torch.zeros((n, 3), dtype=torch.float64, device="cuda")
d_res = torch.zeros(n * capi.RESULT_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
for _ in range(2):
    g.match_pairs(d_src.data_ptr(), c5["so"], d_tgt.data_ptr(), c5["to"], d_g.data_ptr(), n, source_leaf=0.05,
                  space=capi.MEM_DEVICE, out=d_res.data_ptr())
    g.synchronize()
res = np.frombuffer(d_res.cpu().numpy().tobytes(), dtype=capi.RESULT_DTYPE)
print("C5", n, "pairs: device ms", g.last_kernel_ms(), "point_evals", int(res["point_evals"].sum()), "evals/match", res["evals"].mean())
