"""Small-batch relocalisation (4,096 hypotheses: k_align_team<4>, a team of four warps per match) for ncu:
    python profiles/prof_team.py [n_hypotheses]"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import torch

import bench
from ndt_slam_b200 import capi

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
wl = bench.build_c4(65536)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
g = capi.Ndt(capi.default_params(resolution=0.5, stream=stream.cuda_stream))
g.set_target(wl["tgt"]); g.set_source(wl["src"])
d_h = torch.from_numpy(np.ascontiguousarray(wl["hyp"][:n])).cuda()
d_r = torch.zeros(n * capi.RESULT_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
ms = []
for it in range(6):
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record(stream); g.align_batch(d_h.data_ptr(), n=n, space=capi.MEM_DEVICE, out=d_r.data_ptr(), want_fitness=False); b.record(stream)
    torch.cuda.synchronize(); ms.append(a.elapsed_time(b))
print("hypotheses", n, "ms per call", float(np.median(ms[2:])))
