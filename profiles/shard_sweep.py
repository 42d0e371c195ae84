"""What every rank of an N-GPU strong-scaling run would take, measured on ONE GPU: the 65,536 C4 hypotheses cut into N
contiguous shards, each matched with 1 / 2 / 4 warps per match (ndt_params.align_team). max over shards = the N-GPU step.
    python profiles/shard_sweep.py [N]"""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import torch

import bench
from ndt_slam_b200 import capi

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
wl = bench.build_c4(65536)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
out = {}
for team in (1, 2, 4):
    g = capi.Ndt(capi.default_params(resolution=0.5, stream=stream.cuda_stream, align_team=team))
    g.set_target(wl["tgt"]); g.set_source(wl["src"])
    per = []
    for r in range(N):
        lo, hi = bench.shard(65536, r, N)
        d_h = torch.from_numpy(np.ascontiguousarray(wl["hyp"][lo:hi])).cuda()
        d_r = torch.zeros((hi - lo) * capi.RESULT_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
        ms = []
        for it in range(7):
            flush.zero_()
            a, b = torch.cuda.Event(True), torch.cuda.Event(True)
            a.record(stream); g.align_batch(d_h.data_ptr(), n=hi - lo, space=capi.MEM_DEVICE, out=d_r.data_ptr(), want_fitness=False); b.record(stream)
            torch.cuda.synchronize()
            if it >= 2:
                ms.append(a.elapsed_time(b))
        per.append(round(float(np.median(ms)), 4))
    out[f"team{team}"] = {"per_shard_ms": per, "max_ms": max(per), "mean_ms": float(np.mean(per))}
print(json.dumps({"shards": N, **out}))
