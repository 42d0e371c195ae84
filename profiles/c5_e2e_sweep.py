"""C5 end to end through ndt_match_pairs with pinned HOST clouds (run on the GPU box):
    python profiles/c5_e2e_sweep.py [n_pairs]
Sweeps ndt_params.pairs_batch_points (0 = the library's default for host inputs) and prints ms per call and pairs/s;
the results of every setting must be identical (digest)."""
import hashlib
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import torch

import bench
from ndt_slam_b200 import capi

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
c5 = bench.build_c5(0, n)
h_src, h_tgt = torch.from_numpy(c5["src"]).pin_memory(), torch.from_numpy(c5["tgt"]).pin_memory()
h_res = torch.zeros(n * capi.RESULT_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
res = h_res.numpy().view(capi.RESULT_DTYPE)
guesses = np.zeros((n, 3))
nt = c5["tgt"].shape[0]
for bp in (0, 32_000_000, nt // 2 + 1, nt // 3 + 1, nt // 4 + 1, nt // 6 + 1, nt // 8 + 1):
    g = capi.Ndt(capi.default_params(resolution=0.5, pairs_batch_points=bp))
    ts = []
    for it in range(7):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        g.match_pairs(h_src.numpy(), c5["so"], h_tgt.numpy(), c5["to"], guesses, n, source_leaf=0.05, space=capi.MEM_HOST, out=res)
        ts.append((time.perf_counter() - t0) * 1e3)
    d_s, d_t, d_g = torch.from_numpy(c5["src"]).cuda(), torch.from_numpy(c5["tgt"]).cuda(), torch.zeros((n, 3), dtype=torch.float64, device="cuda")
    d_r = torch.zeros(n * capi.RESULT_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
    td = []
    for it in range(6):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        g.match_pairs(d_s.data_ptr(), c5["so"], d_t.data_ptr(), c5["to"], d_g.data_ptr(), n, source_leaf=0.05, space=capi.MEM_DEVICE, out=d_r.data_ptr())
        torch.cuda.synchronize()
        td.append((time.perf_counter() - t0) * 1e3)
    dev_ms = float(np.median(td[2:]))
    dig = hashlib.sha256(res["pose"].tobytes() + res["fitness"].tobytes() + res["evals"].tobytes()).hexdigest()[:12]
    ms = float(np.median(ts[2:]))
    hx, tx = torch.from_numpy(np.ascontiguousarray(c5["src"][:, :2])).pin_memory(), torch.from_numpy(np.ascontiguousarray(c5["tgt"][:, :2])).pin_memory()
    tx_ms = []
    for it in range(7):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        g.match_pairs(hx.numpy(), c5["so"], tx.numpy(), c5["to"], guesses, n, source_leaf=0.05, space=capi.MEM_HOST, out=res, xy=True)
        tx_ms.append((time.perf_counter() - t0) * 1e3)
    xy_ms = float(np.median(tx_ms[2:]))
    print(json.dumps({"pairs": n, "pairs_batch_points": bp, "xy_ms_per_call": xy_ms, "xy_pairs_per_sec": n / (xy_ms * 1e-3), "ms_per_call": ms, "device_inputs_ms_per_call_host_clock": dev_ms, "pairs_per_sec": n / (ms * 1e-3), "digest": dig}), flush=True)
