"""Timing breakdown of one estimatePose-sized problem taken from the C2 sequence (large local map).
  python profiles/c2_state_bench.py [n_scans_to_reach_state]"""
import sys, time
from pathlib import Path
import numpy as np
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from ndt_slam_b200 import capi, host_api as ha, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 190
seq = synth.c2_sequence(seed=2, n_scans=2000)
odo = np.column_stack([seq["odo"][:, 0], seq["odo"][:, 1], np.rad2deg(seq["odo"][:, 2])])
odo[:, 2] = (odo[:, 2] + 180.0) % 360.0 - 180.0
ha.set_params(Resolution=0.5)
slam = ha.Slam()
for i in range(n):
    slam.process(i, odo[i], seq["scans"][i])
tgt = slam.local_map(); poses = slam.poses()
src = ha.voxel_filter(synth.to_xyzw(ha.resample(seq["scans"][n])), 0.05)
guess = np.array([poses[-1, 0], poses[-1, 1], np.deg2rad(poses[-1, 2])])
print("target", tgt.shape, "source", src.shape, "stats", slam.stats())
g = capi.Ndt(capi.default_params(resolution=0.5))
torch.cuda.cudart().cudaProfilerStart()      # ncu --profile-from-start off: only the state benchmark below
d_t = torch.from_numpy(tgt).cuda()
def timeit(f, k=10):
    f(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(k): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / k * 1e3
print("set_target host   wall ms", timeit(lambda: g.set_target(tgt)), "device ms", g.last_kernel_ms())
print("set_target device wall ms", timeit(lambda: g.set_target(d_t.data_ptr(), n=tgt.shape[0], space=capi.MEM_DEVICE)), "device ms", g.last_kernel_ms())
gi = g.grid_info(); print("leaves", gi.n_leaves, "tree", gi.n_slots, "cells", list(gi.div_b))
rb = g.grid_readback(); print("bucket sizes: max", np.abs(rb["nr_points"]).max(), "mean", np.abs(rb["nr_points"]).mean())
print("set_source        wall ms", timeit(lambda: g.set_source(src)))
print("align             wall ms", timeit(lambda: g.align(guess)), "kernel ms", g.last_kernel_ms(), "evals", g.align(guess).evals)
# sporadic launches (one per 8 ms, like the per-scan loop): does the device time change when the GPU idles in between?
import subprocess
for gap in (0.0, 0.002, 0.008, 0.03):
    ms = []
    for _ in range(40):
        time.sleep(gap)
        g.set_target(tgt); ms.append(g.last_kernel_ms())
    clk = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,pstate,power.draw", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
    print(f"gap {gap*1e3:.0f} ms: set_target device ms median {np.median(ms):.3f} min {np.min(ms):.3f} max {np.max(ms):.3f} | {clk}")
