"""How often does a growing local map take the incremental path, and what does a call cost? (GPU box)
Replays the C2 local maps (host PointCloudMap) into one handle with the settled-prefix hints the host classes give."""
import sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from ndt_slam_b200 import capi, synth, host_api as ha
import ndt_common as common
from oracle import oracle_api as oa

n = int(sys.argv[1]) if len(sys.argv) > 1 else 400
seq = synth.c2_sequence(seed=2, n_scans=2000)
traj = seq["traj"]
g = capi.Ndt(capi.default_params(resolution=0.5))
gf = capi.Ndt(capi.default_params(resolution=0.5))
# local maps per scan from the host map classes (truth poses): [prev submap][thinned prefix][tail]
L = ha.load()
import ctypes as C
ha.set_params(sepThre=10.0, LeafSize=0.05)
rows = []
prev = None; settled_prev = 0
# drive a PointCloudMap through the harness one scan at a time is not exposed; emulate: prefix grows, tail replaced
pts_all = []
for i in range(n):
    sc = synth.transform(oa.resample(seq["scans"][i], 0.05, 0.25), traj[i])
    pts_all.append(synth.to_xyzw(sc))
cloud = np.zeros((0, 4), np.float32); settled = 0
for i in range(n):
    # settled part: all earlier scans; tail: this scan (a stand-in with the same sizes as the real local map's growth)
    new_cloud = np.ascontiguousarray(np.concatenate([cloud[:settled], pts_all[i - 1] if i else np.zeros((0, 4), np.float32), pts_all[i]]))
    settled_new = settled + (pts_all[i - 1].shape[0] if i else 0)
    t0 = time.perf_counter()
    g.set_target(new_cloud, n_same=settled if i else 0, n_stable=settled_new)
    t1 = time.perf_counter()
    inc = g.grid_info().reserved
    ms = g.last_kernel_ms()
    t2 = time.perf_counter()
    gf.set_target(new_cloud, n_same=settled if i else 0)
    t3 = time.perf_counter()
    rows.append((i, new_cloud.shape[0], inc, ms, (t1 - t0) * 1e3, gf.last_kernel_ms(), (t3 - t2) * 1e3))
    cloud, settled = new_cloud, settled_new
r = np.array(rows)
inc = r[:, 2] == 1
print("calls", n, "incremental", int(inc.sum()), "points at end", int(r[-1, 1]))
print("incremental: device ms median %.4f  wall ms median %.4f" % (np.median(r[inc, 3]), np.median(r[inc, 4])))
print("full (same clouds, prefix upload): device ms median %.4f  wall ms median %.4f" % (np.median(r[inc, 5]), np.median(r[inc, 6])))
late = inc & (r[:, 0] > n * 0.75)
print("last quarter: incremental %.4f / %.4f ms   full %.4f / %.4f ms" % (np.median(r[late, 3]), np.median(r[late, 4]), np.median(r[late, 5]), np.median(r[late, 6])))
