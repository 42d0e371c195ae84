"""Per-scan time series of the C2 FrontEnd loop (where does set_target's time go as the local map grows?)
  python profiles/c2_series.py [n_scans]"""
import sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from ndt_slam_b200 import host_api as ha, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
seq = synth.c2_sequence(seed=2, n_scans=2000)
odo = np.column_stack([seq["odo"][:, 0], seq["odo"][:, 1], np.rad2deg(seq["odo"][:, 2])])
odo[:, 2] = (odo[:, 2] + 180.0) % 360.0 - 180.0
ha.set_params(Resolution=0.5)
warm = ha.Slam()
for i in range(5):
    warm.process(i, odo[i], seq["scans"][i])
del warm
slam = ha.Slam()
prev = slam.stats(); t_prev = time.perf_counter()
keys = ["set_target_wall_ms", "device_grid_ms", "align_wall_ms", "device_match_ms", "growmap_ms", "estimate_ms", "set_source_wall_ms", "host_filter_ms", "resample_ms"]
print("scan  wall/scan | " + " ".join(k[:14].rjust(14) for k in keys) + " | target pts")
for i in range(n):
    slam.process(i, odo[i], seq["scans"][i])
    if (i + 1) % 25 == 0:
        st = slam.stats(); t = time.perf_counter()
        print(f"{i + 1:4d} {(t - t_prev) / 25 * 1e3:9.3f} | " + " ".join(f"{(st[k] - prev[k]) / 25:14.3f}" for k in keys) + f" | {slam.local_map().shape[0]}")
        prev = st; t_prev = time.perf_counter()
