"""Compare two result dumps written by `NDT_AB_DUMP=1 python profiles/ab_c4.py a.so b.so`:
    python profiles/compare_dumps.py gpurun_out/r2/c4_<a>.npy gpurun_out/r2/c4_<b>.npy
Prints how far the two builds' results are apart (pose, score, Hessian, iteration counts)."""
import sys

import numpy as np

a, b = np.load(sys.argv[1]), np.load(sys.argv[2])
assert a.shape == b.shape
dp = np.abs(a["pose"] - b["pose"])
dp[:, 2] = np.abs((dp[:, 2] + np.pi) % (2 * np.pi) - np.pi)
ds = np.abs(a["score"] - b["score"]) / np.maximum(np.abs(b["score"]), 1e-300)
dh = np.abs(a["hess"] - b["hess"]).max(axis=1) / np.maximum(np.abs(b["hess"]).max(axis=1), 1e-300)
same = (a["iters"] == b["iters"]) & (a["evals"] == b["evals"]) & (a["converged"] == b["converged"])
print({"n": int(a.shape[0]), "bit_identical_poses": int((a["pose"] == b["pose"]).all(axis=1).sum()),
       "max_xy_diff_m": float(dp[:, :2].max()), "max_yaw_diff_rad": float(dp[:, 2].max()),
       "max_score_rel": float(ds.max()), "max_hess_rel": float(dh.max()), "same_iters_evals_converged": int(same.sum()),
       "xy_diff_over_1e-6": int((dp[:, :2].max(axis=1) > 1e-6).sum())})
