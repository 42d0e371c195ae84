"""Shared problem builders for the tests. The oracle is imported here (tests only)."""
from __future__ import annotations

import numpy as np

from ndt_slam_b200 import capi, synth
from oracle import oracle_api as oa

from pathlib import Path

GOLD = Path(__file__).resolve().parent / "golden"
LAUNCH = dict(space=0.05, space_thre=0.25, leaf=0.05, trans_eps=0.01, step_size=0.1, max_iter=35)


def params(resolution=0.5, quirks=capi.QUIRKS_PCL_1_10, **kw):
    p = capi.NdtParams(resolution=resolution, step_size=LAUNCH["step_size"], trans_eps=LAUNCH["trans_eps"],
                       max_iter=LAUNCH["max_iter"], outlier_ratio=0.55, min_points=6, eig_mult=0.01,
                       quirks=quirks, device=0, stream=None)
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def prep_scan(xy):
    """resample (ScanPointResampler) -> float32 cloud, like matchScan + setScanPair"""
    return oa.resample(xy, LAUNCH["space"], LAUNCH["space_thre"])


def c1_problem(seed=1):
    d = synth.c1_pair(seed)
    ra, rb = prep_scan(d["scan_a"]), prep_scan(d["scan_b"])
    tgt = synth.to_xyzw(synth.transform(ra, d["pose_a"]))
    src = oa.approx_voxel_filter(synth.to_xyzw(rb), LAUNCH["leaf"])
    return dict(tgt=tgt, src=src, guess=np.array(d["pose_a"]), truth=np.array(d["pose_b"]), raw=d)


def random_cloud(seed, n, extent=40.0, walls=True):
    rng = synth.rng_for(seed)
    if walls:
        segs = synth.office(seed, extent, extent * 0.6, 12)
        pts = synth.sample_walls(segs, extent * 0.6 * 8 / max(n, 1) + 0.01, 0.01, rng)
        if pts.shape[0] > n:
            pts = pts[rng.permutation(pts.shape[0])[:n]]
    else:
        pts = rng.uniform(-extent, extent, size=(n, 2))
    return synth.to_xyzw(pts)


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    scale = max(np.max(np.abs(b)), 1e-300)
    return float(np.max(np.abs(a - b)) / scale)


def oracle_align_many(prm, tgt, src, guesses, want_fitness=False, threads=None):
    """Full oracle matches of every guess on host threads (one oracle per thread; ctypes drops the GIL)."""
    import os
    import threading
    from concurrent.futures import ThreadPoolExecutor

    guesses = np.asarray(guesses, dtype=np.float64)
    threads = threads or min(os.cpu_count() or 1, 32)
    out = [None] * guesses.shape[0]
    todo = list(range(guesses.shape[0]))
    lock = threading.Lock()

    def work(_):
        o = oa.Oracle(prm)
        o.set_target(tgt); o.set_source(src); o.want_fitness(want_fitness)
        while True:
            with lock:
                if not todo:
                    return
                ids = [todo.pop() for _ in range(min(8, len(todo)))]
            for i in ids:
                out[i] = o.align(guesses[i])

    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(work, range(threads)))
    return out


def oracle_eval_many(prm, tgt, src, poses, want_hessian=True, threads=None):
    """Oracle objective at every pose -> (n, 14) like ndt_eval_batch: score, g[3], H[9], n_pairs."""
    import os
    import threading
    from concurrent.futures import ThreadPoolExecutor

    poses = np.asarray(poses, dtype=np.float64)
    threads = threads or min(os.cpu_count() or 1, 32)
    out = np.zeros((poses.shape[0], 14))
    todo = list(range(poses.shape[0]))
    lock = threading.Lock()

    def work(_):
        o = oa.Oracle(prm)
        o.set_target(tgt); o.set_source(src)
        while True:
            with lock:
                if not todo:
                    return
                ids = [todo.pop() for _ in range(min(16, len(todo)))]
            for i in ids:
                e = o.eval(poses[i], want_hessian)
                out[i, 0] = e.score; out[i, 1:4] = e.grad; out[i, 4:13] = e.hess; out[i, 13] = e.n_pairs

    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(work, range(threads)))
    return out
