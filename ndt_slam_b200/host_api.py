"""ctypes binding of libndt_slam_host.so: the C++ mirror of the reference's classes (PoseEstimator,
ScanMatcher, PointCloudMap, ScanPointResampler, PoseFuser, FrontEnd, SlamLauncher) running on top of
the CUDA C ABI. Test / bench plumbing only."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import build as _build
from .capi import NdtResult

# ndt_mapping.launch values (reference config), Resolution per BASELINE configs; removeMoving off for the headline run
LAUNCH_PARAMS = {
    "space": "0.05", "space_thre": "0.25", "LeafSize": "0.05", "TransformationEpsilon": "0.01", "StepSize": "0.1",
    "Resolution": "0.5", "MaximumIterations": "35", "coeNDTCov": "1.0", "score_thre": "0.5", "sepThre": "10.0",
    "removeMoving": "false", "resol": "0.05", "thre_neighbor": "0.2", "delTime": "0.5", "coeVel": "0.1",
    "coeOmega": "0.5", "keyframe_skip": "5", "start_frame": "0",
}
_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not _build.LIB_HOST.exists():
        raise RuntimeError(f"{_build.LIB_HOST} is not built -- run `python -m ndt_slam_b200.build`")
    C.CDLL(str(_build.LIB_CUDA), mode=C.RTLD_GLOBAL)
    L = C.CDLL(str(_build.LIB_HOST))
    vp, i64, dp, d = C.c_void_p, C.c_int64, C.POINTER(C.c_double), C.c_double
    L.host_last_error.restype = C.c_char_p
    L.host_param_set.argtypes = [C.c_char_p, C.c_char_p]
    L.host_resample.argtypes = [vp, i64, vp, i64]; L.host_resample.restype = i64
    L.host_voxel_filter.argtypes = [vp, i64, C.c_float, vp]; L.host_voxel_filter.restype = i64
    L.host_add_angle.argtypes = [d, d]; L.host_add_angle.restype = d
    L.host_sub_angle.argtypes = [d, d]; L.host_sub_angle.restype = d
    L.host_cal_motion.argtypes = [dp, dp, dp]
    L.host_cal_pred_pose.argtypes = [dp, dp, dp]
    L.host_odometry_cov.argtypes = [dp, dp, dp, dp]
    L.host_fuse_pose.argtypes = [dp] * 8
    L.host_estimate_pose.argtypes = [vp, i64, vp, i64, dp, dp, dp, C.POINTER(NdtResult)]; L.host_estimate_pose.restype = d
    L.host_slam_create.restype = vp
    L.host_slam_destroy.argtypes = [vp]
    L.host_slam_process.argtypes = [vp, C.c_int, dp, vp, i64]; L.host_slam_process.restype = C.c_int
    L.host_slam_process_forced.argtypes = [vp, C.c_int, dp, vp, i64, dp, dp]; L.host_slam_process_forced.restype = C.c_int
    L.host_slam_covs.argtypes = [vp, vp, i64]; L.host_slam_covs.restype = i64
    L.host_slam_poses.argtypes = [vp, vp, i64]; L.host_slam_poses.restype = i64
    L.host_slam_local_map.argtypes = [vp, vp, i64]; L.host_slam_local_map.restype = i64
    L.host_slam_global_map.argtypes = [vp, vp, i64]; L.host_slam_global_map.restype = i64
    L.host_slam_submaps.argtypes = [vp]; L.host_slam_submaps.restype = C.c_int
    L.host_slam_stats.argtypes = [vp, dp, C.POINTER(C.c_int64)]
    L.host_launcher_run.restype = C.c_int
    L.host_map_replay.argtypes = [vp, vp, vp, C.c_int, C.c_int, vp, i64, vp, vp, i64, vp]; L.host_map_replay.restype = i64
    _lib = L
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _d(v):
    return (C.c_double * len(v))(*v)


def set_params(**kw):
    L = load()
    L.host_param_clear()
    p = dict(LAUNCH_PARAMS)
    p.update({k: str(v) for k, v in kw.items()})
    for k, v in p.items():
        L.host_param_set(k.encode(), v.encode())


def resample(xy):
    L = load()
    xy = np.ascontiguousarray(xy, np.float64)
    out = np.zeros((4 * xy.shape[0] + 16, 2))
    m = L.host_resample(_p(xy), xy.shape[0], _p(out), out.shape[0])
    assert m >= 0
    return np.ascontiguousarray(out[:m])


def voxel_filter(xyzw, leaf):
    L = load()
    xyzw = np.ascontiguousarray(xyzw, np.float32)
    out = np.zeros_like(xyzw)
    m = L.host_voxel_filter(_p(xyzw), xyzw.shape[0], leaf, _p(out))
    return np.ascontiguousarray(out[:m])


def pcfilter(base_xyzw, test_xyzw):
    """PCFilter::difference_extraction(base, test) and remove_neighborPoint(test, diff) -> (diff xyzw, kept xyzw)."""
    L = load()
    base = np.ascontiguousarray(base_xyzw, np.float32); test = np.ascontiguousarray(test_xyzw, np.float32)
    d, k = np.zeros_like(test), np.zeros_like(test)
    nd, nk = C.c_int64(), C.c_int64()
    L.host_pcfilter.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.POINTER(C.c_int64), C.c_void_p, C.POINTER(C.c_int64)]
    L.host_pcfilter.restype = None
    L.host_pcfilter(_p(base), base.shape[0], _p(test), test.shape[0], _p(d), C.byref(nd), _p(k), C.byref(nk))
    return np.ascontiguousarray(d[: nd.value]), np.ascontiguousarray(k[: nk.value])


def map_replay(poses_deg, scans_map_xy, check_every=0):
    """PointCloudMap alone: (addPose, addPoints, makeLocalMap) per scan, makeGlobalMap at the end.
    Returns (n_submaps, local_map xyzw, global_map xyzw). Raises if the incremental local map ever differs from a
    from-scratch rebuild (check_every > 0)."""
    L = load()
    poses = np.ascontiguousarray(poses_deg, np.float64)
    off = np.zeros(len(scans_map_xy) + 1, np.int64)
    off[1:] = np.cumsum([s.shape[0] for s in scans_map_xy])
    xy = np.ascontiguousarray(np.concatenate(scans_map_xy, axis=0), np.float64)
    cap = int(off[-1]) + 1024
    lo, go = np.zeros((cap, 4), np.float32), np.zeros((cap, 4), np.float32)
    nl, ng = C.c_int64(), C.c_int64()
    ns = L.host_map_replay(_p(poses), _p(xy), _p(off), len(scans_map_xy), check_every, _p(lo), cap, C.byref(nl), _p(go), cap, C.byref(ng))
    if ns < 0:
        raise RuntimeError(f"incremental local map differs from the from-scratch rebuild at scan {-ns - 1}")
    return int(ns), np.ascontiguousarray(lo[: nl.value]), np.ascontiguousarray(go[: ng.value])


def cal_motion(cur, prev):
    L = load(); o = (C.c_double * 3)(); L.host_cal_motion(_d(cur), _d(prev), o); return np.array(o)


def cal_pred_pose(motion, last):
    L = load(); o = (C.c_double * 3)(); L.host_cal_pred_pose(_d(motion), _d(last), o); return np.array(o)


def odometry_cov(motion, last, last_cov):
    L = load(); c = (C.c_double * 9)()
    L.host_odometry_cov(_d(motion), _d(last), _d(list(np.ravel(last_cov))), c)
    return np.array(c).reshape(3, 3)


def fuse_pose(pred, est, motion, last, last_cov, Q):
    L = load(); f = (C.c_double * 3)(); c = (C.c_double * 9)()
    L.host_fuse_pose(_d(pred), _d(est), _d(motion), _d(last), _d(list(np.ravel(last_cov))), _d(list(np.ravel(Q))), f, c)
    return np.array(f), np.array(c).reshape(3, 3)


def estimate_pose(scan_xy, tgt_xyzw, init_deg):
    """PoseEstimator::setScanPair + estimatePose on the GPU. Returns (cost, est (x, y, th_deg), cov, ndt_result)."""
    L = load()
    scan_xy = np.ascontiguousarray(scan_xy, np.float64); tgt_xyzw = np.ascontiguousarray(tgt_xyzw, np.float32)
    est = (C.c_double * 3)(); cov = (C.c_double * 9)(); res = NdtResult()
    cost = L.host_estimate_pose(_p(scan_xy), scan_xy.shape[0], _p(tgt_xyzw), tgt_xyzw.shape[0], _d(init_deg), est, cov, C.byref(res))
    if cost == -1e300:
        raise RuntimeError(L.host_last_error().decode())
    return cost, np.array(est), np.array(cov).reshape(3, 3), res


class Slam:
    """PointCloudMap + FrontEnd + PoseEstimator wired like SlamLauncher::init."""

    def __init__(self):
        self.L = load()
        self.h = C.c_void_p(self.L.host_slam_create())

    def __del__(self):
        try:
            self.L.host_slam_destroy(self.h)
        except Exception:
            pass

    def process(self, sid, odo_deg, xy):
        xy = np.ascontiguousarray(xy, np.float64)
        if self.L.host_slam_process(self.h, sid, _d(list(odo_deg)), _p(xy), xy.shape[0]) != 0:
            raise RuntimeError(self.L.host_last_error().decode())

    def process_forced(self, sid, odo_deg, xy, pose_deg, cov):
        """Teacher forcing: match this scan, then continue from the given pose / covariance (the reference's)."""
        xy = np.ascontiguousarray(xy, np.float64)
        if self.L.host_slam_process_forced(self.h, sid, _d(list(odo_deg)), _p(xy), xy.shape[0], _d(list(pose_deg)),
                                           _d(list(np.asarray(cov, dtype=np.float64).ravel()))) != 0:
            raise RuntimeError(self.L.host_last_error().decode())

    def covs(self):
        n = self.L.host_slam_covs(self.h, None, 0)
        out = np.zeros((n, 3, 3)); self.L.host_slam_covs(self.h, _p(out), n)
        return out

    def loops(self):
        """What the LoopDetector did for the last processed scan and the pose-graph totals:
        (rows: cur node, ref node, rel x, rel y, rel th_deg, cost, accepted, iters, evals, converged), {nodes, arcs, loop_arcs}"""
        f = self.L.host_slam_loops
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]; f.restype = C.c_int64
        rows = np.zeros((64, 10)); cnt = np.zeros(3, np.int64)
        n = f(self.h, _p(rows), 64, _p(cnt))
        return rows[:n].copy(), dict(nodes=int(cnt[0]), arcs=int(cnt[1]), loop_arcs=int(cnt[2]))

    def loop_arcs(self):
        f = self.L.host_slam_loop_arcs
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]; f.restype = C.c_int64
        n = f(self.h, None, 0)
        rows = np.zeros((max(n, 1), 6)); f(self.h, _p(rows), n)
        return rows[:n]

    def nodes(self):
        f = self.L.host_slam_nodes
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]; f.restype = C.c_int64
        n = f(self.h, None, 0)
        rows = np.zeros((max(n, 1), 3)); f(self.h, _p(rows), n)
        return rows[:n]

    def poses(self):
        n = self.L.host_slam_poses(self.h, None, 0)
        out = np.zeros((n, 3)); self.L.host_slam_poses(self.h, _p(out), n)
        return out

    def local_map(self):
        n = self.L.host_slam_local_map(self.h, None, 0)
        out = np.zeros((n, 4), np.float32); self.L.host_slam_local_map(self.h, _p(out), n)
        return out

    def global_map(self):
        n = self.L.host_slam_global_map(self.h, None, 0)
        out = np.zeros((n, 4), np.float32); self.L.host_slam_global_map(self.h, _p(out), n)
        return out

    def submaps(self):
        return self.L.host_slam_submaps(self.h)

    def stats(self):
        ms = (C.c_double * 10)(); cnt = (C.c_int64 * 3)()
        self.L.host_slam_stats(self.h, ms, cnt)
        return dict(resample_ms=ms[0], estimate_ms=ms[1], fuse_ms=ms[2], growmap_ms=ms[3], device_grid_ms=ms[4],
                    device_match_ms=ms[5], host_filter_ms=ms[6], set_source_wall_ms=ms[7], set_target_wall_ms=ms[8],
                    align_wall_ms=ms[9], matches=cnt[0], evals=cnt[1], point_evals=cnt[2])


def launcher_run() -> int:
    L = load()
    n = L.host_launcher_run()
    if n < 0:
        raise RuntimeError(L.host_last_error().decode())
    return n


def write_scan_log(path, odo_deg, scans, header_lines=("# synthetic", "# ndt_slam text log", "#", "#"), left=None, right=None):
    """The reference's text scan log (SlamLauncher.cpp:37-105): 4 header lines, then per
    scan 'stamp x y theta_deg image' and three point groups 'count x y x y ...' (front, left, right)."""
    def group(xy):
        return f"{xy.shape[0]} " + " ".join(f"{float(x)!r} {float(y)!r}" for x, y in xy) + (" " if xy.shape[0] else "")
    empty = np.zeros((0, 2))
    with open(path, "w") as f:
        for h in header_lines:
            f.write(h + "\n")
        for i, (o, xy) in enumerate(zip(odo_deg, scans)):
            f.write(f"{i} {float(o[0])!r} {float(o[1])!r} {float(o[2])!r} img{i}.png\n")
            f.write(group(xy) + "\n")
            lg = group(left[i] if left is not None else empty)
            rg = group(right[i] if right is not None else empty)
            # The file must end right after the last token: the reader detects the end of data by hitting EOF inside
            # the last getline (SlamLauncher.cpp:96-99), so the record that reaches EOF is parsed but not processed.
            f.write(lg + "\n" + (rg + "\n" if i + 1 < len(scans) else rg.rstrip(" ")))


def launcher_parse(path, sidelidar=True, cap_scans=4096, cap_points=4_000_000):
    """host's SlamLauncher::readFormat + input_file_line over a text log -> (meta (n, 5): sid x y th n_points, points (m, 2))."""
    L = load()
    set_params(filename_in=str(path), poses_name=str(path) + ".poses.tmp", sidelidar="true" if sidelidar else "false")
    meta = np.zeros((cap_scans, 5)); xy = np.zeros((cap_points, 2)); npts = C.c_int64()
    f = L.host_launcher_parse
    f.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]; f.restype = C.c_int64
    n = f(_p(meta), cap_scans, _p(xy), cap_points, C.byref(npts))
    return meta[:n].copy(), xy[: npts.value].copy()


def launcher_write_poses(path, poses_deg):
    """host's SlamLauncher::output_file_poses into `path`."""
    L = load()
    import tempfile
    dummy = tempfile.NamedTemporaryFile(suffix=".log", delete=False); dummy.write(b"#\n#\n#\n#\n"); dummy.close()
    set_params(filename_in=dummy.name, poses_name=str(path))
    poses = np.ascontiguousarray(poses_deg, np.float64)
    f = L.host_launcher_write_poses
    f.argtypes = [C.c_void_p, C.c_int64]; f.restype = None
    f(_p(poses), poses.shape[0])


def save_maps(map_name, separated_name, global_xyzw, submaps_xyzw):
    """host's PointCloudMap::saveGlobalMap: global PCD + one PCD per sub-map."""
    L = load()
    set_params(map_name=str(map_name), separated_map_name=str(separated_name))
    g = np.ascontiguousarray(global_xyzw, np.float32)
    off = np.zeros(len(submaps_xyzw) + 1, np.int64); off[1:] = np.cumsum([s.shape[0] for s in submaps_xyzw])
    sub = np.ascontiguousarray(np.concatenate(submaps_xyzw, axis=0), np.float32) if len(submaps_xyzw) else np.zeros((0, 4), np.float32)
    f = L.host_save_maps
    f.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int]; f.restype = None
    f(_p(g), g.shape[0], _p(sub), _p(off), len(submaps_xyzw))


def loop_candidates(poses_deg, atd):
    """LoopDetector::findCandidates for the last pose against all earlier key frames (host only)."""
    L = load()
    f = L.host_loop_candidates
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64]; f.restype = C.c_int64
    poses = np.ascontiguousarray(poses_deg, np.float64); a = np.ascontiguousarray(atd, np.float64)
    out = np.zeros(256, np.int32)
    n = f(_p(poses), _p(a), poses.shape[0], _p(out), 256)
    return out[:n].copy()


def relocalize(devices, map_xyzw, scan_xyzw, hypotheses, resolution=0.5, want_all=False):
    """Relocalizer (C++ host over the C ABI): returns (best index, best NdtResult, all results or None, slowest shard ms)."""
    from .capi import NdtResult, RESULT_DTYPE
    L = load()
    f = L.host_relocalize
    f.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_double,
                  C.POINTER(NdtResult), C.c_void_p, C.POINTER(C.c_double)]
    f.restype = C.c_int64
    dev = np.ascontiguousarray(devices, np.int32)
    m = np.ascontiguousarray(map_xyzw, np.float32); sc = np.ascontiguousarray(scan_xyzw, np.float32)
    hyp = np.ascontiguousarray(hypotheses, np.float64)
    best = NdtResult(); ms = C.c_double()
    res = np.zeros(hyp.shape[0], RESULT_DTYPE) if want_all else None
    bi = f(_p(dev), dev.shape[0], _p(m), m.shape[0], _p(sc), sc.shape[0], _p(hyp), hyp.shape[0], resolution, C.byref(best),
           _p(res) if want_all else None, C.byref(ms))
    if bi == -2:
        raise RuntimeError(L.host_last_error().decode())
    return int(bi), best, res, ms.value
