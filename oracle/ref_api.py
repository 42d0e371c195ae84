"""ctypes binding of oracle/_ref/libndt_slam_ref.so: the reference's own sources + the restated
mini-PCL (TEST INFRASTRUCTURE). Available only where the library was built (it needs /root/reference
at build time; the built .so travels to the GPU box)."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

LIB = Path(__file__).resolve().parent / "_ref" / "libndt_slam_ref.so"
_lib = None

# ndt_mapping.launch values; Resolution per BASELINE configs
LAUNCH_PARAMS = {
    "space": "0.05", "space_thre": "0.25", "LeafSize": "0.05", "TransformationEpsilon": "0.01", "StepSize": "0.1",
    "Resolution": "0.5", "MaximumIterations": "35", "coeNDTCov": "1.0", "score_thre": "0.5", "sepThre": "10.0",
    "removeMoving": "false", "resol": "0.05", "thre_neighbor": "0.2", "delTime": "0.5", "coeVel": "0.1",
    "coeOmega": "0.5", "keyframe_skip": "5", "start_frame": "0",
}


def available() -> bool:
    return LIB.exists()


def load():
    global _lib
    if _lib is not None:
        return _lib
    L = C.CDLL(str(LIB))
    vp, i64, dp, d = C.c_void_p, C.c_int64, C.POINTER(C.c_double), C.c_double
    L.ref_param_set.argtypes = [C.c_char_p, C.c_char_p]
    L.ref_resample.argtypes = [vp, i64, vp, i64]; L.ref_resample.restype = i64
    L.ref_add_angle.argtypes = [d, d]; L.ref_add_angle.restype = d
    L.ref_sub_angle.argtypes = [d, d]; L.ref_sub_angle.restype = d
    L.ref_cal_motion.argtypes = [dp, dp, dp]
    L.ref_cal_pred_pose.argtypes = [dp, dp, dp]
    L.ref_odometry_cov.argtypes = [dp, dp, dp, dp]
    L.ref_fuse_pose.argtypes = [dp] * 8
    L.ref_estimate_pose.argtypes = [vp, i64, vp, i64, dp, dp, dp]; L.ref_estimate_pose.restype = d
    L.ref_ndt_create.argtypes = [C.c_float, d, d, C.c_int]; L.ref_ndt_create.restype = vp
    L.ref_ndt_destroy.argtypes = [vp]
    L.ref_ndt_set_target.argtypes = [vp, vp, i64]
    L.ref_ndt_set_source.argtypes = [vp, vp, i64]
    L.ref_ndt_grid.argtypes = [vp, i64, vp, vp, vp, vp, vp, vp]; L.ref_ndt_grid.restype = i64
    L.ref_ndt_eval.argtypes = [vp, dp, C.c_int, dp]; L.ref_ndt_eval.restype = d
    L.ref_ndt_align.argtypes = [vp, dp, dp]
    L.ref_voxel_filter.argtypes = [vp, i64, C.c_float, vp]; L.ref_voxel_filter.restype = i64
    L.ref_slam_create.restype = vp
    L.ref_slam_destroy.argtypes = [vp]
    L.ref_slam_process.argtypes = [vp, C.c_int, dp, vp, i64]
    L.ref_slam_poses.argtypes = [vp, vp, i64]; L.ref_slam_poses.restype = i64
    L.ref_slam_local_map.argtypes = [vp, vp, i64]; L.ref_slam_local_map.restype = i64
    L.ref_slam_global_map.argtypes = [vp, vp, i64]; L.ref_slam_global_map.restype = i64
    L.ref_slam_submaps.argtypes = [vp]; L.ref_slam_submaps.restype = C.c_int
    L.ref_map_replay.argtypes = [vp, vp, vp, C.c_int, vp, i64, vp, vp, i64, vp]; L.ref_map_replay.restype = i64
    _lib = L
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _d(v):
    return (C.c_double * len(v))(*v)


def set_params(**kw):
    """ROS parameter server stand-in: the reference classes read it in their constructors."""
    L = load()
    L.ref_param_clear()
    p = dict(LAUNCH_PARAMS)
    p.update({k: str(v) for k, v in kw.items()})
    for k, v in p.items():
        L.ref_param_set(k.encode(), v.encode())


def resample(xy):
    L = load()
    xy = np.ascontiguousarray(xy, np.float64)
    out = np.zeros((4 * xy.shape[0] + 16, 2))
    m = L.ref_resample(_p(xy), xy.shape[0], _p(out), out.shape[0])
    assert m >= 0
    return np.ascontiguousarray(out[:m])


def voxel_filter(xyzw, leaf):
    L = load()
    xyzw = np.ascontiguousarray(xyzw, np.float32)
    out = np.zeros_like(xyzw)
    m = L.ref_voxel_filter(_p(xyzw), xyzw.shape[0], leaf, _p(out))
    return np.ascontiguousarray(out[:m])


def pcfilter(base_xyzw, test_xyzw):
    """The reference's PCFilter: (difference_extraction(base, test), remove_neighborPoint(test, diff))."""
    L = load()
    base = np.ascontiguousarray(base_xyzw, np.float32); test = np.ascontiguousarray(test_xyzw, np.float32)
    d, k = np.zeros_like(test), np.zeros_like(test)
    nd, nk = C.c_int64(), C.c_int64()
    L.ref_pcfilter.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.POINTER(C.c_int64), C.c_void_p, C.POINTER(C.c_int64)]
    L.ref_pcfilter.restype = None
    L.ref_pcfilter(_p(base), base.shape[0], _p(test), test.shape[0], _p(d), C.byref(nd), _p(k), C.byref(nk))
    return np.ascontiguousarray(d[: nd.value]), np.ascontiguousarray(k[: nk.value])


def map_replay(poses_deg, scans_map_xy):
    """The reference's PointCloudMap alone (same driver as ndt_slam_b200.host_api.map_replay)."""
    L = load()
    poses = np.ascontiguousarray(poses_deg, np.float64)
    off = np.zeros(len(scans_map_xy) + 1, np.int64)
    off[1:] = np.cumsum([s.shape[0] for s in scans_map_xy])
    xy = np.ascontiguousarray(np.concatenate(scans_map_xy, axis=0), np.float64)
    cap = int(off[-1]) + 1024
    lo, go = np.zeros((cap, 4), np.float32), np.zeros((cap, 4), np.float32)
    nl, ng = C.c_int64(), C.c_int64()
    ns = L.ref_map_replay(_p(poses), _p(xy), _p(off), len(scans_map_xy), _p(lo), cap, C.byref(nl), _p(go), cap, C.byref(ng))
    return int(ns), np.ascontiguousarray(lo[: nl.value]), np.ascontiguousarray(go[: ng.value])


def fuse_pose(pred, est, motion, last, last_cov, Q):
    L = load()
    fused = (C.c_double * 3)(); cov = (C.c_double * 9)()
    L.ref_fuse_pose(_d(pred), _d(est), _d(motion), _d(last), _d(list(np.ravel(last_cov))), _d(list(np.ravel(Q))), fused, cov)
    return np.array(fused), np.array(cov).reshape(3, 3)


def odometry_cov(motion, last, last_cov):
    L = load()
    cov = (C.c_double * 9)()
    L.ref_odometry_cov(_d(motion), _d(last), _d(list(np.ravel(last_cov))), cov)
    return np.array(cov).reshape(3, 3)


def cal_motion(cur, prev):
    L = load(); o = (C.c_double * 3)(); L.ref_cal_motion(_d(cur), _d(prev), o); return np.array(o)


def cal_pred_pose(motion, last):
    L = load(); o = (C.c_double * 3)(); L.ref_cal_pred_pose(_d(motion), _d(last), o); return np.array(o)


def estimate_pose(scan_xy, tgt_xyzw, init_deg):
    """PoseEstimator::setScanPair + estimatePose (reference code). init/est are (x, y, th_deg)."""
    L = load()
    scan_xy = np.ascontiguousarray(scan_xy, np.float64); tgt_xyzw = np.ascontiguousarray(tgt_xyzw, np.float32)
    est = (C.c_double * 3)(); cov = (C.c_double * 9)()
    cost = L.ref_estimate_pose(_p(scan_xy), scan_xy.shape[0], _p(tgt_xyzw), tgt_xyzw.shape[0], _d(init_deg), est, cov)
    return cost, np.array(est), np.array(cov).reshape(3, 3)


class RefNdt:
    """The restated 6-DoF pcl::NormalDistributionsTransform object."""

    def __init__(self, resolution=0.5, step=0.1, eps=0.01, max_iter=35):
        self.L = load()
        self.h = C.c_void_p(self.L.ref_ndt_create(resolution, step, eps, max_iter))

    def __del__(self):
        try:
            self.L.ref_ndt_destroy(self.h)
        except Exception:
            pass

    def set_target(self, xyzw):
        xyzw = np.ascontiguousarray(xyzw, np.float32); self.L.ref_ndt_set_target(self.h, _p(xyzw), xyzw.shape[0])

    def set_source(self, xyzw):
        xyzw = np.ascontiguousarray(xyzw, np.float32); self.L.ref_ndt_set_source(self.h, _p(xyzw), xyzw.shape[0])

    def grid(self):
        dims = np.zeros(4, np.int32)
        n = self.L.ref_ndt_grid(self.h, 0, None, None, None, None, None, _p(dims))
        cell = np.zeros(n, np.int32); nr = np.zeros(n, np.int32); mean = np.zeros((n, 2)); icov = np.zeros((n, 4)); cen = np.zeros((n, 2), np.float32)
        self.L.ref_ndt_grid(self.h, n, _p(cell), _p(nr), _p(mean), _p(icov), _p(cen), _p(dims))
        return dict(cell_idx=cell, nr_points=nr, mean=mean, icov=icov, centroid=cen, min_b=dims[:2].copy(), div_b=dims[2:].copy())

    def eval(self, pose, want_hessian=True):
        out = (C.c_double * 13)()
        off = self.L.ref_ndt_eval(self.h, _d(list(pose)), int(want_hessian), out)
        o = np.array(out)
        return dict(score=o[0], grad=o[1:4], hess=o[4:13], off_block=off)

    def align(self, guess, want_fitness=True):
        out = (C.c_double * 17)()
        if want_fitness:
            self.L.ref_ndt_align(self.h, _d(list(guess)), out)
        else:
            self.L.ref_ndt_align_nofit.argtypes = self.L.ref_ndt_align.argtypes
            self.L.ref_ndt_align_nofit(self.h, _d(list(guess)), out)
        o = np.array(out)
        return dict(pose=o[0:3], score=o[3], iters=int(o[4]), converged=int(o[5]), evals=int(o[6]), fitness=o[7], hess=o[8:17])


class RefSlam:
    """SlamLauncher::init wiring (PointCloudMap + FrontEnd + PoseEstimator) from the reference sources."""

    def __init__(self):
        self.L = load()
        self.h = C.c_void_p(self.L.ref_slam_create())

    def __del__(self):
        try:
            self.L.ref_slam_destroy(self.h)
        except Exception:
            pass

    def process(self, sid, odo_deg, xy):
        xy = np.ascontiguousarray(xy, np.float64)
        self.L.ref_slam_process(self.h, sid, _d(list(odo_deg)), _p(xy), xy.shape[0])

    def poses(self):
        n = self.L.ref_slam_poses(self.h, None, 0)
        out = np.zeros((n, 3))
        self.L.ref_slam_poses(self.h, _p(out), n)
        return out

    def covs(self):
        self.L.ref_slam_covs.restype = C.c_int64
        self.L.ref_slam_covs.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
        n = self.L.ref_slam_covs(self.h, None, 0)
        out = np.zeros((n, 3, 3))
        self.L.ref_slam_covs(self.h, _p(out), n)
        return out

    def local_map(self):
        n = self.L.ref_slam_local_map(self.h, None, 0)
        out = np.zeros((n, 4), np.float32)
        self.L.ref_slam_local_map(self.h, _p(out), n)
        return out

    def global_map(self):
        n = self.L.ref_slam_global_map(self.h, None, 0)
        out = np.zeros((n, 4), np.float32)
        self.L.ref_slam_global_map(self.h, _p(out), n)
        return out

    def submaps(self):
        return self.L.ref_slam_submaps(self.h)


def launcher_parse(path, sidelidar=True, cap_scans=4096, cap_points=4_000_000):
    """the reference's SlamLauncher::readFormat + input_file_line over a text log -> (meta (n, 5): sid x y th n_points, points (m, 2))."""
    L = load()
    set_params(filename_in=str(path), poses_name=str(path) + ".poses.tmp", sidelidar="true" if sidelidar else "false")
    meta = np.zeros((cap_scans, 5)); xy = np.zeros((cap_points, 2)); npts = C.c_int64()
    f = L.ref_launcher_parse
    f.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]; f.restype = C.c_int64
    n = f(_p(meta), cap_scans, _p(xy), cap_points, C.byref(npts))
    return meta[:n].copy(), xy[: npts.value].copy()


def launcher_write_poses(path, poses_deg):
    """the reference's SlamLauncher::output_file_poses into `path`."""
    L = load()
    import tempfile
    dummy = tempfile.NamedTemporaryFile(suffix=".log", delete=False); dummy.write(b"#\n#\n#\n#\n"); dummy.close()
    set_params(filename_in=dummy.name, poses_name=str(path))
    poses = np.ascontiguousarray(poses_deg, np.float64)
    f = L.ref_launcher_write_poses
    f.argtypes = [C.c_void_p, C.c_int64]; f.restype = None
    f(_p(poses), poses.shape[0])


def save_maps(map_name, separated_name, global_xyzw, submaps_xyzw):
    """the reference's PointCloudMap::saveGlobalMap: global PCD + one PCD per sub-map."""
    L = load()
    set_params(map_name=str(map_name), separated_map_name=str(separated_name))
    g = np.ascontiguousarray(global_xyzw, np.float32)
    off = np.zeros(len(submaps_xyzw) + 1, np.int64); off[1:] = np.cumsum([s.shape[0] for s in submaps_xyzw])
    sub = np.ascontiguousarray(np.concatenate(submaps_xyzw, axis=0), np.float32) if len(submaps_xyzw) else np.zeros((0, 4), np.float32)
    f = L.ref_save_maps
    f.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int]; f.restype = None
    f(_p(g), g.shape[0], _p(sub), _p(off), len(submaps_xyzw))
