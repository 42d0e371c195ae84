"""ctypes binding of the CPU oracle (oracle/libndt_oracle.so). TEST INFRASTRUCTURE ONLY:
import this from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

from ndt_slam_b200.capi import NdtEvalOut, NdtGridInfo, NdtParams, NdtResult

HERE = Path(__file__).resolve().parent
LIB = HERE / "libndt_oracle.so"
_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not LIB.exists():
        from ndt_slam_b200 import build
        build.build_oracle()
    L = C.CDLL(str(LIB))
    vp, i64, dp = C.c_void_p, C.c_int64, C.POINTER(C.c_double)
    L.oracle_create.argtypes = [C.POINTER(NdtParams)]; L.oracle_create.restype = vp
    L.oracle_destroy.argtypes = [vp]
    L.oracle_want_fitness.argtypes = [vp, C.c_int]
    L.oracle_gauss.argtypes = [vp, dp]
    L.oracle_set_target.argtypes = [vp, vp, i64]
    L.oracle_set_source.argtypes = [vp, vp, i64]
    L.oracle_grid_info.argtypes = [vp, C.POINTER(NdtGridInfo)]
    L.oracle_grid_readback.argtypes = [vp, i64, vp, vp, vp, vp, vp]; L.oracle_grid_readback.restype = i64
    L.oracle_cell_index.argtypes = [vp, vp, i64, vp]
    L.oracle_eval.argtypes = [vp, dp, C.c_int, C.POINTER(NdtEvalOut)]
    L.oracle_eval_stats.argtypes = [vp, C.POINTER(i64), dp]
    L.oracle_align.argtypes = [vp, dp, C.POINTER(NdtResult)]
    L.oracle_align_trace.argtypes = [vp, dp, C.POINTER(NdtResult), vp, i64]; L.oracle_align_trace.restype = i64
    L.oracle_fitness.argtypes = [vp, dp]; L.oracle_fitness.restype = C.c_double
    L.oracle_approx_voxel_filter.argtypes = [vp, i64, C.c_float, vp]; L.oracle_approx_voxel_filter.restype = i64
    L.oracle_resample.argtypes = [vp, i64, C.c_double, C.c_double, vp, i64]; L.oracle_resample.restype = i64
    for f in ("oracle_add_angle", "oracle_sub_angle"):
        getattr(L, f).argtypes = [C.c_double, C.c_double]; getattr(L, f).restype = C.c_double
    L.oracle_cal_motion.argtypes = [dp, dp, dp]
    L.oracle_cal_pred_pose.argtypes = [dp, dp, dp]
    L.oracle_odometry_cov.argtypes = [dp, dp, dp, C.c_double, C.c_double, C.c_double, dp]
    L.oracle_fuse_pose.argtypes = [dp, dp, dp, dp, dp, dp, C.c_double, C.c_double, C.c_double, dp, dp]
    _lib = L
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _d(seq):
    return (C.c_double * len(seq))(*seq)


class Oracle:
    def __init__(self, params: NdtParams):
        self.L = load()
        self.params = params
        self.h = C.c_void_p(self.L.oracle_create(C.byref(params)))

    def __del__(self):
        try:
            if self.h:
                self.L.oracle_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def want_fitness(self, on: bool):
        self.L.oracle_want_fitness(self.h, int(on))

    def gauss(self):
        o = (C.c_double * 2)()
        self.L.oracle_gauss(self.h, o)
        return o[0], o[1]

    def set_target(self, xyzw):
        xyzw = np.ascontiguousarray(xyzw, np.float32)
        self.L.oracle_set_target(self.h, _p(xyzw), xyzw.shape[0])

    def set_source(self, xyzw):
        xyzw = np.ascontiguousarray(xyzw, np.float32)
        self.L.oracle_set_source(self.h, _p(xyzw), xyzw.shape[0])

    def grid_info(self) -> NdtGridInfo:
        gi = NdtGridInfo()
        self.L.oracle_grid_info(self.h, C.byref(gi))
        return gi

    def grid_readback(self):
        n = self.grid_info().n_leaves
        idx = np.zeros(n, np.int32); cnt = np.zeros(n, np.int32)
        mean = np.zeros((n, 2)); icov = np.zeros((n, 4)); cen = np.zeros((n, 2), np.float32)
        self.L.oracle_grid_readback(self.h, n, _p(idx), _p(cnt), _p(mean), _p(icov), _p(cen))
        return dict(cell_idx=idx, nr_points=cnt, mean=mean, icov=icov, centroid=cen)

    def cell_index(self, xyzw):
        xyzw = np.ascontiguousarray(xyzw, np.float32)
        out = np.zeros(xyzw.shape[0], np.int32)
        self.L.oracle_cell_index(self.h, _p(xyzw), xyzw.shape[0], _p(out))
        return out

    def eval(self, pose, want_hessian=True) -> NdtEvalOut:
        out = NdtEvalOut()
        self.L.oracle_eval(self.h, _d(list(pose)), int(want_hessian), C.byref(out))
        return out

    def eval_stats(self):
        a = C.c_int64(); g = C.c_double()
        self.L.oracle_eval_stats(self.h, C.byref(a), C.byref(g))
        return a.value, g.value

    def align(self, guess) -> NdtResult:
        r = NdtResult()
        self.L.oracle_align(self.h, _d(list(guess)), C.byref(r))
        return r

    def align_trace(self, guess, cap=4096):
        r = NdtResult()
        tr = np.zeros((cap, 6))
        n = self.L.oracle_align_trace(self.h, _d(list(guess)), C.byref(r), _p(tr), cap)
        return r, tr[: min(n, cap)]

    def fitness(self, pose) -> float:
        return self.L.oracle_fitness(self.h, _d(list(pose)))


def approx_voxel_filter(xyzw, leaf: float):
    L = load()
    xyzw = np.ascontiguousarray(xyzw, np.float32)
    out = np.zeros_like(xyzw)
    m = L.oracle_approx_voxel_filter(_p(xyzw), xyzw.shape[0], leaf, _p(out))
    return np.ascontiguousarray(out[:m])


def resample(xy, space: float, space_thre: float):
    L = load()
    xy = np.ascontiguousarray(xy, np.float64)
    cap = max(4 * xy.shape[0] + 16, 64)
    while True:
        out = np.zeros((cap, 2))
        m = L.oracle_resample(_p(xy), xy.shape[0], space, space_thre, _p(out), cap)
        if m >= 0:
            return np.ascontiguousarray(out[:m])
        cap = -m + 16
