"""ctypes binding of the C ABI in include/ndt_b200.h (libndt_b200.so).

This is plumbing for tests and bench.py: numpy host buffers or raw device pointers (torch tensors'
data_ptr()) go straight to the C entry points. There is no fallback of any kind: if the CUDA
library is missing, importing `load()` raises.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

from . import build as _build

MEM_HOST, MEM_DEVICE = 0, 1
PAIRS_AUTO, PAIRS_WARP, PAIRS_CTA = 0, 1, 2
BLOB_POINTS = 1

QUIRK_COV_INIT_IDENTITY = 1 << 0
QUIRK_COV_SCALE_NM1_N = 1 << 1
QUIRK_MT_INTERVAL_LT0 = 1 << 2
QUIRK_ANGLE_SNAP = 1 << 3
QUIRK_TRANSFORM_SSE_ORDER = 1 << 4
QUIRKS_PCL_1_10 = QUIRK_COV_INIT_IDENTITY | QUIRK_COV_SCALE_NM1_N | QUIRK_MT_INTERVAL_LT0 | QUIRK_ANGLE_SNAP


class NdtParams(C.Structure):
    _fields_ = [("resolution", C.c_float), ("step_size", C.c_double), ("trans_eps", C.c_double),
                ("max_iter", C.c_int32), ("outlier_ratio", C.c_double), ("min_points", C.c_int32),
                ("eig_mult", C.c_double), ("quirks", C.c_int32), ("device", C.c_int32),
                ("stream", C.c_void_p), ("align_skip_fitness", C.c_int32), ("pairs_schedule", C.c_int32),
                ("pairs_batch_points", C.c_int64), ("align_team", C.c_int32), ("reserved0", C.c_int32)]


class NdtEvalOut(C.Structure):
    _fields_ = [("score", C.c_double), ("grad", C.c_double * 3), ("hess", C.c_double * 9),
                ("n_pairs", C.c_int64)]


class NdtResult(C.Structure):
    _fields_ = [("pose", C.c_double * 3), ("T", C.c_float * 16), ("score", C.c_double),
                ("trans_prob", C.c_double), ("fitness", C.c_double), ("hess", C.c_double * 9),
                ("converged", C.c_int32), ("iters", C.c_int32), ("evals", C.c_int32),
                ("passes_run", C.c_int32), ("point_evals", C.c_int64)]


class NdtGridInfo(C.Structure):
    _fields_ = [("min_b", C.c_int32 * 2), ("div_b", C.c_int32 * 2), ("n_points", C.c_int64),
                ("n_leaves", C.c_int32), ("n_slots", C.c_int32), ("n_valid", C.c_int32),
                ("reserved", C.c_int32)]


# numpy view of ndt_result for batched calls (must match the C layout; checked in tests)
RESULT_DTYPE = np.dtype([("pose", "<f8", 3), ("T", "<f4", 16), ("score", "<f8"), ("trans_prob", "<f8"),
                         ("fitness", "<f8"), ("hess", "<f8", 9), ("converged", "<i4"), ("iters", "<i4"),
                         ("evals", "<i4"), ("passes_run", "<i4"), ("point_evals", "<i8")], align=True)

EXPORTS = [
    "ndt_params_default", "ndt_create", "ndt_destroy", "ndt_last_error", "ndt_version",
    "ndt_set_target", "ndt_set_target_prefix", "ndt_set_target_incremental", "ndt_set_target_incremental_async", "ndt_get_grid_info", "ndt_grid_readback", "ndt_cell_index",
    "ndt_set_source", "ndt_approx_voxel_filter", "ndt_eval", "ndt_eval_batch",
    "ndt_align", "ndt_align_batch", "ndt_best_of", "ndt_match_pairs", "ndt_match_pairs_xy",
    "ndt_grid_blob_size", "ndt_grid_export", "ndt_grid_import", "ndt_replicate_grid", "ndt_best_of_multi", "ndt_trim",
    "ndt_alloc", "ndt_free", "ndt_upload", "ndt_download",
    "ndt_launch_count", "ndt_last_kernel_ms", "ndt_synchronize",
]

_lib = None


def lib_path() -> Path:
    """In-tree library; NDT_B200_LIB overrides it for A/B experiments with differently tuned builds."""
    import os
    return Path(os.environ["NDT_B200_LIB"]) if os.environ.get("NDT_B200_LIB") else _build.LIB_CUDA


def load() -> C.CDLL:
    """dlopen libndt_b200.so (no compute call, works without a GPU). Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    p = lib_path()
    if not p.exists():
        raise RuntimeError(f"{p} is not built -- run `python -m ndt_slam_b200.build` (there is no CPU fallback)")
    L = C.CDLL(str(p))
    vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int
    L.ndt_params_default.argtypes = [C.POINTER(NdtParams)]
    L.ndt_create.argtypes = [C.POINTER(NdtParams), C.POINTER(vp)]
    L.ndt_destroy.argtypes = [vp]
    L.ndt_last_error.argtypes = [vp]
    L.ndt_last_error.restype = C.c_char_p
    L.ndt_version.restype = C.c_char_p
    L.ndt_set_target.argtypes = [vp, vp, i64, i32]
    L.ndt_set_target_prefix.argtypes = [vp, vp, i64, i64, i32]
    L.ndt_set_target_incremental.argtypes = [vp, vp, i64, i64, i64, i32]
    L.ndt_set_target_incremental_async.argtypes = [vp, vp, i64, i64, i64, i32]
    L.ndt_get_grid_info.argtypes = [vp, C.POINTER(NdtGridInfo)]
    L.ndt_grid_readback.argtypes = [vp, i64, vp, vp, vp, vp, vp, C.POINTER(i64)]
    L.ndt_cell_index.argtypes = [vp, vp, i64, i32, vp]
    L.ndt_set_source.argtypes = [vp, vp, i64, i32]
    L.ndt_approx_voxel_filter.argtypes = [vp, vp, i64, C.c_float, i32, vp, C.POINTER(i64)]
    L.ndt_eval.argtypes = [vp, C.POINTER(C.c_double), i32, C.POINTER(NdtEvalOut)]
    L.ndt_eval_batch.argtypes = [vp, vp, i64, i32, i32, vp]
    L.ndt_align.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(NdtResult)]
    L.ndt_align_batch.argtypes = [vp, vp, i64, i32, i32, vp]
    L.ndt_best_of.argtypes = [vp, vp, i64, i32, C.POINTER(i64), C.POINTER(NdtResult)]
    L.ndt_match_pairs.argtypes = [vp, vp, vp, vp, vp, vp, i64, C.c_float, i32, vp]
    L.ndt_match_pairs_xy.argtypes = [vp, vp, vp, vp, vp, vp, i64, C.c_float, i32, vp]
    L.ndt_grid_blob_size.argtypes = [vp, i32, C.POINTER(i64)]
    L.ndt_grid_export.argtypes = [vp, i32, vp, i64]
    L.ndt_grid_import.argtypes = [vp, vp, i64]
    L.ndt_replicate_grid.argtypes = [C.POINTER(vp), i32, i32]
    L.ndt_best_of_multi.argtypes = [C.POINTER(vp), C.POINTER(vp), C.POINTER(i64), i32, C.POINTER(i32), C.POINTER(i64), C.POINTER(NdtResult)]
    L.ndt_trim.argtypes = [vp]
    L.ndt_alloc.argtypes = [vp, i64, C.POINTER(vp)]
    L.ndt_free.argtypes = [vp, vp]
    L.ndt_upload.argtypes = [vp, vp, vp, i64]
    L.ndt_download.argtypes = [vp, vp, vp, i64]
    L.ndt_launch_count.argtypes = [vp, C.POINTER(i64)]
    L.ndt_last_kernel_ms.argtypes = [vp, C.POINTER(C.c_float)]
    L.ndt_synchronize.argtypes = [vp]
    for name in EXPORTS:
        if name not in ("ndt_last_error", "ndt_version"):
            getattr(L, name).restype = C.c_int
    _lib = L
    return L


class NdtError(RuntimeError):
    pass


def _ptr(a):
    """numpy array -> void*, int -> device pointer passthrough."""
    if a is None:
        return None
    if isinstance(a, (int,)):
        return C.c_void_p(a)
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


def default_params(**kw) -> NdtParams:
    p = NdtParams()
    load().ndt_params_default(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


class Ndt:
    """Thin object wrapper over one ndt_handle."""

    def __init__(self, params: NdtParams | None = None, **kw):
        self.L = load()
        self.params = params if params is not None else default_params(**kw)
        self.h = C.c_void_p()
        rc = self.L.ndt_create(C.byref(self.params), C.byref(self.h))
        if rc != 0:
            msg = self.L.ndt_last_error(None)
            raise NdtError(f"ndt_create failed ({rc}): {msg.decode() if msg else ''}")

    def close(self):
        if self.h:
            self.L.ndt_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            msg = self.L.ndt_last_error(self.h)
            raise NdtError(f"ndt call failed ({rc}): {msg.decode() if msg else ''}")

    # -- grid --
    def set_target(self, xyzw, n=None, space=MEM_HOST, n_same=0, n_stable=None, queued=False):
        """n_same > 0: the first n_same points equal the previous target's (only the rest is uploaded).
        n_stable: the first n_stable points will stay a prefix of later targets (incremental grid update on the device);
        queued=True: return once the update is launched (ndt_set_target_incremental_async), the next call waits for it."""
        if n is None:
            n = xyzw.shape[0]
        if n_stable is not None:
            fn = self.L.ndt_set_target_incremental_async if queued else self.L.ndt_set_target_incremental
            self._ck(fn(self.h, _ptr(xyzw), n, n_same, n_stable, space))
        elif n_same:
            self._ck(self.L.ndt_set_target_prefix(self.h, _ptr(xyzw), n, n_same, space))
        else:
            self._ck(self.L.ndt_set_target(self.h, _ptr(xyzw), n, space))

    def grid_info(self) -> NdtGridInfo:
        gi = NdtGridInfo()
        self._ck(self.L.ndt_get_grid_info(self.h, C.byref(gi)))
        return gi

    def grid_readback(self):
        gi = self.grid_info()
        n = gi.n_leaves
        idx = np.zeros(n, np.int32); cnt = np.zeros(n, np.int32)
        mean = np.zeros((n, 2)); icov = np.zeros((n, 4)); cen = np.zeros((n, 2), np.float32)
        nout = C.c_int64()
        self._ck(self.L.ndt_grid_readback(self.h, n, _ptr(idx), _ptr(cnt), _ptr(mean), _ptr(icov), _ptr(cen), C.byref(nout)))
        assert nout.value == n
        return dict(cell_idx=idx, nr_points=cnt, mean=mean, icov=icov, centroid=cen)

    def cell_index(self, xyzw, n=None, space=MEM_HOST):
        if n is None:
            n = xyzw.shape[0]
        out = np.zeros(n, np.int32)
        self._ck(self.L.ndt_cell_index(self.h, _ptr(xyzw), n, space, _ptr(out)))
        return out

    # -- source --
    def set_source(self, xyzw, n=None, space=MEM_HOST):
        if n is None:
            n = xyzw.shape[0]
        self._ck(self.L.ndt_set_source(self.h, _ptr(xyzw), n, space))

    def approx_voxel_filter(self, xyzw, leaf: float):
        n = xyzw.shape[0]
        out = np.zeros((n, 4), np.float32)
        nout = C.c_int64()
        self._ck(self.L.ndt_approx_voxel_filter(self.h, _ptr(xyzw), n, leaf, MEM_HOST, _ptr(out), C.byref(nout)))
        return np.ascontiguousarray(out[: nout.value])

    # -- objective --
    def eval(self, pose, want_hessian=True) -> NdtEvalOut:
        p = (C.c_double * 3)(*pose)
        out = NdtEvalOut()
        self._ck(self.L.ndt_eval(self.h, p, int(want_hessian), C.byref(out)))
        return out

    def eval_batch(self, poses, n=None, want_hessian=True, space=MEM_HOST, out=None):
        if n is None:
            n = poses.shape[0]
        if out is None:
            out = np.zeros((n, 14))
        self._ck(self.L.ndt_eval_batch(self.h, _ptr(poses), n, int(want_hessian), space, _ptr(out)))
        return out

    # -- matching --
    def align(self, guess) -> NdtResult:
        g = (C.c_double * 3)(*guess)
        r = NdtResult()
        self._ck(self.L.ndt_align(self.h, g, C.byref(r)))
        return r

    def align_batch(self, guesses, n=None, space=MEM_HOST, out=None, want_fitness=None):
        """want_fitness=None: small batches (n < 64) carry the fitness score, large ones are ranked by score only."""
        if n is None:
            n = guesses.shape[0]
        if out is None:
            out = np.zeros(n, RESULT_DTYPE)
        if want_fitness is None:
            want_fitness = n < 64
        self._ck(self.L.ndt_align_batch(self.h, _ptr(guesses), n, space, int(bool(want_fitness)), _ptr(out)))
        return out

    def best_of(self, results, n=None, space=MEM_HOST):
        if n is None:
            n = results.shape[0]
        bi = C.c_int64(); best = NdtResult()
        self._ck(self.L.ndt_best_of(self.h, _ptr(results), n, space, C.byref(bi), C.byref(best)))
        return bi.value, best

    def match_pairs(self, src, src_off, tgt, tgt_off, guesses, n_pairs, source_leaf=0.0, space=MEM_HOST, out=None, xy=False):
        """xy=True: src / tgt hold (x, y) float pairs, 8 bytes per point (ndt_match_pairs_xy)."""
        if out is None:
            out = np.zeros(n_pairs, RESULT_DTYPE)
        fn = self.L.ndt_match_pairs_xy if xy else self.L.ndt_match_pairs
        self._ck(fn(self.h, _ptr(src), _ptr(src_off), _ptr(tgt), _ptr(tgt_off), _ptr(guesses), n_pairs, source_leaf, space, _ptr(out)))
        return out

    # -- replication --
    def grid_blob_size(self, flags=BLOB_POINTS) -> int:
        b = C.c_int64()
        self._ck(self.L.ndt_grid_blob_size(self.h, flags, C.byref(b)))
        return b.value

    def grid_export(self, dev_ptr: int, nbytes: int, flags=BLOB_POINTS):
        self._ck(self.L.ndt_grid_export(self.h, flags, C.c_void_p(dev_ptr), nbytes))

    def trim(self):
        self._ck(self.L.ndt_trim(self.h))

    def grid_import(self, dev_ptr: int, nbytes: int):
        self._ck(self.L.ndt_grid_import(self.h, C.c_void_p(dev_ptr), nbytes))

    # -- instrumentation --
    def launch_count(self) -> int:
        n = C.c_int64()
        self._ck(self.L.ndt_launch_count(self.h, C.byref(n)))
        return n.value

    def last_kernel_ms(self) -> float:
        ms = C.c_float()
        self._ck(self.L.ndt_last_kernel_ms(self.h, C.byref(ms)))
        return ms.value

    def synchronize(self):
        self._ck(self.L.ndt_synchronize(self.h))


def replicate_grid(handles, flags=BLOB_POINTS):
    """ndt_replicate_grid: copy the grid of handles[0] to the other handles (peer copies over NVLink)."""
    L = load()
    arr = (C.c_void_p * len(handles))(*[h.h for h in handles])
    rc = L.ndt_replicate_grid(arr, len(handles), flags)
    if rc != 0:
        msg = L.ndt_last_error(handles[0].h)
        raise NdtError(f"ndt_replicate_grid failed ({rc}): {msg.decode() if msg else ''}")


def best_of_multi(handles, device_ptrs, counts):
    """ndt_best_of_multi: (handle index, result index, NdtResult) of the best converged result over all shards."""
    L = load()
    n = len(handles)
    hs = (C.c_void_p * n)(*[h.h for h in handles])
    ps = (C.c_void_p * n)(*[C.c_void_p(p) for p in device_ptrs])
    cs = (C.c_int64 * n)(*counts)
    bh, bi, best = C.c_int32(), C.c_int64(), NdtResult()
    rc = L.ndt_best_of_multi(hs, ps, cs, n, C.byref(bh), C.byref(bi), C.byref(best))
    if rc != 0:
        raise NdtError(f"ndt_best_of_multi failed ({rc})")
    return bh.value, bi.value, best
