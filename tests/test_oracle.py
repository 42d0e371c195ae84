"""CPU tests of the oracle itself: known answers, finite differences, ground-truth recovery.
The reference pins nothing here (no tests, no golden vectors: SURVEY.md 4), so these are the pins."""
import ctypes as C

import numpy as np
import pytest

from ndt_slam_b200 import capi, synth
from oracle import oracle_api as oa
import ndt_common as common

ORACLE_DEBUG_DOUBLE_TRANSFORM = 1 << 16


@pytest.mark.parametrize("res,d1,d2", [
    (1.0, -2.217225244042889, 0.433123004703555),
    (0.5, -0.704446735813879, 0.756362730327364),
    (0.3, -0.199595759306291, 0.924143615677333),
    (0.1, -0.008148528926693, 0.996798231013647),
])
def test_gauss_constants_known_answers(res, d1, d2):
    """SURVEY.md App. A.3 table (PCL computeTransformation, reached from PoseEstimator.cpp:28)."""
    o = oa.Oracle(common.params(resolution=res))
    g1, g2 = o.gauss()
    assert g1 == pytest.approx(d1, rel=1e-12)
    assert g2 == pytest.approx(d2, rel=1e-12)


def test_angles_wrap():
    L = oa.load()
    assert L.oracle_add_angle(170.0, 20.0) == -170.0
    assert L.oracle_sub_angle(-170.0, 20.0) == 170.0
    assert L.oracle_add_angle(90.0, 90.0) == -180.0      # [-180, 180)
    assert L.oracle_sub_angle(10.0, 5.0) == 5.0


def test_resampler_straight_wall():
    """Points every 1 cm along a wall -> exactly `space` apart (ScanPointResampler.cpp:41-62)."""
    xy = np.stack([np.arange(0, 2.0, 0.01), np.full(200, 1.0)], axis=1)
    r = oa.resample(xy, 0.05, 0.25)
    d = np.hypot(*(r[1:] - r[:-1]).T)
    assert np.allclose(d, 0.05, atol=1e-12)
    assert r.shape[0] == 40
    # a gap wider than space_thre is kept as a raw point, not interpolated
    xy2 = np.array([[0.0, 0.0], [1.0, 0.0], [1.01, 0.0]])
    r2 = oa.resample(xy2, 0.05, 0.25)
    assert np.allclose(r2[:2], [[0, 0], [1, 0]])
    assert oa.resample(np.zeros((0, 2)), 0.05, 0.25).shape[0] == 0


def test_voxel_filter_basic():
    """ApproximateVoxelGrid: points of one voxel collapse to their float centroid; output order =
    flush order (collisions first, then hash order)."""
    pts = np.zeros((4, 4), np.float32)
    pts[:, 0] = [0.01, 0.02, 0.03, 0.26]
    pts[:, 1] = [0.01, 0.01, 0.02, 0.01]
    out = oa.approx_voxel_filter(pts, 0.05)
    assert out.shape[0] == 2
    got = sorted(map(tuple, np.round(out[:, :2], 6)))
    assert got[0] == pytest.approx((0.02, 0.013333), abs=1e-5)
    assert got[1] == pytest.approx((0.26, 0.01), abs=1e-6)
    assert oa.approx_voxel_filter(np.zeros((0, 4), np.float32), 0.05).shape[0] == 0


def _one_cell_problem():
    rng = synth.rng_for(7)
    # 40 target points in one 0.5 m cell: anisotropic blob
    t = np.stack([0.25 + rng.normal(0, 0.08, 40), 0.25 + rng.normal(0, 0.02, 40)], axis=1)
    t = np.clip(t, 0.01, 0.49)
    s = np.stack([0.25 + rng.normal(0, 0.05, 12), 0.25 + rng.normal(0, 0.05, 12)], axis=1)
    return synth.to_xyzw(t), synth.to_xyzw(s)


def test_one_cell_hand_computation():
    """One cell, known answer: score/gradient recomputed in numpy from the read-back mean/icov."""
    tgt, src = _one_cell_problem()
    o = oa.Oracle(common.params(resolution=0.5))
    o.set_target(tgt); o.set_source(src)
    gi = o.grid_info()
    assert gi.n_leaves == 1 and gi.n_slots == 1 and gi.n_valid == 1
    g = o.grid_readback()
    assert g["nr_points"][0] == 40
    d1, d2 = o.gauss()
    pose = np.array([0.01, -0.02, 0.03])
    e = o.eval(pose)
    yaw = np.float32(pose[2])
    c, s = np.float32(np.cos(np.float64(yaw))), np.float32(np.sin(np.float64(yaw)))
    x, y = src[:, 0], src[:, 1]
    xt = (c * x + (-s) * y) + np.float32(pose[0])
    yt = (s * x + c * y) + np.float32(pose[1])
    assert xt.dtype == np.float32
    dd = np.stack([xt.astype(np.float64) - g["mean"][0, 0], yt.astype(np.float64) - g["mean"][0, 1]], axis=1)
    Cm = g["icov"][0].reshape(2, 2)
    q = np.einsum("ni,ij,nj->n", dd, Cm, dd)
    ex = np.exp(-d2 * q / 2)
    assert e.n_pairs == 12
    assert e.score == pytest.approx(float(np.sum(-d1 * ex)), rel=1e-12)
    gx = np.sum(d1 * d2 * ex * (dd @ Cm[:, 0]))
    assert e.grad[0] == pytest.approx(float(gx), rel=1e-10)
    # covariance of the cell: biased + I/n, then (n-1)/n  (PCL 1.10 quirks)
    t64 = tgt[:, :2].astype(np.float64)
    n = 40
    cov = (np.cov(t64.T, bias=True) + np.eye(2) / n) * (n - 1) / n
    assert np.allclose(np.linalg.inv(cov), Cm, rtol=1e-7)


@pytest.mark.parametrize("quirks", [capi.QUIRKS_PCL_1_10, capi.QUIRK_MT_INTERVAL_LT0 | capi.QUIRK_ANGLE_SNAP])
def test_gradient_and_hessian_match_finite_differences(quirks):
    """Analytic gradient / Hessian vs central differences (fp64 transform switch, one cell so the
    neighbour set cannot change)."""
    tgt, src = _one_cell_problem()
    o = oa.Oracle(common.params(resolution=0.5, quirks=quirks | ORACLE_DEBUG_DOUBLE_TRANSFORM))
    o.set_target(tgt); o.set_source(src)
    pose = np.array([0.013, -0.021, 0.034])
    e = o.eval(pose)
    h = 1e-5
    g_fd = np.zeros(3); H_fd = np.zeros((3, 3))
    for k in range(3):
        pp, pm = pose.copy(), pose.copy()
        pp[k] += h; pm[k] -= h
        ep, em = o.eval(pp), o.eval(pm)
        assert ep.n_pairs == e.n_pairs == em.n_pairs
        g_fd[k] = (ep.score - em.score) / (2 * h)
        H_fd[:, k] = (np.array(ep.grad) - np.array(em.grad)) / (2 * h)
    assert common.rel_err(e.grad, g_fd) < 1e-7
    assert common.rel_err(np.array(e.hess).reshape(3, 3), H_fd) < 1e-7


def test_grid_indexing_is_float32():
    """Cell assignment follows float32 floor(x * inv_leaf) - min_b, not double floor(x / leaf)."""
    rng = synth.rng_for(11)
    pts = rng.uniform(-200, 200, size=(200_000, 2))
    xyzw = synth.to_xyzw(pts)
    o = oa.Oracle(common.params(resolution=0.1))
    o.set_target(xyzw)
    gi = o.grid_info()
    idx = o.cell_index(xyzw)
    inv = np.float32(1.0) / np.float32(0.1)
    i0 = (np.floor(xyzw[:, 0] * inv) - np.float32(gi.min_b[0])).astype(np.int64)
    i1 = (np.floor(xyzw[:, 1] * inv) - np.float32(gi.min_b[1])).astype(np.int64)
    assert np.array_equal(idx.astype(np.int64), i0 + i1 * gi.div_b[0])
    dbl = (np.floor(xyzw[:, 0].astype(np.float64) / 0.1) - gi.min_b[0]).astype(np.int64)
    assert np.count_nonzero(dbl != i0) > 0      # the double formula disagrees on some points
    g = o.grid_readback()
    assert g["nr_points"][g["nr_points"] > 0].sum() + 0 <= 200_000
    cnt = np.bincount(idx, minlength=gi.div_b[0] * gi.div_b[1])
    assert np.array_equal(np.nonzero(cnt)[0], g["cell_idx"])
    assert np.array_equal(cnt[g["cell_idx"]], np.abs(g["nr_points"]).clip(min=0) + (g["nr_points"] == -1) * cnt[g["cell_idx"]])


def test_c1_match_recovers_truth_and_trace_is_consistent():
    pb = common.c1_problem()
    o = oa.Oracle(common.params(resolution=0.5))
    o.set_target(pb["tgt"]); o.set_source(pb["src"])
    r, tr = o.align_trace(pb["guess"])
    assert r.converged == 1 and 2 <= r.iters <= 36
    assert r.evals == tr.shape[0]
    err = np.array(r.pose) - pb["truth"]
    assert np.hypot(err[0], err[1]) < 0.03 and abs(err[2]) < np.deg2rad(0.5)
    # every accepted step length lies in [trans_eps / 2, step_size]
    steps = tr[1:, 4]
    assert np.all(steps >= 0.005 - 1e-15) and np.all(steps <= 0.1 + 1e-15)
    assert r.fitness < 0.01 and r.fitness == pytest.approx(o.fitness(list(r.pose)), rel=1e-15)
    assert r.point_evals == r.evals * pb["src"].shape[0]
    outside, _ = o.eval_stats()
    assert outside == 0        # radius hits never leave the 3x3 block


def test_empty_overlap_returns_guess_converged():
    """No neighbours anywhere -> g = H = 0 -> delta_p = 0 -> converged at the guess."""
    tgt, src = _one_cell_problem()
    o = oa.Oracle(common.params(resolution=0.5))
    o.set_target(tgt)
    far = src.copy(); far[:, 0] += 100.0
    o.set_source(far)
    r = o.align([0.0, 0.0, 0.0])
    assert r.converged == 1 and r.iters == 0 and r.evals == 1
    assert list(r.pose) == [0.0, 0.0, 0.0] and r.trans_prob == 0.0


def test_pose_fuser_diagonal():
    """PoseFuser with diagonal covariances (PoseFuser.cpp:3-36): K = S_hat (Q + S_hat)^-1."""
    L = oa.load()
    d = lambda v: (C.c_double * len(v))(*v)
    pred, est = d([1.0, 2.0, 10.0]), d([1.2, 2.1, 12.0])
    motion, last = d([0.0, 0.0, 0.0]), d([1.0, 2.0, 10.0])
    lastCov = d([0.04, 0, 0, 0, 0.04, 0, 0, 0, 0.01])
    Q = d([0.04, 0, 0, 0, 0.04, 0, 0, 0, 0.01])
    fused = (C.c_double * 3)(); cov = (C.c_double * 9)()
    L.oracle_fuse_pose(pred, est, motion, last, lastCov, Q, 0.5, 0.1, 0.1, fused, cov)
    assert list(fused) == pytest.approx([1.1, 2.05, 11.0], rel=1e-12)
    assert np.allclose(np.array(cov).reshape(3, 3), np.diag([0.02, 0.02, 0.005]), atol=1e-15)
