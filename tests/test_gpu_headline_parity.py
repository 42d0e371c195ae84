"""Parity under the headline: the BASELINE generators themselves (C4 relocalisation, C3 dense single match)
through the C ABI against the CPU oracle.

C4 (`synth.c4_reloc`, the workload bench.py's `value` is quoted on): a 402 x 402 grid, hypotheses tens of metres off
the map's structure, the persistent warp-per-match kernel (`k_align_warp`, n >= 64: staged occupancy bitmap + staged
scan, chunked work counter, out-of-range branch of `candidate_cell`) and the persistent score sweep (`k_eval_warp`).
C3 (`synth.c3_dense`): 4 M target points, 0.1 m cells (4096 x 4096), 65,536 source points, the whole-GPU cooperative
matcher (`k_align_grid`).

Bars (BASELINE.json north_star): iterations / evaluation counts identical, n_pairs identical, score / gradient /
Hessian within 1e-6 relative, final pose within 1e-4 m and 1e-5 rad. [REF src/PoseEstimator.cpp:17-29]"""
import numpy as np
import pytest

import ndt_common as common
from ndt_slam_b200 import capi, synth
from oracle import oracle_api as oa

pytestmark = pytest.mark.gpu

REL_EVAL = 1e-6
POSE_M, POSE_RAD = 1e-4, 1e-5


def c4_hypothesis_sample(d, n_lattice=1984, seed=77):
    """n_lattice hypotheses drawn across the whole 64 x 64 x 16 lattice + 64 adversarial ones: poses outside the map
    (the scan lies wholly off the grid: every point takes the out-of-range branch), poses on the border looking out,
    and poses next to the hidden true pose (long, well-conditioned Newton runs)."""
    rng = synth.rng_for(seed)
    hyp = d["hypotheses"]
    ids = np.sort(rng.choice(hyp.shape[0], n_lattice, replace=False))
    size = 200.0
    outside = np.stack([rng.uniform(-90.0, -40.0, 16), rng.uniform(0.0, size, 16), rng.uniform(-np.pi, np.pi, 16)], axis=1)
    outside2 = np.stack([rng.uniform(0.0, size, 16), rng.uniform(size + 35.0, size + 80.0, 16), rng.uniform(-np.pi, np.pi, 16)], axis=1)
    border = np.stack([rng.uniform(-2.0, 2.0, 16), rng.uniform(0.0, size, 16), np.pi + rng.uniform(-0.5, 0.5, 16)], axis=1)
    near = np.array(d["true_pose"]) + rng.normal(0.0, [0.25, 0.25, 0.03], size=(16, 3))
    return np.ascontiguousarray(np.concatenate([hyp[ids], outside, outside2, border, near], axis=0)), ids


@pytest.fixture(scope="module")
def c4():
    d = synth.c4_reloc(seed=4)
    src = oa.approx_voxel_filter(synth.to_xyzw(common.prep_scan(d["scan"])), common.LAUNCH["leaf"])
    tgt = synth.to_xyzw(d["map_pts"])
    prm = common.params(resolution=0.5)
    g = capi.Ndt(prm)
    g.set_target(tgt); g.set_source(src)
    return d, prm, g, tgt, src


def test_c4_grid_bit_exact(c4):
    d, prm, g, tgt, src = c4
    o = oa.Oracle(prm)
    o.set_target(tgt)
    a, b = g.grid_readback(), o.grid_readback()
    gi, go = g.grid_info(), o.grid_info()
    assert list(gi.div_b) == list(go.div_b) == [402, 402] and gi.n_slots == go.n_slots and gi.n_valid == go.n_valid
    for k in ("cell_idx", "nr_points", "centroid", "mean", "icov"):
        assert np.array_equal(a[k], b[k]), k


def test_c4_align_batch_matches_oracle_on_2048_hypotheses(c4):
    """k_align_warp on the headline workload: every sampled hypothesis ends where the oracle ends, by the same path."""
    d, prm, g, tgt, src = c4
    guesses, _ = c4_hypothesis_sample(d)
    assert guesses.shape[0] == 2048
    res = g.align_batch(guesses)                       # n >= 64: persistent warp-per-match kernel
    ref = common.oracle_align_many(prm, tgt, src, guesses)
    worst_m = worst_rad = worst_s = worst_h = 0.0
    iters_hist = {}
    for k, b in enumerate(ref):
        r = res[k]
        assert r["converged"] == b.converged and r["iters"] == b.iters and r["evals"] == b.evals, (k, guesses[k])
        assert r["point_evals"] == b.evals * src.shape[0]
        worst_m = max(worst_m, float(np.hypot(r["pose"][0] - b.pose[0], r["pose"][1] - b.pose[1])))
        worst_rad = max(worst_rad, float(abs(r["pose"][2] - b.pose[2])))
        if b.score != 0.0:
            worst_s = max(worst_s, abs(r["score"] - b.score) / abs(b.score))
        else:
            assert r["score"] == 0.0
        hb = np.array(b.hess)
        if np.max(np.abs(hb)) > 0:
            worst_h = max(worst_h, common.rel_err(r["hess"], hb))
        iters_hist[b.iters] = iters_hist.get(b.iters, 0) + 1
    assert worst_m < POSE_M and worst_rad < POSE_RAD, (worst_m, worst_rad)
    assert worst_s < REL_EVAL and worst_h < REL_EVAL, (worst_s, worst_h)
    # the sample exercises what the headline exercises: scans wholly off the map (no hit at all: the matcher stops at
    # once) next to long Newton / More-Thuente runs
    assert iters_hist.get(0, 0) >= 16 and max(iters_hist) >= 10, iters_hist
    # the device arg-max agrees with the oracle's ranking
    bi, best = g.best_of(res)
    sc = np.array([b.score if b.converged else -np.inf for b in ref])
    assert bi == int(np.argmax(sc))
    # small batches (one CTA / cluster per match) give the same answers for the same guesses
    few = g.align_batch(np.ascontiguousarray(guesses[-16:]))
    for k in range(16):
        b = ref[2048 - 16 + k]
        assert few[k]["iters"] == b.iters and few[k]["evals"] == b.evals
        assert np.hypot(few[k]["pose"][0] - b.pose[0], few[k]["pose"][1] - b.pose[1]) < POSE_M
    o1 = oa.Oracle(prm); o1.set_target(tgt); o1.set_source(src)
    assert few[3]["fitness"] == pytest.approx(o1.fitness(list(few[3]["pose"])), rel=1e-9)


@pytest.mark.parametrize("team", [1, 2, 4, 8])
def test_c4_align_batch_teams_of_warps_match_oracle(c4, team):
    """ndt_params.align_team: 1, 2, 4 or 8 warps share one match (`k_align_warp` / `k_align_team<WPM>`). Every team size ends
    where the oracle ends, by the same path, and gives the same bytes when the call is repeated."""
    d, prm, g, tgt, src = c4
    guesses, _ = c4_hypothesis_sample(d)
    pick = np.r_[0:192, 1984:2048]                   # lattice sample + off-map + border + near-truth poses
    gs = np.ascontiguousarray(guesses[pick])
    gt = capi.Ndt(common.params(resolution=0.5, align_team=team))
    gt.set_target(tgt); gt.set_source(src)
    res = gt.align_batch(gs, want_fitness=True)
    ref = common.oracle_align_many(prm, tgt, src, gs)
    o1 = oa.Oracle(prm); o1.set_target(tgt); o1.set_source(src)
    for k, b in enumerate(ref):
        r = res[k]
        assert r["converged"] == b.converged and r["iters"] == b.iters and r["evals"] == b.evals, (team, k)
        assert np.hypot(r["pose"][0] - b.pose[0], r["pose"][1] - b.pose[1]) < POSE_M and abs(r["pose"][2] - b.pose[2]) < POSE_RAD
        if b.score != 0.0:
            assert abs(r["score"] - b.score) / abs(b.score) < REL_EVAL
        if np.max(np.abs(np.array(b.hess))) > 0:
            assert common.rel_err(r["hess"], np.array(b.hess)) < REL_EVAL
    for k in (3, 200, 250):
        assert res[k]["fitness"] == pytest.approx(o1.fitness(list(res[k]["pose"])), rel=1e-9)
    again = gt.align_batch(gs, want_fitness=True)
    assert again.tobytes() == res.tobytes()          # fixed-order team reduction: run-to-run deterministic


def test_c4_passes_run_are_the_oracle_passes_minus_exact_repeats(c4):
    """ndt_result.passes_run: the device skips exactly (a) line-search trials whose step equals the step of the trial before
    (same pose: the oracle's trace shows the same score again) and (b) the Hessian-only pass after a search (its trial passes
    carry the Hessian). Everything else the oracle evaluates is evaluated on the device: counted from the oracle's trace."""
    d, prm, g, tgt, src = c4
    guesses, _ = c4_hypothesis_sample(d)
    pick = np.r_[0:160, 1984:2048]
    res = g.align_batch(np.ascontiguousarray(guesses[pick]))     # n >= 64: k_align_warp
    o = oa.Oracle(prm); o.set_target(tgt); o.set_source(src); o.want_fitness(False)
    skipped_total = 0
    for k, i in enumerate(pick):
        b, tr = o.align_trace(list(guesses[i]))
        expect, last_step = 0, None
        for row in tr:                                 # x, y, yaw, score, a_t, kind
            kind = int(row[5])
            if kind == 0:
                expect += 1; last_step = None
            elif kind == 1:
                if last_step is None or row[4] != last_step:
                    expect += 1
                else:
                    assert row[3] == prev_score        # the oracle itself got the same number again
                last_step = row[4]; prev_score = row[3]
        assert res[k]["evals"] == b.evals == len(tr), (i, res[k]["evals"], b.evals)
        assert res[k]["passes_run"] == expect, (i, res[k]["passes_run"], expect)
        skipped_total += b.evals - expect
    assert skipped_total > 0


def test_c4_eval_batch_matches_oracle_score_sweep(c4):
    """k_eval_warp (the relocalisation score sweep) on 4,096 lattice hypotheses + the off-map ones."""
    d, prm, g, tgt, src = c4
    guesses, _ = c4_hypothesis_sample(d, n_lattice=4032, seed=78)
    assert guesses.shape[0] == 4096
    out = g.eval_batch(guesses)
    ref = common.oracle_eval_many(prm, tgt, src, guesses)
    assert np.array_equal(out[:, 13], ref[:, 13])                     # identical neighbour sets (float32 radius test)
    assert int(np.sum(ref[:, 13] == 0)) >= 32                          # scans wholly off the map
    nz = ref[:, 13] > 0
    assert np.all(out[~nz, :13] == 0.0)
    assert np.max(np.abs(out[nz, 0] - ref[nz, 0]) / np.abs(ref[nz, 0])) < REL_EVAL
    for k in np.flatnonzero(nz):
        assert common.rel_err(out[k, 1:4], ref[k, 1:4]) < REL_EVAL, k
        assert common.rel_err(out[k, 4:13], ref[k, 4:13]) < REL_EVAL, k
    nog = g.eval_batch(guesses, want_hessian=False)
    assert np.array_equal(nog[:, 13], ref[:, 13])
    assert np.allclose(nog[:, :4], out[:, :4], rtol=1e-11, atol=1e-12)


@pytest.fixture(scope="module")
def c3():
    d = synth.c3_dense(seed=3)
    tgt, src = synth.to_xyzw(d["target"]), synth.to_xyzw(d["source"])
    prm = common.params(resolution=0.1)
    g, o = capi.Ndt(prm), oa.Oracle(prm)
    g.set_target(tgt); g.set_source(src)
    o.set_target(tgt); o.set_source(src); o.want_fitness(False)
    return d, prm, g, o, tgt, src


def test_c3_grid_bit_exact_4m_points(c3):
    d, prm, g, o, tgt, src = c3
    gi, go = g.grid_info(), o.grid_info()
    assert list(gi.div_b) == list(go.div_b) and min(gi.div_b) >= 3900      # walls span 25 .. 384.6 m of the 409.6 m world
    for f in ("n_points", "n_leaves", "n_slots", "n_valid"):
        assert getattr(gi, f) == getattr(go, f), f
    a, b = g.grid_readback(), o.grid_readback()
    for k in ("cell_idx", "nr_points", "centroid", "mean", "icov"):
        assert np.array_equal(a[k], b[k]), k


def test_c3_eval_and_match_parity_whole_gpu_kernel(c3):
    d, prm, g, o, tgt, src = c3
    assert src.shape[0] == 65536
    rng = synth.rng_for(33)
    for k in range(4):
        pose = np.array(d["guess"]) + rng.normal(0, [0.03, 0.03, 1e-4])
        a, b = g.eval(pose), o.eval(pose)
        assert a.n_pairs == b.n_pairs
        assert a.score == pytest.approx(b.score, rel=REL_EVAL)
        assert common.rel_err(a.grad, b.grad) < REL_EVAL and common.rel_err(a.hess, b.hess) < REL_EVAL
    a, b = g.align(list(d["guess"])), o.align(list(d["guess"]))       # > 16,384 source points: k_align_grid
    assert a.converged == b.converged and a.iters == b.iters and a.evals == b.evals
    assert np.hypot(a.pose[0] - b.pose[0], a.pose[1] - b.pose[1]) < POSE_M and abs(a.pose[2] - b.pose[2]) < POSE_RAD
    assert a.score == pytest.approx(b.score, rel=REL_EVAL) and common.rel_err(a.hess, b.hess) < REL_EVAL
    tp = d["true_pose"]
    assert np.hypot(a.pose[0] - tp[0], a.pose[1] - tp[1]) < 0.01
    assert a.fitness == pytest.approx(o.fitness(list(a.pose)), rel=1e-9)
