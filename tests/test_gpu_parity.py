"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on identical seeded inputs.

Bars (BASELINE.json north_star): integer cell indexing and counts bit-exact; per-evaluation
score / gradient / Hessian within 1e-6 relative; final pose within 1e-4 m and 1e-5 rad."""
import numpy as np
import pytest
import torch

import ndt_common as common
from ndt_slam_b200 import capi, synth
from oracle import oracle_api as oa

pytestmark = pytest.mark.gpu

REL_EVAL = 1e-6
POSE_M, POSE_RAD = 1e-4, 1e-5


@pytest.fixture(scope="module")
def c1():
    pb = common.c1_problem()
    prm = common.params(resolution=0.5)
    g = capi.Ndt(prm)
    o = oa.Oracle(prm)
    g.set_target(pb["tgt"]); g.set_source(pb["src"])
    o.set_target(pb["tgt"]); o.set_source(pb["src"])
    return pb, g, o


def _assert_grid_equal(g, o, exact_float=True):
    gi_g, gi_o = g.grid_info(), o.grid_info()
    for f in ("n_points", "n_leaves", "n_slots", "n_valid"):
        assert getattr(gi_g, f) == getattr(gi_o, f), f
    assert list(gi_g.min_b) == list(gi_o.min_b) and list(gi_g.div_b) == list(gi_o.div_b)
    a, b = g.grid_readback(), o.grid_readback()
    assert np.array_equal(a["cell_idx"], b["cell_idx"])          # bit-exact indexing
    assert np.array_equal(a["nr_points"], b["nr_points"])        # bit-exact counts (incl. -1 flags)
    assert np.array_equal(a["centroid"], b["centroid"])          # float32 centroid, input-order sum
    if exact_float:
        assert np.array_equal(a["mean"], b["mean"])
        assert np.array_equal(a["icov"], b["icov"])
    else:
        assert common.rel_err(a["mean"], b["mean"]) < 1e-9
        assert common.rel_err(a["icov"], b["icov"]) < 1e-6


def test_c1_grid_bit_exact(c1):
    pb, g, o = c1
    _assert_grid_equal(g, o)
    assert np.array_equal(g.cell_index(pb["tgt"]), o.cell_index(pb["tgt"]))


@pytest.mark.parametrize("res,n,extent,seed", [(0.5, 20_000, 60.0, 3), (0.1, 200_000, 150.0, 4), (1.0, 3_000, 30.0, 5)])
def test_grid_bit_exact_random_walls(res, n, extent, seed):
    tgt = common.random_cloud(seed, n, extent)
    prm = common.params(resolution=res)
    g, o = capi.Ndt(prm), oa.Oracle(prm)
    g.set_target(tgt); o.set_target(tgt)
    _assert_grid_equal(g, o)
    assert np.array_equal(g.cell_index(tgt), o.cell_index(tgt))


def test_grid_scattered_points_and_quirk_variants():
    """Uniform scatter (many n < 6 leaves, degenerate cells) and the non-default covariance switches."""
    tgt = common.random_cloud(9, 50_000, 20.0, walls=False)
    for quirks in (capi.QUIRKS_PCL_1_10, capi.QUIRK_MT_INTERVAL_LT0 | capi.QUIRK_ANGLE_SNAP,
                   capi.QUIRKS_PCL_1_10 & ~capi.QUIRK_COV_INIT_IDENTITY):
        prm = common.params(resolution=0.5, quirks=quirks)
        g, o = capi.Ndt(prm), oa.Oracle(prm)
        g.set_target(tgt); o.set_target(tgt)
        _assert_grid_equal(g, o)


def test_grid_device_resident_input_and_nonfinite_points():
    tgt = common.random_cloud(6, 10_000, 40.0).copy()
    tgt[17, 0] = np.nan; tgt[4000, 1] = np.inf
    prm = common.params(resolution=0.5)
    g, o = capi.Ndt(prm), oa.Oracle(prm)
    o.set_target(tgt)
    d = torch.from_numpy(tgt).cuda()
    g.set_target(d.data_ptr(), n=tgt.shape[0], space=capi.MEM_DEVICE)
    _assert_grid_equal(g, o)
    assert g.grid_info().n_points == tgt.shape[0] - 2


def test_set_target_prefix_uploads_only_the_tail():
    """ndt_set_target_prefix: a cloud that only changed at its end (a local map growing scan by scan). The grid must be
    the one ndt_set_target builds from the whole cloud, across device-buffer reallocations."""
    prm = common.params(resolution=0.5)
    g, o = capi.Ndt(prm), oa.Oracle(prm)
    full = common.random_cloud(41, 120000, extent=60.0)
    n_prev, cloud = 0, None
    for n in (900, 2500, 2600, 40000, 40010, 120000):
        stable = max(n_prev - 300, 0)                      # the last 300 points of the previous version were replaced
        cloud = full[:n].copy()
        cloud[stable:n_prev, 0] += 0.013                  # ... by different ones
        full[stable:n_prev] = cloud[stable:n_prev]
        g.set_target(cloud, n_same=stable)
        o.set_target(cloud)
        _assert_grid_equal(g, o)
        n_prev = n
    # a stale or oversized promise degrades to a full upload
    g.set_target(cloud[:5000], n_same=10**9); o.set_target(cloud[:5000])
    _assert_grid_equal(g, o)


def test_grid_edge_cases():
    prm = common.params(resolution=0.5)
    g, o = capi.Ndt(prm), oa.Oracle(prm)
    empty = np.zeros((0, 4), np.float32)
    g.set_target(empty); o.set_target(empty)
    assert g.grid_info().n_leaves == 0 == o.grid_info().n_leaves
    one = np.array([[1.0, 2.0, 0, 0]], np.float32)
    g.set_target(one); o.set_target(one)
    _assert_grid_equal(g, o)
    # many identical points in one cell (long bucket)
    same = np.tile(np.array([[3.3, -4.4, 0, 0]], np.float32), (5000, 1))
    same[:, 0] += np.linspace(0, 0.05, 5000, dtype=np.float32)
    g.set_target(same); o.set_target(same)
    _assert_grid_equal(g, o)


def test_c1_eval_parity(c1):
    pb, g, o = c1
    rng = synth.rng_for(21)
    worst = 0.0
    for k in range(24):
        pose = pb["guess"] + rng.normal(0, [0.05, 0.05, 0.01])
        for want_h in (True, False):
            a, b = g.eval(pose, want_h), o.eval(pose, want_h)
            assert a.n_pairs == b.n_pairs            # identical neighbour sets (float32 radius test)
            assert a.score == pytest.approx(b.score, rel=REL_EVAL)
            worst = max(worst, common.rel_err(a.grad, b.grad))
            if want_h:
                worst = max(worst, common.rel_err(a.hess, b.hess))
            else:
                assert not np.any(np.array(a.hess))
    assert worst < REL_EVAL, worst


def test_eval_batch_matches_single(c1):
    pb, g, o = c1
    rng = synth.rng_for(22)
    poses = pb["guess"] + rng.normal(0, [0.2, 0.2, 0.05], size=(300, 3))
    out = g.eval_batch(poses)
    for k in (0, 17, 299):
        b = o.eval(poses[k])
        assert out[k, 0] == pytest.approx(b.score, rel=REL_EVAL)
        assert common.rel_err(out[k, 1:4], b.grad) < REL_EVAL
        assert common.rel_err(out[k, 4:13], b.hess) < REL_EVAL
        assert out[k, 13] == b.n_pairs
    d_p = torch.from_numpy(poses).cuda(); d_o = torch.zeros((300, 14), dtype=torch.float64, device="cuda")
    g.eval_batch(d_p.data_ptr(), n=300, space=capi.MEM_DEVICE, out=d_o.data_ptr())
    g.synchronize()
    assert np.array_equal(d_o.cpu().numpy(), out)      # deterministic, same kernel either way
    # thousands of poses take the persistent warp-per-pose kernel (score sweep): same numbers up to summation order
    many = pb["guess"] + rng.normal(0, [0.2, 0.2, 0.05], size=(3000, 3))
    many[:300] = poses
    big = g.eval_batch(many)
    assert np.array_equal(big[:300, 13], out[:, 13])
    assert np.allclose(big[:300, :13], out[:, :13], rtol=1e-11, atol=1e-12)
    b = o.eval(many[2999])
    assert big[2999, 0] == pytest.approx(b.score, rel=REL_EVAL) and big[2999, 13] == b.n_pairs
    assert common.rel_err(big[2999, 4:13], b.hess) < REL_EVAL
    nog = g.eval_batch(many, want_hessian=False)
    assert np.allclose(nog[:, :4], big[:, :4], rtol=1e-11, atol=1e-12)


def _assert_result_close(a, b):
    assert a.converged == b.converged and a.iters == b.iters and a.evals == b.evals
    assert a.point_evals == b.point_evals
    pa, pb_ = np.array(a.pose), np.array(b.pose)
    assert np.hypot(pa[0] - pb_[0], pa[1] - pb_[1]) < POSE_M
    assert abs(pa[2] - pb_[2]) < POSE_RAD
    assert a.score == pytest.approx(b.score, rel=REL_EVAL)
    assert a.trans_prob == pytest.approx(b.trans_prob, rel=REL_EVAL)
    assert common.rel_err(a.hess, b.hess) < REL_EVAL
    assert np.allclose(np.array(a.T), np.array(b.T), rtol=0, atol=2e-7)
    assert a.fitness == pytest.approx(b.fitness, rel=1e-9)


def test_c1_align_parity(c1):
    pb, g, o = c1
    a, b = g.align(pb["guess"]), o.align(pb["guess"])
    _assert_result_close(a, b)
    err = np.array(a.pose) - pb["truth"]
    assert np.hypot(err[0], err[1]) < 0.03


def test_align_parity_many_guesses_block_and_warp_kernels(c1):
    """The one-CTA-per-match kernel (n < 64) and the persistent warp-per-match kernel (n >= 64)."""
    pb, g, o = c1
    rng = synth.rng_for(23)
    guesses = pb["guess"] + rng.normal(0, [0.15, 0.15, 0.03], size=(96, 3))
    few = g.align_batch(guesses[:8])
    many = g.align_batch(guesses)
    trial_evals = 0
    for k in range(96):
        b = o.align(guesses[k])
        r = many[k]
        assert r["converged"] == b.converged and r["iters"] == b.iters and r["evals"] == b.evals, k
        assert np.hypot(r["pose"][0] - b.pose[0], r["pose"][1] - b.pose[1]) < POSE_M
        assert abs(r["pose"][2] - b.pose[2]) < POSE_RAD
        assert r["score"] == pytest.approx(b.score, rel=REL_EVAL)
        assert common.rel_err(r["hess"], b.hess) < REL_EVAL
        trial_evals += b.evals - b.iters - 1
        if k < 8:
            f = few[k]
            assert f["iters"] == b.iters and f["evals"] == b.evals
            assert np.allclose(f["pose"], r["pose"], rtol=0, atol=1e-9)
            assert f["fitness"] == pytest.approx(b.fitness, rel=1e-9)
    assert trial_evals > 0      # the More-Thuente inner loop was exercised
    bi, best = g.best_of(many)
    conv = many["converged"] == 1
    assert bi == int(np.argmax(np.where(conv, many["score"], -np.inf)))
    assert best.score == many["score"][bi]


def test_align_small_source_one_cta_kernel(c1):
    """Sources below the cluster threshold (600 points) take k_align_block: whole-grid tile in shared memory when it fits
    (C1-sized grid), global tables otherwise (a grid too large for the tile)."""
    pb, g, o = c1
    small = np.ascontiguousarray(pb["src"][::3])
    assert small.shape[0] < 600
    g.set_source(small); o.set_source(small)
    rng = synth.rng_for(29)
    for k in range(6):
        guess = pb["guess"] + rng.normal(0, [0.1, 0.1, 0.02])
        _assert_result_close(g.align(guess), o.align(guess))
    g.set_source(pb["src"]); o.set_source(pb["src"])
    # large grid: the tile does not fit in shared memory
    prm = common.params(resolution=0.5)
    g2, o2 = capi.Ndt(prm), oa.Oracle(prm)
    big = common.random_cloud(31, 60000, extent=400.0)
    sub = np.ascontiguousarray(big[1000:1400])
    c, s_ = np.cos(0.01), np.sin(0.01)
    srcp = sub.copy(); srcp[:, 0] = c * sub[:, 0] + s_ * sub[:, 1] - 0.05; srcp[:, 1] = -s_ * sub[:, 0] + c * sub[:, 1] + 0.04
    g2.set_target(big); o2.set_target(big); g2.set_source(srcp); o2.set_source(srcp)
    _assert_result_close(g2.align([0.0, 0.0, 0.0]), o2.align([0.0, 0.0, 0.0]))


def test_align_empty_overlap(c1):
    pb, g, o = c1
    far = pb["src"].copy(); far[:, 0] += 500.0
    g.set_source(far); o.set_source(far)
    a, b = g.align([0.0, 0.0, 0.0]), o.align([0.0, 0.0, 0.0])
    assert a.converged == 1 == b.converged and a.iters == 0 and a.evals == 1
    assert list(a.pose) == [0.0, 0.0, 0.0] and a.trans_prob == 0.0
    assert a.fitness == pytest.approx(b.fitness, rel=1e-9)   # exhaustive 1-NN fallback
    g.set_source(pb["src"]); o.set_source(pb["src"])


def test_voxel_filter_device_matches_oracle():
    for seed in (1, 2, 3):
        d = synth.c1_pair(seed)
        xyzw = synth.to_xyzw(common.prep_scan(d["scan_b"]))
        g = capi.Ndt(common.params())
        for leaf in (0.05, 0.1, 0.2):
            assert np.array_equal(g.approx_voxel_filter(xyzw, leaf), oa.approx_voxel_filter(xyzw, leaf))


def test_large_source_cluster_and_whole_gpu_kernels():
    """20,000 source points: the cooperative whole-GPU matcher (grid-wide reduction); the first 9,000 of them: the
    thread-block-cluster matcher (DSMEM reduction)."""
    rng = synth.rng_for(31)
    segs = synth.office(31, 60.0, 40.0, 20)
    tgt_xy = synth.sample_walls(segs, 0.005, 0.01, rng)
    pick = np.sort(rng.choice(tgt_xy.shape[0], 20_000, replace=False))
    true = (0.06, -0.04, np.deg2rad(0.4))
    c, s = np.cos(true[2]), np.sin(true[2])
    dd = tgt_xy[pick] + rng.normal(0, 0.005, (20_000, 2)) - np.array(true[:2])
    src = synth.to_xyzw(np.stack([c * dd[:, 0] + s * dd[:, 1], -s * dd[:, 0] + c * dd[:, 1]], axis=1))
    tgt = synth.to_xyzw(tgt_xy)
    prm = common.params(resolution=0.5)
    g, o = capi.Ndt(prm), oa.Oracle(prm)
    g.set_target(tgt); g.set_source(src); o.set_target(tgt); o.set_source(src)
    _assert_grid_equal(g, o)
    a, b = g.align([0.0, 0.0, 0.0]), o.align([0.0, 0.0, 0.0])
    _assert_result_close(a, b)
    assert np.hypot(a.pose[0] - true[0], a.pose[1] - true[1]) < 0.02
    a2 = g.align([0.0, 0.0, 0.0])                       # run-to-run determinism of the grid-wide reduction
    assert list(a2.pose) == list(a.pose) and a2.score == a.score
    few = g.align_batch(np.array([[0.0, 0.0, 0.0], [0.02, -0.01, 0.001]]))      # two matches, one cooperative launch each
    assert np.allclose(few[0]["pose"], a.pose, rtol=0, atol=0) and few[1]["converged"] == 1
    part = np.ascontiguousarray(src[:9000])
    g.set_source(part); o.set_source(part)
    _assert_result_close(g.align([0.0, 0.0, 0.0]), o.align([0.0, 0.0, 0.0]))


def test_grid_export_import_roundtrip(c1):
    pb, g, o = c1
    nbytes = g.grid_blob_size()
    blob = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    g.grid_export(blob.data_ptr(), nbytes)
    g2 = capi.Ndt(common.params(resolution=0.5))
    g2.grid_import(blob.data_ptr(), nbytes)
    g2.set_source(pb["src"])
    a, b = g2.align(pb["guess"]), g.align(pb["guess"])
    assert bytes(a) == bytes(b)        # bit-identical result from the replicated grid
    # the per-leaf read-back tables are not replicated: a defined error, not a read of stale buffers
    with pytest.raises(capi.NdtError, match="imported"):
        g2.grid_readback()
    # a replica without the target points: same match, no fitness; less than half the bytes
    nb2 = g.grid_blob_size(flags=0)
    assert nb2 < 0.6 * nbytes
    g.grid_export(blob.data_ptr(), nb2, flags=0)
    g3 = capi.Ndt(common.params(resolution=0.5))
    g3.grid_import(blob.data_ptr(), nb2)
    g3.set_source(pb["src"])
    c = g3.align(pb["guess"])
    assert list(c.pose) == list(b.pose) and c.score == b.score and c.evals == b.evals and np.isnan(c.fitness)
    # ndt_replicate_grid (peer copy between handles; same device here) + ndt_best_of_multi
    g4, g5 = capi.Ndt(common.params(resolution=0.5)), capi.Ndt(common.params(resolution=0.5))
    capi.replicate_grid([g, g4, g5])
    rng = synth.rng_for(61)
    guesses = pb["guess"] + rng.normal(0, [0.1, 0.1, 0.02], size=(192, 3))
    outs, ptrs = [], []
    for k, gg in enumerate((g, g4, g5)):
        gg.set_source(pb["src"])
        d_g = torch.from_numpy(np.ascontiguousarray(guesses[64 * k:64 * k + 64])).cuda()
        d_r = torch.zeros(64 * capi.RESULT_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
        gg.align_batch(d_g.data_ptr(), n=64, space=capi.MEM_DEVICE, out=d_r.data_ptr())
        gg.synchronize()
        outs.append(d_r); ptrs.append(d_r.data_ptr())
    allr = np.concatenate([np.frombuffer(t.cpu().numpy().tobytes(), dtype=capi.RESULT_DTYPE) for t in outs])
    ref = g.align_batch(np.ascontiguousarray(guesses))
    assert np.array_equal(allr["pose"], ref["pose"]) and np.array_equal(allr["score"], ref["score"])
    bh, bi, best = capi.best_of_multi([g, g4, g5], ptrs, [64, 64, 64])
    conv = allr["converged"] == 1
    k = int(np.argmax(np.where(conv, allr["score"], -np.inf)))
    assert (bh, bi) == (k // 64, k % 64) and best.score == allr["score"][k]


def test_grid_import_rejects_bad_blobs(c1):
    """Nothing in a blob header is trusted: wrong magic, truncated, offsets out of range, another resolution."""
    pb, g, o = c1
    nbytes = g.grid_blob_size()
    blob = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    g.grid_export(blob.data_ptr(), nbytes)
    g2 = capi.Ndt(common.params(resolution=0.5))
    with pytest.raises(capi.NdtError, match="truncated"):
        g2.grid_import(blob.data_ptr(), nbytes - 4096)
    bad = blob.clone(); bad[0] ^= 0xFF
    with pytest.raises(capi.NdtError, match="magic"):
        g2.grid_import(bad.data_ptr(), nbytes)
    hdr = blob[:512].cpu().numpy().copy()
    # find the off_recs field: corrupt every int64 of the offset table in turn, each must be rejected or harmless
    rejected = 0
    words = hdr.view(np.int64)
    for w in range(words.shape[0]):
        if 256 <= words[w] <= nbytes and words[w] % 256 == 0:        # looks like a section offset / the total
            bad = blob.clone()
            h2 = hdr.copy(); h2.view(np.int64)[w] = nbytes + 256 * (w + 1)
            bad[:512] = torch.from_numpy(h2).cuda()
            try:
                g2.grid_import(bad.data_ptr(), nbytes)
            except capi.NdtError:
                rejected += 1
    assert rejected >= 8
    g3 = capi.Ndt(common.params(resolution=1.0))
    with pytest.raises(capi.NdtError, match="resolution"):
        g3.grid_import(blob.data_ptr(), nbytes)
    with pytest.raises(capi.NdtError):
        capi.Ndt(common.params(resolution=0.5)).grid_blob_size()          # no target set


def test_run_to_run_determinism(c1):
    pb, g, o = c1
    r1, r2 = g.align(pb["guess"]), g.align(pb["guess"])
    assert bytes(r1) == bytes(r2)
    e1, e2 = g.eval(pb["guess"]), g.eval(pb["guess"])
    assert bytes(e1) == bytes(e2)


# ---- golden vectors generated by oracle/_ref (reference sources + 6-DoF mini-PCL) -------------------
from pathlib import Path  # noqa: E402

GOLD = Path(__file__).resolve().parent / "golden"


@pytest.mark.parametrize("path", sorted(GOLD.glob("c1_seed*.npz")), ids=lambda p: p.stem)
def test_cuda_path_matches_golden_vectors(path):
    z = np.load(path)
    prm = common.params(resolution=float(z["resolution"]))
    g = capi.Ndt(prm)
    assert np.array_equal(g.approx_voxel_filter(synth.to_xyzw(z["resampled_b"]), 0.05), z["src"])
    g.set_target(z["tgt"]); g.set_source(z["src"])
    a = g.grid_readback(); gi = g.grid_info()
    assert list(gi.min_b) == list(z["grid_min_b"]) and list(gi.div_b) == list(z["grid_div_b"])
    assert np.array_equal(a["cell_idx"], z["grid_cell"]) and np.array_equal(a["nr_points"], z["grid_nr"])
    assert np.array_equal(a["centroid"], z["grid_centroid"]) and np.array_equal(a["mean"], z["grid_mean"])
    m = z["grid_nr"] >= 6
    assert common.rel_err(a["icov"][m], z["grid_icov"][m]) < 1e-9
    out = g.eval_batch(np.ascontiguousarray(z["eval_poses"]))
    for k in range(out.shape[0]):
        assert out[k, 0] == pytest.approx(z["eval_out"][k, 0], rel=REL_EVAL)
        assert common.rel_err(out[k, 1:4], z["eval_out"][k, 1:4]) < REL_EVAL
        assert common.rel_err(out[k, 4:13], z["eval_out"][k, 4:13]) < REL_EVAL
    res = g.align_batch(np.ascontiguousarray(z["align_guesses"]))
    for k in range(res.shape[0]):
        assert res[k]["iters"] == z["align_iters"][k] and res[k]["evals"] == z["align_evals"][k]
        assert np.hypot(*(res[k]["pose"][:2] - z["align_pose"][k][:2])) < POSE_M
        assert abs(res[k]["pose"][2] - z["align_pose"][k][2]) < POSE_RAD
        assert res[k]["fitness"] == pytest.approx(z["align_fitness"][k], rel=1e-4)


def test_fitness_exact_with_dense_buckets():
    """Dense NDT buckets switch the 1-NN of getFitnessScore to the finer lattice: the result stays exact."""
    rng = synth.rng_for(51)
    segs = synth.office(51, 24.0, 16.0, 8)
    tgt = synth.to_xyzw(np.concatenate([synth.sample_walls(segs, 0.004, 0.01, rng) for _ in range(3)], axis=0))
    scan = synth.raycast(segs, (6.0, 5.0, 0.3), rng)
    src = oa.approx_voxel_filter(synth.to_xyzw(common.prep_scan(scan)), 0.05)
    prm = common.params(resolution=0.5)
    g, o = capi.Ndt(prm), oa.Oracle(prm)
    g.set_target(tgt); g.set_source(src); o.set_target(tgt); o.set_source(src)
    gi = g.grid_info()
    assert gi.n_points / gi.n_leaves > 24          # dense enough to trigger the fine lattice
    guess = [6.02, 4.97, 0.31]
    a, b = g.align(guess), o.align(guess)
    _assert_result_close(a, b)
    # a source far outside the map still gets the exact answer (ring search gives up, exhaustive scan)
    far = src.copy(); far[:, 1] += 300.0
    g.set_source(far); o.set_source(far)
    a, b = g.align([0.0, 0.0, 0.0]), o.align([0.0, 0.0, 0.0])
    assert a.fitness == pytest.approx(b.fitness, rel=1e-9)


def _c5_batch(ids):
    """Scan pairs as ndt_match_pairs takes them: raw (resampled) source clouds, target clouds, offsets."""
    srcs, tgts = [], []
    for i in ids:
        d = synth.c5_pair(i)
        tgts.append(synth.to_xyzw(common.prep_scan(d["scan_a"])))
        srcs.append(synth.to_xyzw(common.prep_scan(d["scan_b"])))
    return srcs, tgts


def _pack(clouds):
    off = np.zeros(len(clouds) + 1, np.int64)
    off[1:] = np.cumsum([c.shape[0] for c in clouds])
    pts = np.concatenate(clouds, axis=0) if off[-1] > 0 else np.zeros((0, 4), np.float32)
    return np.ascontiguousarray(pts, dtype=np.float32), off


def test_match_pairs_parity_with_oracle_per_pair():
    """C5 (loop-closure verification): one batched call == n x (filter, grid build, match, fitness) of the oracle.
    Includes ragged sizes, an empty target, an empty source and a single-point target."""
    prm = common.params(resolution=0.5)
    srcs, tgts = _c5_batch(range(70))
    tgts[5] = np.zeros((0, 4), np.float32)            # empty target
    srcs[9] = np.zeros((0, 4), np.float32)            # empty source
    tgts[11] = tgts[11][:1]                            # one target point: a grid with no tree cell
    srcs[13] = srcs[13][:40]; tgts[17] = tgts[17][:100]
    srcs[21] = np.concatenate([srcs[21], srcs[21][::2] + np.float32(0.003), srcs[21][::3] - np.float32(0.002)])   # > 1152 points: not staged in smem
    n = len(srcs)
    src, so = _pack(srcs); tgt, to = _pack(tgts)
    guesses = np.zeros((n, 3))
    g = capi.Ndt(prm)
    res = g.match_pairs(src, so, tgt, to, guesses, n, source_leaf=common.LAUNCH["leaf"])
    o = oa.Oracle(prm)
    worst_t = 0.0
    for k in range(n):
        fs = oa.approx_voxel_filter(srcs[k], common.LAUNCH["leaf"]) if srcs[k].shape[0] else srcs[k]
        o.set_target(tgts[k]); o.set_source(fs)
        b = o.align(guesses[k])
        r = res[k]
        assert r["converged"] == b.converged and r["iters"] == b.iters and r["evals"] == b.evals, k
        assert r["point_evals"] == b.evals * fs.shape[0], k           # identical filtered source size
        assert np.hypot(r["pose"][0] - b.pose[0], r["pose"][1] - b.pose[1]) < POSE_M, k
        assert abs(r["pose"][2] - b.pose[2]) < POSE_RAD, k
        assert r["score"] == pytest.approx(b.score, rel=REL_EVAL, abs=1e-12), k
        if fs.shape[0] and tgts[k].shape[0]:
            assert r["fitness"] == pytest.approx(b.fitness, rel=1e-9), k
        worst_t = max(worst_t, float(np.hypot(r["pose"][0] - b.pose[0], r["pose"][1] - b.pose[1])))
    # the batched call agrees with the single-match entry points of the same library (different reduction tree:
    # one CTA per match there, one warp per pair here)
    for k in (0, 3, 33):
        g1 = capi.Ndt(prm)
        g1.set_target(tgts[k]); g1.set_source(g1.approx_voxel_filter(srcs[k], common.LAUNCH["leaf"]))
        a = g1.align(guesses[k])
        assert np.allclose(a.pose, res[k]["pose"], rtol=0, atol=1e-9) and a.score == pytest.approx(res[k]["score"], rel=1e-10)
        assert a.iters == res[k]["iters"] and a.evals == res[k]["evals"]
    # the warp-per-pair schedule (used for large batches) agrees with the CTA-per-pair one used above
    g_w = capi.Ndt(common.params(resolution=0.5, pairs_schedule=capi.PAIRS_WARP))
    res_w = g_w.match_pairs(src, so, tgt, to, guesses, n, source_leaf=common.LAUNCH["leaf"])
    assert np.array_equal(res_w["iters"], res["iters"]) and np.array_equal(res_w["evals"], res["evals"])
    assert np.allclose(res_w["pose"], res["pose"], rtol=0, atol=1e-9)
    # small batches (several grid-build + match rounds inside one call) give the same bytes
    g_b = capi.Ndt(common.params(resolution=0.5, pairs_batch_points=9000))
    res_b = g_b.match_pairs(src, so, tgt, to, guesses, n, source_leaf=common.LAUNCH["leaf"])
    assert np.array_equal(res_b["pose"], res["pose"]) and np.array_equal(res_b["fitness"], res["fitness"])
    # device-resident inputs / outputs give the same bytes
    d_src, d_tgt = torch.from_numpy(src).cuda(), torch.from_numpy(tgt).cuda()
    d_g = torch.from_numpy(guesses).cuda()
    d_res = torch.zeros(n * capi.RESULT_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
    g.match_pairs(d_src.data_ptr(), so, d_tgt.data_ptr(), to, d_g.data_ptr(), n, source_leaf=common.LAUNCH["leaf"],
                  space=capi.MEM_DEVICE, out=d_res.data_ptr())
    g.synchronize()
    res2 = np.frombuffer(d_res.cpu().numpy().tobytes(), dtype=capi.RESULT_DTYPE)
    assert np.array_equal(res2["pose"], res["pose"]) and np.array_equal(res2["score"], res["score"])
    # compact (x, y) clouds, 8 bytes per point (ndt_match_pairs_xy): the same bytes, from host and from device buffers
    src0 = src.copy(); src0[:, 2:] = 0.0             # pair 21 above carries z != 0 (it takes part in the voxel hash): planar copy
    assert not tgt[:, 2:].any()
    res = g.match_pairs(src0, so, tgt, to, guesses, n, source_leaf=common.LAUNCH["leaf"])
    src_xy, tgt_xy = np.ascontiguousarray(src0[:, :2]), np.ascontiguousarray(tgt[:, :2])
    res_xy = g.match_pairs(src_xy, so, tgt_xy, to, guesses, n, source_leaf=common.LAUNCH["leaf"], xy=True)
    for f in ("pose", "score", "fitness", "iters", "evals", "hess"):
        assert np.array_equal(res_xy[f], res[f], equal_nan=(f == "fitness")), f
    d_sxy, d_txy = torch.from_numpy(src_xy).cuda(), torch.from_numpy(tgt_xy).cuda()
    d_res.zero_()
    g.match_pairs(d_sxy.data_ptr(), so, d_txy.data_ptr(), to, d_g.data_ptr(), n, source_leaf=common.LAUNCH["leaf"],
                  space=capi.MEM_DEVICE, out=d_res.data_ptr(), xy=True)
    g.synchronize()
    res3 = np.frombuffer(d_res.cpu().numpy().tobytes(), dtype=capi.RESULT_DTYPE)
    assert np.array_equal(res3["pose"], res["pose"]) and np.array_equal(res3["fitness"], res["fitness"], equal_nan=True)
    res_bxy = g_b.match_pairs(src_xy, so, tgt_xy, to, guesses, n, source_leaf=common.LAUNCH["leaf"], xy=True)   # several batches
    assert np.array_equal(res_bxy["pose"], res["pose"])
    # recovery of the true offset for ordinary pairs
    ok = 0
    for k in range(20, 70):
        off = synth.c5_pair(k)["offset"]
        if res[k]["converged"] and np.hypot(res[k]["pose"][0] - off[0], res[k]["pose"][1] - off[1]) < 0.05:
            ok += 1
    assert ok >= 35


def test_batched_entry_points_reject_bad_arguments_and_odd_knobs():
    """Error behaviour of the batched calls (no exception crosses the ABI: a negative status + ndt_last_error), and knobs
    outside their range fall back to the library's choice."""
    prm = common.params(resolution=0.5)
    g = capi.Ndt(prm)
    srcs, tgts = _c5_batch(range(3))
    src, so = _pack(srcs); tgt, to = _pack(tgts)
    guesses = np.zeros((3, 3))
    for xy in (False, True):
        s_in = np.ascontiguousarray(src[:, :2]) if xy else src
        t_in = np.ascontiguousarray(tgt[:, :2]) if xy else tgt
        bad = so.copy(); bad[0] = 1
        with pytest.raises(capi.NdtError, match="start at 0"):
            g.match_pairs(s_in, bad, t_in, to, guesses, 3, source_leaf=0.05, xy=xy)
        bad = to.copy(); bad[2] = bad[1] - 1
        with pytest.raises(capi.NdtError, match="non-decreasing"):
            g.match_pairs(s_in, so, t_in, bad, guesses, 3, source_leaf=0.05, xy=xy)
        ok = g.match_pairs(s_in, so, t_in, to, guesses, 3, source_leaf=0.05, xy=xy)      # the handle is still usable
        assert ok["converged"].all()
    assert g.match_pairs(src, so, tgt, to, guesses, 0).shape[0] == 0                       # nothing to do
    # align_team outside {0, 1, 2, 4, 8}: the library picks (same answers as the default)
    pb = common.c1_problem()
    rng = np.random.Generator(np.random.PCG64(3))
    hyp = pb["guess"] + rng.normal(0, [0.2, 0.2, 0.03], size=(80, 3))
    ga, gb = capi.Ndt(common.params(resolution=0.5)), capi.Ndt(common.params(resolution=0.5, align_team=3))
    for gg in (ga, gb):
        gg.set_target(pb["tgt"]); gg.set_source(pb["src"])
    assert ga.align_batch(hyp).tobytes() == gb.align_batch(hyp).tobytes()


def test_hit_rejected_by_the_guard_on_e_contributes_nothing(c1):
    """PCL's updateDerivatives returns 0 for a hit whose e = d2 * exp(-d2 q / 2) is NaN, negative or > 1: no score term, no
    gradient, no Hessian (the hit still counts as a neighbour). Unreachable with the inverse covariances the grid build
    produces, so the records are crafted: a replica is imported from a blob in which one tree cell's inverse covariance is
    strongly negative definite (q < 0, e > 1). It must give exactly what a grid without that cell gives. A blob with a
    non-finite record (the build never writes one) is refused by ndt_grid_import."""
    import torch
    pb, g, o = c1
    guess = np.array(pb["guess"])
    e0 = g.eval(guess)
    nbytes = g.grid_blob_size(flags=0)
    blob = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    g.grid_export(blob.data_ptr(), nbytes, flags=0)
    b = blob.cpu().numpy()
    gi = g.grid_info()
    npad = (gi.div_b[0] + 4) * (gi.div_b[1] + 4)
    hdr = b[:512].view(np.int64)
    at = [w for w in range(hdr.shape[0] - 1, 0, -1) if hdr[w] == nbytes][0]
    off = hdr[at - 10: at]                         # slot, cen, occ, recs, ...
    slot = b[off[0]: off[0] + 4 * npad].view(np.int32)
    cen = b[off[1]: off[1] + 8 * npad].view(np.float32).reshape(npad, 2)
    tree = np.flatnonzero(~np.isnan(cen[:, 0]))
    # the tree cell closest to the transformed scan: it certainly has hits at this pose
    c, s = np.cos(guess[2]), np.sin(guess[2])
    pts = pb["src"][:, :2].astype(np.float64) @ np.array([[c, s], [-s, c]]) + guess[:2]
    d = ((cen[tree][:, None, :].astype(np.float64) - pts[None, :, :]) ** 2).sum(axis=2)
    victim = int(tree[np.argmin(d.min(axis=1))])
    hits_on_victim = int((d[np.flatnonzero(tree == victim)[0]] < 0.25 - 1e-6).sum())
    assert hits_on_victim > 0

    def replica(edit):
        bb = b.copy()
        edit(bb)
        gg = capi.Ndt(common.params(resolution=0.5))
        t = torch.from_numpy(bb).cuda()
        gg.grid_import(t.data_ptr(), nbytes)
        gg.set_source(pb["src"])
        return gg.eval(guess)

    def icov_of(bb):
        r0 = off[3] + 64 * int(slot[victim]) + 32
        return bb[r0: r0 + 32].view(np.float64)

    def set_nan(bb): icov_of(bb)[:] = np.nan
    def set_negative(bb): icov_of(bb)[:] = [-1e12, 0.0, 0.0, -1e12]     # d2 * exp(d2 * 1e12 |d|^2 / 2) > 1 for every |d| > 4e-6 m
    def remove_cell(bb): bb[off[1] + 8 * victim: off[1] + 8 * victim + 8].view(np.uint32)[:] = 0xFFFFFFFF

    without = replica(remove_cell)
    removed = e0.n_pairs - without.n_pairs                 # (the float64 count above may differ by a point on the radius)
    assert removed >= 1 and abs(removed - hits_on_victim) <= 1 and without.score != e0.score
    e = replica(set_negative)
    assert e.n_pairs == e0.n_pairs                           # still a neighbour
    assert np.isfinite(e.score) and e.score == pytest.approx(without.score, rel=1e-12)
    assert np.allclose(np.array(e.grad), np.array(without.grad), rtol=1e-10, atol=1e-12)
    assert np.allclose(np.array(e.hess), np.array(without.hess), rtol=1e-10, atol=1e-10)
    with pytest.raises(capi.NdtError, match="non-finite"):
        replica(set_nan)
