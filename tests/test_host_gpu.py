"""GPU tests of the host mirror: PoseEstimator::estimatePose and the FrontEnd sequence on the CUDA
path, against what the reference's own classes produced (golden vectors from oracle/_ref)."""
from pathlib import Path

import numpy as np
import pytest

from ndt_slam_b200 import build, host_api as ha, synth

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"


@pytest.fixture(scope="module", autouse=True)
def _host_built():
    build.build_host()


@pytest.mark.parametrize("path", sorted(GOLD.glob("c1_seed*.npz")), ids=lambda p: p.stem)
def test_estimate_pose_matches_reference_pose_estimator(path):
    z = np.load(path)
    ha.set_params(Resolution=float(z["resolution"]))
    cost, est, cov, res = ha.estimate_pose(z["resampled_b"], z["tgt"], list(z["est_init_deg"]))
    ref = z["est_pose_deg"]
    assert np.hypot(est[0] - ref[0], est[1] - ref[1]) < 1e-4                 # 1e-4 m
    assert abs(np.deg2rad(est[2] - ref[2])) < 1e-5                            # 1e-5 rad
    assert cost == pytest.approx(float(z["est_cost"]), rel=1e-3)
    assert np.max(np.abs(cov - z["est_cov"])) / np.max(np.abs(z["est_cov"])) < 1e-3
    assert res.converged == 1


def _odo_deg(seq, n):
    o = np.column_stack([seq["odo"][:n, 0], seq["odo"][:n, 1], np.rad2deg(seq["odo"][:n, 2])])
    o[:, 2] = (o[:, 2] + 180.0) % 360.0 - 180.0
    return o


def test_frontend_sequence_matches_reference_frontend_scan_by_scan():
    """320 scans through FrontEnd::process (crossing the sub-map split at sepThre = 10 m) against the reference FrontEnd
    compiled from its own sources (reproducible: oracle/ref_shim.cpp zeroes ScanMatcher::lastCov).

    Teacher-forced: after every scan the map and the fusion state continue from the REFERENCE's pose / covariance for that
    scan, so every one of the 319 matches sees exactly the inputs the reference's match saw (the local maps are then
    bit-identical) and the per-match bar applies to each: 1e-4 m, 1e-5 rad."""
    z = np.load(GOLD / "c2_first320.npz")
    ref, ref_cov = z["poses"], z["covs"]
    n = ref.shape[0]
    assert n == 320 and ref_cov.shape == (n, 3, 3)
    ha.set_params(Resolution=0.5)
    seq = synth.c2_sequence(seed=2, n_scans=2000)
    odo = _odo_deg(seq, n)
    assert np.array_equal(odo, z["odo_deg"])
    slam = ha.Slam()
    for i in range(n):
        slam.process_forced(i, odo[i], seq["scans"][i], ref[i], ref_cov[i])
    poses, covs = slam.poses(), slam.covs()
    dpos = np.hypot(poses[:, 0] - ref[:, 0], poses[:, 1] - ref[:, 1])
    dyaw = np.abs(np.deg2rad(poses[:, 2] - ref[:, 2]))
    assert np.max(dpos) < 1e-4 and np.max(dyaw) < 1e-5, (np.max(dpos), np.max(dyaw), int(np.argmax(dpos)))
    scale = np.max(np.abs(ref_cov[1:]), axis=(1, 2))
    # the covariance is (-H)^-1 at the final pose: a final pose that differs inside the pose bar moves H a little
    assert np.max(np.max(np.abs(covs[1:] - ref_cov[1:]), axis=(1, 2)) / scale) < 1e-3
    assert np.median(np.max(np.abs(covs[1:] - ref_cov[1:]), axis=(1, 2)) / scale) < 1e-6
    assert slam.submaps() == int(z["n_submaps"]) == 2
    lm = slam.local_map()
    assert np.array_equal(lm[:, :2], z["local_map"])          # same poses in -> the reference's local map, bit for bit
    st = slam.stats()
    assert st["matches"] == n - 1 and st["point_evals"] > 0


def test_frontend_sequence_free_running_stays_on_the_reference_trajectory():
    """The same 320 scans without forcing. A match only converges to within TransformationEpsilon (0.01), so its result
    moves by up to ~1e-3 m when an upstream 1e-9 difference flips one discrete decision (a point across a 5 cm voxel face
    of the map filter, one More-Thuente trial more); the free-running trajectories therefore agree to millimetres, not to
    the per-match bar -- that one is checked scan by scan above."""
    z = np.load(GOLD / "c2_first320.npz")
    ref = z["poses"]
    n = ref.shape[0]
    ha.set_params(Resolution=0.5)
    seq = synth.c2_sequence(seed=2, n_scans=2000)
    odo = _odo_deg(seq, n)
    slam = ha.Slam()
    for i in range(n):
        slam.process(i, odo[i], seq["scans"][i])
    poses = slam.poses()
    dpos = np.hypot(poses[:, 0] - ref[:, 0], poses[:, 1] - ref[:, 1])
    dyaw = np.abs(np.deg2rad(poses[:, 2] - ref[:, 2]))
    assert np.max(dpos[:8]) < 1e-6                       # identical until the first discrete flip
    assert np.max(dpos) < 5e-3 and np.max(dyaw) < 1e-3, (np.max(dpos), np.max(dyaw))
    assert slam.submaps() == int(z["n_submaps"])
    from scipy.spatial import cKDTree
    lm, ref_lm = slam.local_map(), z["local_map"]
    assert abs(lm.shape[0] - ref_lm.shape[0]) < 0.01 * ref_lm.shape[0]
    dist, _ = cKDTree(ref_lm).query(lm[:, :2])
    assert np.max(dist) < 0.05 and np.mean(dist) < 0.002


def test_frontend_with_moving_object_removal_matches_reference_scan_by_scan():
    """removeMoving = true (the launch default, ndt_mapping.launch:20), 150 scans, sub-map split every 3 m: the reference's
    FrontEnd with its own PCFilter / PointCloudMap against the product (teacher-forced like the test above): every match
    within the per-match bar, the filtered local map bit-identical."""
    z = np.load(GOLD / "moving_removal.npz")
    ref, ref_cov = z["fe_poses"], z["fe_covs"]
    n = ref.shape[0]
    assert n == 150
    ha.set_params(Resolution=0.5, removeMoving="true", sepThre=3.0, thre_neighbor=0.2)
    seq = synth.c2_sequence(seed=2, n_scans=2000)
    odo = _odo_deg(seq, n)
    slam = ha.Slam()
    for i in range(n):
        slam.process_forced(i, odo[i], seq["scans"][i], ref[i], ref_cov[i])
    poses = slam.poses()
    dpos = np.hypot(poses[:, 0] - ref[:, 0], poses[:, 1] - ref[:, 1])
    dyaw = np.abs(np.deg2rad(poses[:, 2] - ref[:, 2]))
    assert np.max(dpos) < 1e-4 and np.max(dyaw) < 1e-5, (np.max(dpos), np.max(dyaw), int(np.argmax(dpos)))
    assert slam.submaps() == int(z["fe_submaps"]) >= 3
    assert np.array_equal(slam.local_map()[:, :2], z["fe_local_map"])
    ha.set_params()


def test_launcher_reads_text_log_and_writes_outputs(tmp_path):
    seq = synth.c2_sequence(seed=2, n_scans=2000)
    n = 25
    odo = _odo_deg(seq, n)
    log = tmp_path / "scan.txt"
    ha.write_scan_log(log, odo, seq["scans"][:n])
    ha.set_params(Resolution=0.5, filename_in=str(log), poses_name=str(tmp_path / "poses.txt"), map_name=str(tmp_path / "map.pcd"),
                  separated_map_name=str(tmp_path / "sub"), end_frame=1000, sidelidar="false")
    done = ha.launcher_run()
    assert done == n - 1          # like the reference, the record that hits EOF is parsed but not processed
    lines = (tmp_path / "poses.txt").read_text().splitlines()
    assert int(lines[0]) == n - 1 and len(lines) == 1 + (n - 1 + 9) // 10
    pcd = (tmp_path / "map.pcd").read_text().splitlines()
    assert pcd[0].startswith("# .PCD v0.7") and pcd[10] == "DATA ascii"
    assert int(pcd[9].split()[1]) == len(pcd) - 11 > 100
    # same run through the in-memory path gives the same trajectory
    ha.set_params(Resolution=0.5)
    slam = ha.Slam()
    for i in range(n - 1):
        slam.process(i, odo[i], seq["scans"][i])
    p0 = slam.poses()[0]
    got = np.array(lines[1].split(), dtype=float)
    assert np.allclose(got, p0, atol=1e-4)


def test_loop_closure_pipeline_verifies_revisits_in_one_batched_call():
    """f3: key frames -> pose-graph nodes + odometry arcs; revisit candidates verified by ONE ndt_match_pairs call per key
    frame; accepted matches become loop arcs. A short closed lap (the robot returns to its start): loop arcs must appear at
    the end, link the last key frames to the first ones, agree with the ground-truth relative pose, and every verification
    must be the oracle's per-pair match (iterations / evaluations identical, pose within the per-match bar)."""
    import ndt_common as common
    from oracle import oracle_api as oa
    n = 240
    segs = synth.office(7, 16.0, 12.0, 6)
    t = np.linspace(0.0, 2.0 * np.pi, n, endpoint=False)
    traj = np.column_stack([8.0 + 3.5 * np.cos(t), 6.0 + 2.5 * np.sin(t), t + np.pi / 2])       # an elliptic lap back to the start
    traj = np.concatenate([traj, traj[:12]])                                                    # ... and a little beyond
    rng = synth.rng_for(707)
    scans = [synth.raycast(segs, tuple(p), rng) for p in traj]
    c0, s0 = np.cos(traj[0, 2]), np.sin(traj[0, 2])
    rel = np.stack([c0 * (traj[:, 0] - traj[0, 0]) + s0 * (traj[:, 1] - traj[0, 1]), -s0 * (traj[:, 0] - traj[0, 0]) + c0 * (traj[:, 1] - traj[0, 1])], axis=1)
    odo = np.column_stack([rel[:, 0], rel[:, 1], np.rad2deg(traj[:, 2] - traj[0, 2])])
    odo[:, 2] = (odo[:, 2] + 180.0) % 360.0 - 180.0
    ha.set_params(Resolution=0.5, loop_closure="true", keyframe_skip=6, loop_radius=1.5, loop_min_travel=10.0, loop_max_candidates=4,
                  loop_score_thre=0.05)
    slam = ha.Slam()
    verified = []
    for i in range(traj.shape[0]):
        slam.process(i, odo[i], scans[i])
        rows, counts = slam.loops()
        if i % 6 == 0 and rows.shape[0]:
            verified.append((i, rows))
    rows, counts = slam.loops()
    arcs, nodes = slam.loop_arcs(), slam.nodes()
    assert counts["nodes"] == (traj.shape[0] + 5) // 6 and counts["arcs"] == counts["nodes"] - 1 + counts["loop_arcs"]
    assert counts["loop_arcs"] >= 3 and len(verified) >= 2
    assert all(i >= 200 for i, _ in verified)                    # nothing is a "revisit" before the lap closes
    # loop arcs: late key frames linked to early ones, relative pose = ground truth within a few centimetres
    for src, dst, rx, ry, rth, cost in arcs:
        a, b = traj[int(src) * 6], traj[int(dst) * 6]
        ca, sa = np.cos(a[2]), np.sin(a[2])
        gt = (ca * (b[0] - a[0]) + sa * (b[1] - a[1]), -sa * (b[0] - a[0]) + ca * (b[1] - a[1]))
        assert int(src) <= 3 and int(dst) >= 34 and np.hypot(rx - gt[0], ry - gt[1]) < 0.05 and cost <= 0.05
    # every verification of the last key frame against the oracle, pair by pair (same inputs: resampled scans, relative guess)
    i_last, rows = verified[-1]
    prm = common.params(resolution=0.5)
    o = oa.Oracle(prm)
    cur = synth.to_xyzw(oa.resample(scans[i_last], 0.05, 0.25))
    for r in rows:
        ref_scan = synth.to_xyzw(oa.resample(scans[int(r[1]) * 6], 0.05, 0.25))
        o.set_target(ref_scan); o.set_source(oa.approx_voxel_filter(cur, 0.05))
        # the guess the detector used: current key-frame pose seen from the candidate's (estimated poses, degrees)
        pc, pr = nodes[int(r[0])], nodes[int(r[1])]
        cr, sr = np.cos(np.deg2rad(pr[2])), np.sin(np.deg2rad(pr[2]))
        g = [cr * (pc[0] - pr[0]) + sr * (pc[1] - pr[1]), -sr * (pc[0] - pr[0]) + cr * (pc[1] - pr[1]), np.deg2rad((pc[2] - pr[2] + 180.0) % 360.0 - 180.0)]
        b = o.align(g)
        assert int(r[7]) == b.iters and int(r[8]) == b.evals and int(r[9]) == b.converged
        assert np.hypot(r[2] - float(np.float32(b.pose[0])), r[3] - float(np.float32(b.pose[1]))) < 1e-4
    ha.set_params()


def test_relocalizer_cpp_host_shards_hypotheses_over_replicated_grids():
    """The C++ Relocalizer (C ABI only: ndt_replicate_grid, device-space ndt_align_batch per handle, ndt_best_of_multi) gives
    the same matches and the same winner as one handle matching every hypothesis; as many GPUs as the box has (two handles
    on one GPU otherwise)."""
    import torch
    import ndt_common as common
    from ndt_slam_b200 import capi
    from oracle import oracle_api as oa
    d = synth.c4_reloc(seed=4)
    src = oa.approx_voxel_filter(synth.to_xyzw(common.prep_scan(d["scan"])), 0.05)
    tgt = synth.to_xyzw(d["map_pts"])
    rng = synth.rng_for(99)
    hyp = np.ascontiguousarray(d["hypotheses"][np.sort(rng.choice(65536, 1023, replace=False))])
    hyp = np.concatenate([hyp, [np.array(d["true_pose"]) + [0.2, -0.1, 0.02]]])
    ndev = torch.cuda.device_count()
    devices = list(range(ndev)) if ndev > 1 else [0, 0, 0]
    bi, best, res, ms = ha.relocalize(devices, tgt, src, hyp, want_all=True)
    g = capi.Ndt(common.params(resolution=0.5))
    g.set_target(tgt); g.set_source(src)
    ref = g.align_batch(hyp, want_fitness=False)
    # a shard may run with another number of warps per match than the whole batch (ndt_params.align_team = auto): same
    # path, another summation tree
    assert np.allclose(res["pose"], ref["pose"], rtol=0, atol=1e-9) and np.allclose(res["score"], ref["score"], rtol=1e-10, atol=0)
    assert np.array_equal(res["evals"], ref["evals"]) and np.array_equal(res["iters"], ref["iters"])
    bi_ref, best_ref = g.best_of(ref)
    assert bi == bi_ref == 1023 and best.score == pytest.approx(best_ref.score, rel=1e-10)
    assert np.hypot(best.pose[0] - d["true_pose"][0], best.pose[1] - d["true_pose"][1]) < 0.05 and ms > 0
