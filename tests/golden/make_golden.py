"""Generates tests/golden/*.npz from oracle/_ref (the reference's own sources + the restated 6-DoF
mini-PCL). Run in the container that has /root/reference:  python tests/golden/make_golden.py
The vectors pin the hot path independently of the plain oracle and travel with the repo (the GPU box
has no /root/reference)."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import ndt_common as common  # noqa: E402
from ndt_slam_b200 import synth  # noqa: E402
from oracle import oracle_api as oa, ref_api as ra  # noqa: E402

OUT = Path(__file__).resolve().parent


def c1_vectors(seed, resolution):
    ra.set_params(Resolution=resolution)
    d = synth.c1_pair(seed)
    rsa, rsb = ra.resample(d["scan_a"]), ra.resample(d["scan_b"])           # ScanPointResampler (reference code)
    tgt = synth.to_xyzw(synth.transform(rsa, d["pose_a"]))
    src = ra.voxel_filter(synth.to_xyzw(rsb), 0.05)                           # ApproximateVoxelGrid (restated)
    n = ra.RefNdt(resolution)
    n.set_target(tgt); n.set_source(src)
    g = n.grid()
    rng = synth.rng_for(900 + seed)
    guess = np.array(d["pose_a"])
    poses = guess + rng.normal(0, [0.05, 0.05, 0.01], size=(12, 3))
    ev = np.array([np.concatenate([[e["score"]], e["grad"], e["hess"]]) for e in (n.eval(p) for p in poses)])
    guesses = guess + rng.normal(0, [0.12, 0.12, 0.025], size=(10, 3))
    guesses[0] = guess
    al = [n.align(q) for q in guesses]
    init_deg = [guess[0], guess[1], np.rad2deg(guess[2])]
    cost, est, cov = ra.estimate_pose(rsb, tgt, init_deg)                     # PoseEstimator::estimatePose (reference code)
    np.savez_compressed(
        OUT / f"c1_seed{seed}_res{resolution}.npz",
        scan_a=d["scan_a"], scan_b=d["scan_b"], pose_a=np.array(d["pose_a"]), pose_b=np.array(d["pose_b"]),
        resampled_a=rsa, resampled_b=rsb, tgt=tgt, src=src,
        grid_cell=g["cell_idx"], grid_nr=g["nr_points"], grid_mean=g["mean"], grid_icov=g["icov"], grid_centroid=g["centroid"],
        grid_min_b=g["min_b"], grid_div_b=g["div_b"],
        eval_poses=poses, eval_out=ev,
        align_guesses=guesses, align_pose=np.array([a["pose"] for a in al]), align_score=np.array([a["score"] for a in al]),
        align_iters=np.array([a["iters"] for a in al]), align_evals=np.array([a["evals"] for a in al]),
        align_fitness=np.array([a["fitness"] for a in al]), align_hess=np.array([a["hess"] for a in al]),
        est_init_deg=np.array(init_deg), est_cost=cost, est_pose_deg=est, est_cov=cov, resolution=resolution)


def host_vectors():
    """Pose algebra + EKF fusion from the reference's Pose2D.cpp / MyUtil.cpp / PoseFuser.cpp."""
    ra.set_params()
    rng = synth.rng_for(77)
    rows = []
    for _ in range(40):
        last = np.array([rng.uniform(-20, 20), rng.uniform(-20, 20), rng.uniform(-180, 180)])
        motion = np.array([rng.uniform(0, 0.3), rng.uniform(-0.05, 0.05), rng.uniform(-5, 5)])
        pred = ra.cal_pred_pose(motion, last)
        est = pred + np.array([rng.normal(0, 0.02), rng.normal(0, 0.02), rng.normal(0, 0.3)])
        A = rng.normal(size=(3, 3)) * 0.01
        last_cov = A @ A.T + np.diag([1e-4, 1e-4, 1e-5])
        B = rng.normal(size=(3, 3)) * 0.01
        Q = B @ B.T + np.diag([2e-4, 2e-4, 5e-6])
        fused, cov = ra.fuse_pose(pred, est, motion, last, last_cov, Q)
        ocov = ra.odometry_cov(motion, last, last_cov)
        cur = np.array([last[0] + rng.uniform(-1, 1), last[1] + rng.uniform(-1, 1), rng.uniform(-180, 180)])
        rows.append(np.concatenate([last, motion, pred, est, last_cov.ravel(), Q.ravel(), fused, cov.ravel(), ocov.ravel(),
                                    cur, ra.cal_motion(cur, last)]))
    np.savez_compressed(OUT / "host_math.npz", rows=np.array(rows),
                        layout="last3 motion3 pred3 est3 lastCov9 Q9 fused3 cov9 odoCov9 cur3 calMotion3 (delTime 0.5, coeVel 0.1, coeOmega 0.5)")


def sequence_vectors(n_scans=320):
    """C2-style run through the reference FrontEnd (FrontEnd.cpp / ScanMatcher.cpp / PointCloudMap.cpp), long enough to
    cross a sub-map split (sepThre 10 m at 0.05 m per scan). oracle/ref_shim.cpp zeroes ScanMatcher::lastCov, which the
    reference leaves uninitialised, so this run is reproducible (tests/test_ref_crosscheck.py checks that)."""
    ra.set_params(Resolution=0.5)
    seq = synth.c2_sequence(seed=2, n_scans=2000)
    slam = ra.RefSlam()
    odo_deg = np.column_stack([seq["odo"][:, 0], seq["odo"][:, 1], np.rad2deg(seq["odo"][:, 2])])
    odo_deg[:, 2] = (odo_deg[:, 2] + 180.0) % 360.0 - 180.0
    for i in range(n_scans):
        slam.process(i, odo_deg[i], seq["scans"][i])
    np.savez_compressed(OUT / f"c2_first{n_scans}.npz", poses=slam.poses(), odo_deg=odo_deg[:n_scans],
                        covs=slam.covs(), local_map=slam.local_map()[:, :2].copy(), n_global=slam.global_map().shape[0],
                        n_submaps=slam.submaps(), truth=seq["traj"][:n_scans])


def moving_object(scans, first=8, last=70):
    """A small box crossing the corridor: 14 points on a 0.35 m face, 0.12 m further along every scan (map frame)."""
    out = []
    for i, sc in enumerate(scans):
        if first <= i < last:
            c = sc.mean(axis=0) + np.array([-1.5 + 0.12 * (i - first), 0.4])
            box = c + np.stack([np.linspace(-0.175, 0.175, 14), 0.02 * np.sin(np.arange(14))], axis=1)
            sc = np.concatenate([sc, box], axis=0)
        out.append(sc)
    return out


def moving_vectors():
    """Moving-object removal (removeMoving = true, the launch default): the reference's PCFilter.h / PointCloudMap.cpp on
    the restated change-detector octree, a map replay with a moving box, and a FrontEnd run with sub-map splits."""
    poses, scans = map_replay_inputs(90)
    scans = moving_object(scans)
    prm = dict(removeMoving="true", sepThre=2.0, LeafSize=0.05, resol=0.05, thre_neighbor=0.2)
    ra.set_params(**prm)
    n_sub, local, glob = ra.map_replay(poses, scans)
    base = synth.to_xyzw(np.concatenate([scans[20], scans[22]])); test = synth.to_xyzw(scans[21])
    diff, kept = ra.pcfilter(base, test)
    # FrontEnd with removeMoving over 150 scans, sub-map split every 3 m
    ra.set_params(Resolution=0.5, removeMoving="true", sepThre=3.0, thre_neighbor=0.2)
    seq = synth.c2_sequence(seed=2, n_scans=2000)
    odo_deg = np.column_stack([seq["odo"][:, 0], seq["odo"][:, 1], np.rad2deg(seq["odo"][:, 2])])
    odo_deg[:, 2] = (odo_deg[:, 2] + 180.0) % 360.0 - 180.0
    slam = ra.RefSlam()
    n = 150
    for i in range(n):
        slam.process(i, odo_deg[i], seq["scans"][i])
    ra.set_params()
    np.savez_compressed(OUT / "moving_removal.npz", n_submaps=n_sub, local_map=local[:, :2].copy(), global_map=glob[:, :2].copy(),
                        pcf_base=base, pcf_test=test, pcf_diff=diff, pcf_kept=kept,
                        fe_poses=slam.poses(), fe_covs=slam.covs(), fe_local_map=slam.local_map()[:, :2].copy(), fe_submaps=slam.submaps())


def launcher_io_inputs():
    """A short text scan log with all three lidar groups populated, a trajectory and clouds for the writers."""
    seq = synth.c2_sequence(seed=2, n_scans=40)
    odo = np.column_stack([seq["odo"][:, 0], seq["odo"][:, 1], np.rad2deg(seq["odo"][:, 2])])[:9]
    rng = synth.rng_for(5150)
    front = [seq["scans"][i][::7] for i in range(9)]
    left = [rng.normal(size=(i % 4, 2)) * 3.0 for i in range(9)]
    right = [rng.normal(size=((i + 2) % 3, 2)) * 3.0 for i in range(9)]
    poses = rng.normal(size=(37, 3)) * [10.0, 10.0, 90.0]
    g = synth.to_xyzw(rng.normal(size=(300, 2)) * 20.0)
    g[5, 0] = 1e-7; g[6, 1] = -123456.789          # formatting corner cases (setprecision(8), exponent notation)
    return odo, front, left, right, poses, g, [g[:100], g[100:180], g[180:]]


def launcher_io_vectors():
    """The reference's own SlamLauncher.cpp (reader, poses writer) and PointCloudMap::saveGlobalMap (PCD writer)."""
    import tempfile
    from ndt_slam_b200 import host_api as ha      # only its log WRITER (test tooling), nothing of the product's reader
    odo, front, left, right, poses, g, subs = launcher_io_inputs()
    d = Path(tempfile.mkdtemp())
    ha.write_scan_log(d / "scan.txt", odo, front, left=left, right=right)
    out = {"log": np.frombuffer((d / "scan.txt").read_bytes(), np.uint8)}
    for side in (True, False):
        meta, xy = ra.launcher_parse(d / "scan.txt", side)
        out[f"meta_{int(side)}"], out[f"xy_{int(side)}"] = meta, xy
    ra.launcher_write_poses(d / "poses.txt", poses)
    out["poses_bytes"] = np.frombuffer((d / "poses.txt").read_bytes(), np.uint8)
    ra.save_maps(d / "map.pcd", d / "sub", g, subs)
    out["map_bytes"] = np.frombuffer((d / "map.pcd").read_bytes(), np.uint8)
    for k in range(3):
        out[f"sub{k}_bytes"] = np.frombuffer((d / f"sub{k}.pcd").read_bytes(), np.uint8)
    ra.set_params()
    np.savez_compressed(OUT / "launcher_io.npz", **out)


def map_replay_inputs(n_scans=150):
    """Map-frame scans along the C2 ground-truth trajectory (the inputs of PointCloudMap::addPose / addPoints)."""
    seq = synth.c2_sequence(seed=2, n_scans=2000)
    traj = seq["traj"][:n_scans]
    poses = np.column_stack([traj[:, 0], traj[:, 1], np.rad2deg(traj[:, 2])])
    scans = [synth.transform(oa.resample(seq["scans"][i], 0.05, 0.25), traj[i]) for i in range(n_scans)]
    return poses, scans


def map_vectors():
    """The reference's own PointCloudMap (src/PointCloudMap.cpp, compiled unmodified) driven scan by scan: sub-map
    splitting, concatenation, ApproximateVoxelGrid thinning, local and global map."""
    poses, scans = map_replay_inputs()
    ra.set_params(sepThre=1.0, LeafSize=0.2)
    n_sub, local, glob = ra.map_replay(poses, scans)
    ra.set_params()
    np.savez_compressed(OUT / "map_replay_sep1_leaf0.2.npz", n_submaps=n_sub, local_map=local[:, :2].copy(), global_map=glob[:, :2].copy())


if __name__ == "__main__":
    if not ra.available():
        raise SystemExit("oracle/_ref is not built (needs /root/reference): run `make -C oracle`")
    only = set(sys.argv[1:])              # e.g. `make_golden.py seq` regenerates one family
    if not only or "c1" in only:
        c1_vectors(1, 0.5)
        c1_vectors(2, 0.5)
        c1_vectors(3, 1.0)
    if not only or "host" in only:
        host_vectors()
    if not only or "seq" in only:
        sequence_vectors()
    if not only or "map" in only:
        map_vectors()
    if not only or "moving" in only:
        moving_vectors()
    if not only or "io" in only:
        launcher_io_vectors()
    for p in sorted(OUT.glob("*.npz")):
        print(p.name, p.stat().st_size)
