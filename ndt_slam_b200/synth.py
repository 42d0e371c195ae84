"""Seeded synthetic 2-D LiDAR data for the NDT hot path.

The reference ships no data (its only dataset is a private path, ndt_mapping.launch:3), so every
workload is generated here, deterministically, with numpy's PCG64. The same arrays are handed to
the CPU oracle and to the CUDA path. Float32 casts happen exactly where the reference casts
(PoseEstimator.h:97-99, PointCloudMap.cpp:65-67); until then everything is float64.

Worlds are lists of wall segments; a scan is 1081 beams over 270 degrees (0.25 degree step),
r_max 30 m, Gaussian range noise, beams without a hit dropped, points in beam order in the sensor
frame -- the Cartesian layout the reference's log reader produces (SlamLauncher.cpp:56-66).
"""
from __future__ import annotations

import numpy as np

N_BEAMS = 1081
FOV_DEG = 270.0
R_MAX = 30.0


def rng_for(seed: int) -> np.random.Generator:
    return np.random.Generator(np.random.PCG64(seed))


def _rect(x0, y0, x1, y1):
    return [(x0, y0, x1, y0), (x1, y0, x1, y1), (x1, y1, x0, y1), (x0, y1, x0, y0)]


def _rot_rect(cx, cy, w, h, ang):
    c, s = np.cos(ang), np.sin(ang)
    pts = [(-w / 2, -h / 2), (w / 2, -h / 2), (w / 2, h / 2), (-w / 2, h / 2)]
    q = [(cx + c * px - s * py, cy + s * px + c * py) for px, py in pts]
    return [(q[i][0], q[i][1], q[(i + 1) % 4][0], q[(i + 1) % 4][1]) for i in range(4)]


def office(seed: int, width: float, height: float, n_boxes: int, x0: float = 0.0, y0: float = 0.0,
           keep_clear=None) -> np.ndarray:
    """Outer rectangle plus interior boxes/partitions, axis-aligned and 30-degree rotated. (S, 4) float64."""
    rng = rng_for(seed)
    segs = _rect(x0, y0, x0 + width, y0 + height)
    placed = 0
    guard = 0
    while placed < n_boxes and guard < 100 * n_boxes + 100:
        guard += 1
        w, h = rng.uniform(0.6, 3.0), rng.uniform(0.4, 2.0)
        cx = rng.uniform(x0 + 1.5, x0 + width - 1.5)
        cy = rng.uniform(y0 + 1.5, y0 + height - 1.5)
        rotated = (placed % 3 == 2)
        if keep_clear is not None:
            d = np.hypot(np.asarray(keep_clear)[:, 0] - cx, np.asarray(keep_clear)[:, 1] - cy)
            if d.min() < 0.5 * np.hypot(w, h) + 0.8:
                continue
        if rotated:
            segs += _rot_rect(cx, cy, w, h, np.deg2rad(30.0))
        else:
            segs += _rect(cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2)
        placed += 1
    return np.asarray(segs, dtype=np.float64)


def raycast(segs: np.ndarray, pose, rng: np.random.Generator | None, n_beams: int = N_BEAMS,
            fov_deg: float = FOV_DEG, r_max: float = R_MAX, sigma: float = 0.01) -> np.ndarray:
    """Scan `segs` from pose (x, y, yaw_rad). Returns (M, 2) float64 points in the sensor frame."""
    x, y, th = pose
    ang = np.deg2rad(np.linspace(-fov_deg / 2, fov_deg / 2, n_beams))
    d = np.stack([np.cos(ang + th), np.sin(ang + th)], axis=1)            # (B, 2) world directions
    p1 = segs[:, 0:2]
    e = segs[:, 2:4] - p1                                                   # (S, 2)
    w = p1 - np.array([x, y])                                               # (S, 2)
    denom = d[:, 0:1] * e[None, :, 1] - d[:, 1:2] * e[None, :, 0]           # (B, S)
    with np.errstate(divide="ignore", invalid="ignore"):
        t = (w[None, :, 0] * e[None, :, 1] - w[None, :, 1] * e[None, :, 0]) / denom
        u = (w[None, :, 0] * d[:, 1:2] - w[None, :, 1] * d[:, 0:1]) / denom
    ok = (np.abs(denom) > 1e-12) & (t > 1e-6) & (u >= 0.0) & (u <= 1.0)
    t = np.where(ok, t, np.inf)
    r = t.min(axis=1)
    hit = np.isfinite(r) & (r < r_max)
    if rng is not None and sigma > 0:
        r = r + rng.normal(0.0, sigma, size=r.shape)
    r = r[hit]
    a = ang[hit]
    return np.stack([r * np.cos(a), r * np.sin(a)], axis=1)


def transform(xy: np.ndarray, pose) -> np.ndarray:
    """Sensor frame -> map frame in float64 (ScanMatcher::growMap, ScanMatcher.cpp:99-100)."""
    x, y, th = pose
    c, s = np.cos(th), np.sin(th)
    return np.stack([c * xy[:, 0] - s * xy[:, 1] + x, s * xy[:, 0] + c * xy[:, 1] + y], axis=1)


def to_xyzw(xy: np.ndarray) -> np.ndarray:
    """double -> float32 pcl::PointXYZ layout {x, y, 0, pad} (PoseEstimator.h:97-99)."""
    out = np.zeros((xy.shape[0], 4), dtype=np.float32)
    out[:, 0] = xy[:, 0].astype(np.float32)
    out[:, 1] = xy[:, 1].astype(np.float32)
    return np.ascontiguousarray(out)


def sample_walls(segs: np.ndarray, spacing: float, sigma: float, rng: np.random.Generator) -> np.ndarray:
    """Points every `spacing` metres along every wall with isotropic Gaussian noise. (N, 2) float64."""
    out = []
    for x0, y0, x1, y1 in segs:
        L = float(np.hypot(x1 - x0, y1 - y0))
        k = max(int(L / spacing), 1)
        s = (np.arange(k) + 0.5) / k
        out.append(np.stack([x0 + s * (x1 - x0), y0 + s * (y1 - y0)], axis=1))
    pts = np.concatenate(out, axis=0)
    if sigma > 0:
        pts = pts + rng.normal(0.0, sigma, size=pts.shape)
    return pts


# ---------------------------------------------------------------------------------------------
# Workloads (BASELINE.json configs). Resampling / voxel filtering is NOT done here: callers run
# the scan through the resampler they are testing (oracle, reference build or product).
# ---------------------------------------------------------------------------------------------

def c1_pair(seed: int = 1):
    """C1: room 20 x 12 m + 6 boxes; pose A = (2, 3, 10 deg), pose B = A (+) (0.10, 0.02, 1.5 deg)."""
    rng = rng_for(seed)
    pose_a = (2.0, 3.0, np.deg2rad(10.0))
    ca, sa = np.cos(pose_a[2]), np.sin(pose_a[2])
    pose_b = (pose_a[0] + ca * 0.10 - sa * 0.02, pose_a[1] + sa * 0.10 + ca * 0.02, pose_a[2] + np.deg2rad(1.5))
    segs = office(seed, 20.0, 12.0, 6, keep_clear=[pose_a[:2], pose_b[:2]])
    scan_a = raycast(segs, pose_a, rng)
    scan_b = raycast(segs, pose_b, rng)
    return dict(segs=segs, pose_a=pose_a, pose_b=pose_b, scan_a=scan_a, scan_b=scan_b)


def loop_trajectory(n: int, width: float, height: float, margin: float = 4.0) -> np.ndarray:
    """Closed rounded-rectangle loop, n poses (x, y, yaw), constant arc-length step."""
    w, h, r = width - 2 * margin, height - 2 * margin, 3.0
    per = 2 * (w - 2 * r) + 2 * (h - 2 * r) + 2 * np.pi * r
    s = np.arange(n) * (per / n)
    out = np.zeros((n, 3))
    legs = [w - 2 * r, np.pi * r / 2, h - 2 * r, np.pi * r / 2, w - 2 * r, np.pi * r / 2, h - 2 * r, np.pi * r / 2]
    for i, si in enumerate(s):
        k = 0
        while si > legs[k] and k < 7:
            si -= legs[k]
            k += 1
        side, corner = k // 2, k % 2
        ang0 = side * np.pi / 2
        # corner centres of the rounded rectangle, counter-clockwise from bottom edge
        starts = [(margin + r, margin), (margin + w, margin + r), (margin + w - r, margin + h), (margin, margin + h - r)]
        cc = [(margin + w - r, margin + r), (margin + w - r, margin + h - r), (margin + r, margin + h - r), (margin + r, margin + r)]
        if corner == 0:
            x = starts[side][0] + np.cos(ang0) * si
            y = starts[side][1] + np.sin(ang0) * si
            th = ang0
        else:
            a = si / r
            x = cc[side][0] + r * np.cos(ang0 - np.pi / 2 + a)
            y = cc[side][1] + r * np.sin(ang0 - np.pi / 2 + a)
            th = ang0 + a
        out[i] = (x, y, th)
    return out


def c2_sequence(seed: int = 2, n_scans: int = 2000, width: float = 40.0, height: float = 25.0):
    """C2: office 40 x 25 m, closed loop of n_scans poses, odometry = truth + seeded drift."""
    rng = rng_for(seed)
    traj = loop_trajectory(n_scans, width, height)
    segs = office(seed, width, height, 24, keep_clear=traj[::20, :2])
    scans = [raycast(segs, tuple(p), rng) for p in traj]
    # odometry: integrate true motions with multiplicative drift (0.5 %/m) and a yaw bias (0.02 deg/scan)
    odo = np.zeros_like(traj)
    odo[0] = (0.0, 0.0, 0.0)
    for i in range(1, n_scans):
        dxw, dyw = traj[i, 0] - traj[i - 1, 0], traj[i, 1] - traj[i - 1, 1]
        c, s = np.cos(traj[i - 1, 2]), np.sin(traj[i - 1, 2])
        mx, my = c * dxw + s * dyw, -s * dxw + c * dyw
        dth = np.arctan2(np.sin(traj[i, 2] - traj[i - 1, 2]), np.cos(traj[i, 2] - traj[i - 1, 2]))
        mx *= 1.005 + rng.normal(0, 0.002)
        my *= 1.005 + rng.normal(0, 0.002)
        dth += np.deg2rad(0.02) + rng.normal(0, np.deg2rad(0.01))
        co, so = np.cos(odo[i - 1, 2]), np.sin(odo[i - 1, 2])
        odo[i] = (odo[i - 1, 0] + co * mx - so * my, odo[i - 1, 1] + so * mx + co * my, odo[i - 1, 2] + dth)
    return dict(segs=segs, traj=traj, odo=odo, scans=scans)


def c3_dense(seed: int = 3, world: float = 409.6, n_target: int = 4_000_000, n_source: int = 65_536,
             spacing: float = 0.0125, yaw_deg: float = 0.01):
    """C3: 409.6 m square world, ~50 km of walls sampled every 12.5 mm (+5 mm noise) -> ~4 M target
    points; source = n_source wall points seen from a sensor at the world centre whose pose the matcher
    must recover from a guess that is off by (0.05 m, -0.03 m, yaw_deg).
    SURVEY.md 8(d) quotes a 0.3 deg yaw offset: over a 409 m cloud that moves far points by > 1 m, ten cell
    sizes at 0.1 m, and NDT (GPU and CPU alike) converges to a wrong local optimum. 0.01 deg (3.5 cm at
    200 m) keeps the workload shape and makes it well-posed; the true pose is returned for checking."""
    rng = rng_for(seed)
    total_len = n_target * spacing
    segs = []
    acc = 0.0
    while acc < total_len:
        L = rng.uniform(5.0, 40.0)
        cx, cy = rng.uniform(25.0, world - 25.0, size=2)
        if rng.random() < 0.7:
            ang = rng.choice([0.0, np.pi / 2])
        else:
            ang = rng.uniform(0, np.pi)
        dx, dy = 0.5 * L * np.cos(ang), 0.5 * L * np.sin(ang)
        segs.append((cx - dx, cy - dy, cx + dx, cy + dy))
        acc += L
    segs = np.asarray(segs)
    target = sample_walls(segs, spacing, 0.005, rng)
    target = np.clip(target, 0.05, world - 0.05)
    pick = rng.choice(target.shape[0], size=n_source, replace=False)
    pick.sort()
    src_map = target[pick] + rng.normal(0.0, 0.005, size=(n_source, 2))
    centre = 0.5 * world
    true_pose = (centre + 0.05, centre - 0.03, np.deg2rad(yaw_deg))   # sensor pose in the map; guess = (centre, centre, 0)
    c, s = np.cos(true_pose[2]), np.sin(true_pose[2])
    d = src_map - np.array(true_pose[:2])
    src = np.stack([c * d[:, 0] + s * d[:, 1], -s * d[:, 0] + c * d[:, 1]], axis=1)
    return dict(target=target, source=src, true_pose=true_pose, guess=(centre, centre, 0.0), segs=segs)


def c4_reloc(seed: int = 4, size: float = 200.0, n_xy: int = 64, n_th: int = 16):
    """C4: 200 x 200 m map of rooms; one 1081-beam scan from a hidden pose; n_xy^2 * n_th jittered hypotheses."""
    rng = rng_for(seed)
    segs = [np.asarray(_rect(0.0, 0.0, size, size))]
    room = 25.0
    k = 0
    for iy in range(int(size // room)):
        for ix in range(int(size // room)):
            segs.append(office(seed * 1000 + k, room - 3.0, room - 3.0, 5, x0=ix * room + 1.5, y0=iy * room + 1.5))
            k += 1
    segs = np.concatenate(segs, axis=0)
    map_pts = sample_walls(segs, 0.05, 0.01, rng)
    true_pose = (size * 0.43 + 3.1, size * 0.57 - 2.2, np.deg2rad(33.0))
    scan = raycast(segs, true_pose, rng)
    pitch = size / n_xy
    gx, gy, gt = np.meshgrid((np.arange(n_xy) + 0.5) * pitch, (np.arange(n_xy) + 0.5) * pitch,
                             np.arange(n_th) * (2 * np.pi / n_th) - np.pi, indexing="ij")
    hyp = np.stack([gx.ravel(), gy.ravel(), gt.ravel()], axis=1)
    hyp[:, 0:2] += rng.uniform(-0.25 * pitch, 0.25 * pitch, size=(hyp.shape[0], 2))
    hyp[:, 2] += rng.uniform(-np.deg2rad(5), np.deg2rad(5), size=hyp.shape[0])
    return dict(segs=segs, map_pts=map_pts, scan=scan, true_pose=true_pose, hypotheses=hyp)


def c5_pair(i: int):
    """C5 pair i: seeds 1000 + i; scan A at the origin pose of a random room, scan B offset by
    U(+-0.3 m, +-0.3 m, +-5 deg). Returns sensor-frame scans and the true offset (B in A's frame)."""
    seed = 1000 + i
    rng = rng_for(seed)
    w, h = rng.uniform(10.0, 24.0), rng.uniform(8.0, 16.0)
    pa = (rng.uniform(3.0, w - 3.0), rng.uniform(3.0, h - 3.0), rng.uniform(-np.pi, np.pi))
    off = (rng.uniform(-0.3, 0.3), rng.uniform(-0.3, 0.3), np.deg2rad(rng.uniform(-5.0, 5.0)))
    c, s = np.cos(pa[2]), np.sin(pa[2])
    pb = (pa[0] + c * off[0] - s * off[1], pa[1] + s * off[0] + c * off[1], pa[2] + off[2])
    segs = office(seed, w, h, 5, keep_clear=[pa[:2], pb[:2]])
    return dict(scan_a=raycast(segs, pa, rng), scan_b=raycast(segs, pb, rng), offset=off)
