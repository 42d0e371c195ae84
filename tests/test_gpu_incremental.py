"""ndt_set_target_incremental: a map with a settled prefix and a provisional tail, maintained on the device (running per-cell
sums of the settled part, only the touched cells re-derived) must give EXACTLY the grid ndt_set_target builds from the
whole cloud -- every table the matcher reads, bit for bit -- and the same matches and fitness scores.
[REF src/PointCloudMap.cpp:119-134 (makeLocalMap: previous sub-map + thinned current sub-map), src/PoseEstimator.cpp:19]"""
import numpy as np
import pytest
import torch

import ndt_common as common
from ndt_slam_b200 import capi, synth
from oracle import oracle_api as oa

pytestmark = pytest.mark.gpu


def _tables(g):
    """The matcher's view of a grid, from the replication blob (flags = 0): per padded cell the probe centroid (NaN pattern
    included), the occupancy bitmap, and the 64-byte record behind every tree cell (slot numbering itself is arbitrary)."""
    nbytes = g.grid_blob_size(flags=0)
    blob = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    g.grid_export(blob.data_ptr(), nbytes, flags=0)
    b = blob.cpu().numpy()
    gi = g.grid_info()
    npad = (gi.div_b[0] + 4) * (gi.div_b[1] + 4)
    # header layout: magic u64, flags i32, reserved i32, GridDims, counters, offsets (int64 x 11); find offsets from the tail of the header
    hdr = b[:512].view(np.int64)
    total = None
    for w in range(hdr.shape[0] - 1, 0, -1):
        if hdr[w] == nbytes:
            total = w
            break
    assert total is not None
    off = hdr[total - 10: total]                 # slot, cen, occ, recs, leaf_id, leaf_range, sorted, tgt, nn_range, nn_pts
    slot = b[off[0]: off[0] + 4 * npad].view(np.int32)
    cen = b[off[1]: off[1] + 8 * npad].view(np.uint32).reshape(npad, 2)
    occ = b[off[2]: off[2] + 4 * ((npad + 31) // 32)].view(np.uint32)
    tree = ~np.isnan(cen.view(np.float32)[:, 0])
    recs = b[off[3]:].view(np.uint8)
    rec_of = np.zeros((npad, 64), np.uint8)
    idx = np.flatnonzero(tree)
    for c in idx:
        rec_of[c] = recs[64 * slot[c]: 64 * slot[c] + 64]
    return dict(cen=np.where(tree[:, None], cen, 0xFFFFFFFF), occ=occ.copy(), tree=tree, rec=rec_of, info=(list(gi.min_b), list(gi.div_b), gi.n_points, gi.n_leaves, gi.n_valid))


def _same(a, b):
    assert a["info"] == b["info"]
    assert np.array_equal(a["tree"], b["tree"]) and np.array_equal(a["cen"], b["cen"]) and np.array_equal(a["occ"], b["occ"])
    ra, rb = a["rec"].copy(), b["rec"].copy()
    assert np.array_equal(ra, rb)              # centroid, count, cell position, mean, inverse covariance: all 64 bytes


def test_incremental_target_equals_full_rebuild_bit_for_bit():
    prm = common.params(resolution=0.5)
    rng = synth.rng_for(2026)
    segs = synth.office(11, 30.0, 20.0, 10)
    walls = synth.sample_walls(segs, 0.004, 0.01, rng)               # dense: hundreds of points per cell
    walls = walls[rng.permutation(walls.shape[0])]
    base = synth.to_xyzw(walls[:60000])                               # the "previous sub-map": settled from the start
    g_inc, src = capi.Ndt(prm), None
    scan = synth.raycast(segs, (12.0, 9.0, 0.4), rng)
    src = oa.approx_voxel_filter(synth.to_xyzw(common.prep_scan(scan)), 0.05)
    g_inc.set_source(src)
    settled = base.shape[0]
    cloud = base
    n_prev = 0
    used_incremental = 0
    for step in range(14):
        # the settled part grows by a few hundred points, the provisional tail (<= ~600 points) is replaced entirely
        grow = synth.to_xyzw(walls[60000 + 700 * step: 60000 + 700 * step + int(rng.integers(200, 700))])
        tail = synth.to_xyzw(walls[rng.integers(0, walls.shape[0], size=int(rng.integers(50, 600)))] + rng.normal(0, 0.01, (1, 2)))
        if step == 9:
            tail = np.zeros((0, 4), np.float32)                        # an empty tail
        if step == 11:
            grow = np.concatenate([grow, np.array([[np.nan, 1.0, 0, 0], [1.0, np.inf, 0, 0]], np.float32)])   # non-finite points are skipped
        new_cloud = np.ascontiguousarray(np.concatenate([cloud[:settled], grow, tail]))
        n_same = settled
        settled_new = settled + grow.shape[0]
        l0 = g_inc.launch_count()
        g_inc.set_target(new_cloud, n_same=n_same if step else 0, n_stable=settled_new)
        g_full = capi.Ndt(prm)
        g_full.set_target(new_cloud)
        _same(_tables(g_inc), _tables(g_full))
        g_full.set_source(src)
        guess = [12.03, 8.96, 0.41]
        a, b = g_inc.align(guess), g_full.align(guess)
        assert list(a.pose) == list(b.pose) and a.score == b.score and a.evals == b.evals and list(a.hess) == list(b.hess)
        assert a.fitness == b.fitness                                  # exact 1-NN on the lattice == exact 1-NN on the ordered buckets
        if step:
            with pytest.raises(capi.NdtError, match="ndt_grid_readback"):
                g_inc.grid_readback()
            used_incremental += 1
        cloud, settled = new_cloud, settled_new
    assert used_incremental >= 10
    # a tail point outside the current grid moves the bounds: falls back to a full build, still identical
    far = np.concatenate([cloud[:settled], np.array([[95.0, -40.0, 0, 0]], np.float32)])
    g_inc.set_target(np.ascontiguousarray(far), n_same=settled, n_stable=settled)
    g_full = capi.Ndt(prm); g_full.set_target(np.ascontiguousarray(far))
    _same(_tables(g_inc), _tables(g_full))
    g_inc.grid_readback()                                              # a full build: the read-back tables are current again
    # the promise broken (the settled prefix shrinks): full build as well
    g_inc.set_target(np.ascontiguousarray(cloud[:50000]), n_same=50000, n_stable=50000)
    g_full = capi.Ndt(prm); g_full.set_target(np.ascontiguousarray(cloud[:50000]))
    _same(_tables(g_inc), _tables(g_full))
    # batched kernels on an incrementally maintained grid (neighbour masks derived on demand)
    g_inc.set_target(np.ascontiguousarray(cloud[:50500]), n_same=50000, n_stable=50200)
    g_full = capi.Ndt(prm); g_full.set_target(np.ascontiguousarray(cloud[:50500])); g_full.set_source(src)
    guesses = np.array([12.0, 9.0, 0.4]) + rng.normal(0, [0.1, 0.1, 0.02], size=(80, 3))
    ra, rb = g_inc.align_batch(guesses), g_full.align_batch(guesses)
    assert np.array_equal(ra["pose"], rb["pose"]) and np.array_equal(ra["score"], rb["score"])
