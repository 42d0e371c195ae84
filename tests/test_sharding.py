"""Multi-GPU logic without GPUs: world_size-2 gloo processes shard a small relocalisation batch, each
runs its shard (CPU oracle standing in for the per-rank matcher), the grid 'blob' is replicated with one
broadcast, and the arg-max over ranks equals the single-process answer."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import ndt_common as common
from ndt_slam_b200 import sharding


def test_shard_range_properties():
    for n in (0, 1, 7, 64, 65_536, 65_537):
        for world in (1, 2, 3, 4, 8):
            r = [sharding.shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, ret):
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    sys.path.insert(0, str(root)); sys.path.insert(0, str(root / "tests"))
    import ndt_common as cm
    from oracle import oracle_api as oa
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pb = cm.c1_problem()
    rng = np.random.Generator(np.random.PCG64(5))
    hyp = pb["guess"] + rng.normal(0, [0.3, 0.3, 0.05], size=(48, 3))
    # "grid replication": rank 0 owns the target cloud bytes, everyone else receives them in one broadcast
    blob = torch.from_numpy(pb["tgt"].view(np.uint8).reshape(-1).copy()) if rank == 0 else None
    blob = sharding.replicate_blob(blob, 0, src=0, device="cpu")
    tgt = blob.numpy().view(np.float32).reshape(-1, 4)
    o = oa.Oracle(cm.params(resolution=0.5)); o.set_target(tgt); o.set_source(pb["src"]); o.want_fitness(False)
    lo, hi = sharding.shard_range(hyp.shape[0], rank, world)
    best = (-np.inf, -1, np.zeros(3))
    for i in range(lo, hi):
        r = o.align(hyp[i])
        if r.converged and (r.score > best[0]):
            best = (r.score, i, np.array(r.pose))
    score, gi, pose, owner = sharding.best_over_ranks(best[0], best[1], best[2], device="cpu")
    ret[rank] = (score, gi, list(pose), owner, lo, hi)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_and_best_matches_single_process():
    from oracle import oracle_api as oa
    world = 2
    mgr = mp.Manager(); ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert ret[0][:4] == ret[1][:4]                        # every rank ends with the same answer
    assert (ret[0][4], ret[0][5], ret[1][4], ret[1][5]) == (0, 24, 24, 48)
    pb = common.c1_problem()
    rng = np.random.Generator(np.random.PCG64(5))
    hyp = pb["guess"] + rng.normal(0, [0.3, 0.3, 0.05], size=(48, 3))
    o = oa.Oracle(common.params(resolution=0.5)); o.set_target(pb["tgt"]); o.set_source(pb["src"]); o.want_fitness(False)
    res = [o.align(h) for h in hyp]
    scores = np.array([r.score if r.converged else -np.inf for r in res])
    bi = int(np.argmax(scores))
    assert ret[0][1] == bi and ret[0][0] == pytest.approx(scores[bi], rel=1e-15)
    assert ret[0][3] == (0 if bi < 24 else 1)


def _pairs_worker(rank, world, port, ret):
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    sys.path.insert(0, str(root)); sys.path.insert(0, str(root / "tests"))
    import ndt_common as cm
    from ndt_slam_b200 import synth
    from oracle import oracle_api as oa
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_pairs = 7                                            # odd on purpose: ragged shards
    lo, hi = sharding.shard_range(n_pairs, rank, world)
    o = oa.Oracle(cm.params(resolution=0.5))
    rows = []
    for i in range(lo, hi):                                # every rank builds its own pairs' grids: nothing is exchanged
        d = synth.c5_pair(i)
        o.set_target(synth.to_xyzw(cm.prep_scan(d["scan_a"])))
        o.set_source(oa.approx_voxel_filter(synth.to_xyzw(cm.prep_scan(d["scan_b"])), cm.LAUNCH["leaf"]))
        r = o.align([0.0, 0.0, 0.0])
        rows.append([r.pose[0], r.pose[1], r.pose[2], r.score, float(r.converged), float(r.evals)])
    allrows = sharding.gather_shards(np.array(rows, dtype=np.float64).reshape(-1, 6), n_pairs, device="cpu")
    ret[rank] = allrows
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_pair_shards_gather_to_the_single_process_result():
    """C5-style sharding: scan pairs are block-partitioned, every rank matches its own, one all_gather at the end."""
    from ndt_slam_b200 import synth
    from oracle import oracle_api as oa
    world = 2
    mgr = mp.Manager(); ret = mgr.dict()
    mp.spawn(_pairs_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert ret[0].shape == (7, 6) and np.array_equal(ret[0], ret[1])
    o = oa.Oracle(common.params(resolution=0.5))
    for i in range(7):
        d = synth.c5_pair(i)
        o.set_target(synth.to_xyzw(common.prep_scan(d["scan_a"])))
        o.set_source(oa.approx_voxel_filter(synth.to_xyzw(common.prep_scan(d["scan_b"])), common.LAUNCH["leaf"]))
        r = o.align([0.0, 0.0, 0.0])
        assert list(ret[0][i, :3]) == list(r.pose) and ret[0][i, 3] == r.score and ret[0][i, 5] == r.evals
