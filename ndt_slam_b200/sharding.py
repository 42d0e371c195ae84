"""Sharding of the batched workloads across GPUs: independent units (pose hypotheses,
scan pairs) are block-partitioned over ranks, the finished grid is replicated once, and there is no
collective per iteration. Only the final arg-max needs an exchange: one small all_gather of each rank's
best (score, global index, pose). One process per GPU; torch.distributed is plumbing (NCCL on GPUs, gloo
in the CPU tests)."""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def shard_range(n_total: int, rank: int, world: int) -> tuple[int, int]:
    """Block partition [r*H/G, (r+1)*H/G): contiguous, disjoint, covers everything, sizes differ by <= 1."""
    return (n_total * rank) // world, (n_total * (rank + 1)) // world


def replicate_blob(blob: torch.Tensor | None, nbytes_hint: int, src: int = 0, device="cuda") -> torch.Tensor:
    """Broadcast a flat uint8 grid blob (ndt_grid_export) from `src` to every rank. Returns the local copy."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return blob
    n = torch.tensor([nbytes_hint if blob is None else blob.numel()], dtype=torch.int64, device=device)
    dist.broadcast(n, src)
    if blob is None:
        blob = torch.empty(int(n.item()), dtype=torch.uint8, device=device)
    dist.broadcast(blob, src)
    return blob


def best_over_ranks(score: float, global_index: int, pose, device="cpu"):
    """Each rank contributes its best converged match; returns (score, global_index, pose, owner_rank) of the
    overall best (highest score, lowest global index on ties) on every rank. A rank with nothing converged
    passes score = -inf."""
    rec = torch.tensor([score, float(global_index), pose[0], pose[1], pose[2]], dtype=torch.float64, device=device)
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        allrec = rec[None, :]
    else:
        buf = [torch.empty_like(rec) for _ in range(world)]
        dist.all_gather(buf, rec)
        allrec = torch.stack(buf)
    a = allrec.cpu().numpy()
    order = np.lexsort((a[:, 1], -a[:, 0]))      # primary: score descending, secondary: index ascending
    w = int(order[0])
    return float(a[w, 0]), int(a[w, 1]), a[w, 2:5].copy(), w


def gather_shards(local: np.ndarray, n_total: int, device="cpu") -> np.ndarray:
    """Concatenate per-rank result rows (block partition by shard_range, in rank order) on every rank: the end of a
    sharded ndt_match_pairs run (C5), where every rank holds the results of its own pairs only. `local` is a 2-D
    float64 array with one row per unit of this rank's shard."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return np.ascontiguousarray(local)
    width = local.shape[1]
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    cap = max(b - a for a, b in sizes)
    pad = torch.zeros((cap, width), dtype=torch.float64, device=device)
    pad[: local.shape[0]] = torch.from_numpy(np.ascontiguousarray(local, dtype=np.float64)).to(device)
    buf = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(buf, pad)
    return np.concatenate([buf[r][: sizes[r][1] - sizes[r][0]].cpu().numpy() for r in range(world)], axis=0)
