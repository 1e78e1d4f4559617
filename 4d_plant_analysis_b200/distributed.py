"""Multi-GPU: one process per GPU (torch.distributed), segment pairs sharded across ranks.

The algorithm has no exchange step: the atomic unit is a matched segment pair (both
directions), ranks work on disjoint shards with no data-path collective, and the only communication
is the final gather of the dense maps.  ``shard_segments`` is the partitioner (longest-processing-time
greedy on a cost estimate); ``gather_maps`` all-gathers the int64 maps (NCCL on GPUs; gloo in the CPU
tests of the host logic).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np


def segment_costs(off1, off2, lm_off) -> np.ndarray:
    """work estimate per segment pair: ICP pairs (n1*n2/K, ~3.6 passes, two directions) + final-map queries"""
    n1, n2, K = np.diff(off1).astype(np.float64), np.diff(off2).astype(np.float64), np.diff(lm_off).astype(np.float64)
    return 2 * 3.6 * n1 * n2 / np.maximum(K, 1) + 40.0 * (n1 + n2)


def shard_segments(costs: Sequence[float], world: int, frame_pair: Optional[Sequence[int]] = None) -> List[np.ndarray]:
    """Deal segment pairs to ranks.  With ``frame_pair`` ids and at least `world` frame pairs, whole frame
    pairs are dealt round-robin (C4: 15 pairs -> 2,2,2,2,2,2,2,1); otherwise segment pairs are dealt greedily,
    heaviest first, to the least loaded rank.  Every rank's list is in ascending segment order."""
    costs = np.asarray(costs, np.float64)
    S = len(costs)
    if frame_pair is not None and len(np.unique(frame_pair)) >= world:
        fp = np.asarray(frame_pair)
        ids = np.unique(fp)
        return [np.nonzero(np.isin(fp, ids[r::world]))[0] for r in range(world)]
    load = np.zeros(world)
    out: List[List[int]] = [[] for _ in range(world)]
    for s in np.argsort(-costs, kind="stable"):
        r = int(np.argmin(load))
        out[r].append(int(s))
        load[r] += costs[s]
    return [np.array(sorted(o), dtype=np.int64) for o in out]


def gather_maps(T_local: np.ndarray, seg_ids: np.ndarray, seg_sizes_all: np.ndarray, *, device=None) -> np.ndarray:
    """All-gather one direction's maps.  ``T_local``: this rank's maps concatenated in the order of
    ``seg_ids``; ``seg_sizes_all``: (S,) entries per segment pair of the WHOLE batch.  Returns the whole-batch
    concatenated map on every rank.  One padded all_gather_into_tensor (int64) + a count exchange."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    ids = [None] * world
    dist.all_gather_object(ids, np.asarray(seg_ids, np.int64))
    counts = [int(seg_sizes_all[i].sum()) for i in ids]
    pad = max(counts)
    buf = torch.zeros(pad, dtype=torch.int64, device=dev)
    buf[:len(T_local)] = torch.as_tensor(T_local, dtype=torch.int64).to(dev)
    out = torch.empty(world * pad, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(out, buf)
    out = out.cpu().numpy().reshape(world, pad)
    off = np.concatenate([[0], np.cumsum(seg_sizes_all)]).astype(np.int64)
    full = np.empty(int(off[-1]), np.int64)
    for r in range(world):
        p = 0
        for s in ids[r]:
            n = int(seg_sizes_all[s])
            full[off[s]:off[s] + n] = out[r, p:p + n]
            p += n
    return full


def local_icp_distributed(batch, *, compute=None, gather=True):
    """Shard ``batch`` (a SegmentPairBatch, identical on every rank) over the ranks of the default process
    group, run the local shard, and (optionally) gather whole-batch T21/T12 on every rank.
    ``compute(sub_batch) -> (T21, T12)`` defaults to the CUDA path on this rank's GPU (LOCAL_RANK)."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    shards = shard_segments(segment_costs(batch.off1, batch.off2, batch.lm_off), world, batch.frame_pair)
    mine = shards[rank]
    T21 = T12 = np.zeros(0, np.int64)
    if len(mine):
        sub = batch.take(mine)
        if compute is None:
            from .batch import local_icp_batch
            r = local_icp_batch(sub)
            T21, T12 = r.T21, r.T12
        else:
            T21, T12 = compute(sub)
    if not gather:
        return T21, T12, mine
    n1, n2 = np.diff(batch.off1), np.diff(batch.off2)
    return gather_maps(T21, mine, n2), gather_maps(T12, mine, n1), mine
