"""Utterance sharding across the GPUs of one box and host-side gathering of the results.

The DP is independent per utterance (the reference decodes one utterance per call,
networks/task/forced_alignment.py:154-186), so multi-GPU is pure data parallelism: no collective on
the data path, no NCCL.  One process per GPU; every rank aligns its own shard and only the compact
per-utterance results travel, over the host (gloo) process group.
"""
from __future__ import annotations

import heapq
from typing import List, Sequence

import numpy as np


def shard_by_cost(T: Sequence[int], S: Sequence[int], world_size: int) -> List[np.ndarray]:
    """Longest-processing-time-first assignment of utterances to ranks, balancing sum(T*S).

    Returns ``world_size`` index arrays (ascending inside a shard, so shards keep corpus order).
    """
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    T = np.asarray(T, dtype=np.int64)
    S = np.asarray(S, dtype=np.int64)
    cost = np.maximum(T, 0) * np.maximum(S, 0)
    order = np.argsort(-cost, kind="stable")
    heap = [(0, r) for r in range(world_size)]
    heapq.heapify(heap)
    shards: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        load, r = heapq.heappop(heap)
        shards[r].append(int(i))
        heapq.heappush(heap, (load + int(cost[i]), r))
    return [np.array(sorted(s), dtype=np.int64) for s in shards]


def chunk_by_bytes(T: Sequence[int], S: Sequence[int], max_cells: int) -> List[np.ndarray]:
    """Splits a shard into consecutive chunks of at most ``max_cells`` DP cells (>= 1 utterance each),
    so corpus-scale inputs stream through a bounded workspace (4 B/cell of emissions dominate it)."""
    T = np.asarray(T, dtype=np.int64)
    S = np.asarray(S, dtype=np.int64)
    chunks, cur, acc = [], [], 0
    for i, c in enumerate(T * S):
        if cur and acc + int(c) > max_cells:
            chunks.append(np.array(cur, dtype=np.int64))
            cur, acc = [], 0
        cur.append(i)
        acc += int(c)
    if cur:
        chunks.append(np.array(cur, dtype=np.int64))
    return chunks


def gather_on_host(local_indices: np.ndarray, local_payload: dict, group=None, dst: int = 0):
    """Gathers per-rank result dicts on rank ``dst`` and restores corpus order.

    local_payload maps names to per-utterance python lists / arrays aligned with local_indices.
    Uses ``torch.distributed.gather_object`` on a host (gloo) group: results are a few bytes per
    phoneme, there is nothing for NVLink to do here.
    """
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized():
        return _merge([(np.asarray(local_indices), local_payload)])
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    box = [None] * world if rank == dst else None
    dist.gather_object((np.asarray(local_indices), local_payload), box, dst=dst, group=group)
    return _merge(box) if rank == dst else None


def _merge(parts):
    total = sum(len(idx) for idx, _ in parts)
    names = list(parts[0][1].keys())
    out = {k: [None] * total for k in names}
    for idx, payload in parts:
        for k in names:
            vals = payload[k]
            for j, i in enumerate(idx):
                out[k][int(i)] = vals[j]
    return out
