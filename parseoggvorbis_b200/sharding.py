"""Multi-GPU sharding of the synthesis stage: independent streams / files, no collective on the data path.

Streams share nothing (reference: std::map<serial, VorbisStream>, src/ParseOggVorbis.hpp:1387; per-stream
VorbisStreamDecodeState, hpp:1122), so rank r of W decodes its own subset and only two scalars ever cross ranks:
the number of PCM values it produced and the time it took (throughput = sum of values / max of times).
torch.distributed is plumbing here (barrier + two tiny all-reduces); the backend is NCCL on GPUs and gloo in the CPU tests.
"""
from typing import List, Sequence, Tuple

import numpy as np


def shard_range(n_items: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous balanced split of n_items over `world` ranks: ranks < n_items % world get one more."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad world/rank %d/%d" % (rank, world))
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_by_cost(costs: Sequence[int], world: int) -> List[List[int]]:
    """Greedy longest-processing-time assignment of items (e.g. files by packet count) to ranks; deterministic."""
    order = sorted(range(len(costs)), key=lambda i: (-int(costs[i]), i))
    load = [0] * world
    out: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += int(costs[i])
    for r in range(world):
        out[r].sort()
    return out


def aggregate_throughput(units_local: int, ms_local: float, dist=None, device=None) -> Tuple[int, float]:
    """-> (units summed over ranks, milliseconds = max over ranks). `dist` = torch.distributed or None (1 rank)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return int(units_local), float(ms_local)
    import torch
    t = torch.tensor([float(ms_local)], dtype=torch.float64, device=device)
    u = torch.tensor([int(units_local)], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return int(u.item()), float(t.item())


def merge_checksums(local: np.ndarray, dist=None) -> np.ndarray:
    """All ranks' per-stream checksums, concatenated in rank order (shards are contiguous, so this is stream order)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return np.asarray(local)
    parts = [None] * dist.get_world_size()
    dist.all_gather_object(parts, np.asarray(local))
    return np.concatenate(parts)
