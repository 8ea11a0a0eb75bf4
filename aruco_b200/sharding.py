"""Frame sharding across GPUs: frames of a batch are independent, so the path shards into
contiguous frame ranges, one per rank, with NO collective on the data path; results (a few KB per frame)
are gathered on the host in frame order."""
from __future__ import annotations

from typing import List, Sequence, Tuple


def shard_range(n_frames: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous split; the first n_frames % world ranks get one extra frame. Returns (start, count)."""
    if world < 1 or not (0 <= rank < world) or n_frames < 0:
        raise ValueError("bad shard request")
    base, extra = divmod(n_frames, world)
    start = rank * base + min(rank, extra)
    return start, base + (1 if rank < extra else 0)


def gather_in_frame_order(local_results: Sequence, n_frames: int, world: int, rank: int, group=None) -> List:
    """All ranks contribute their shard's per-frame results; every rank gets the full list in frame order.
    Uses torch.distributed object gathering (host side, works with gloo and nccl groups)."""
    start, count = shard_range(n_frames, world, rank)
    if len(local_results) != count:
        raise ValueError("rank %d holds %d results for a shard of %d frames" % (rank, len(local_results), count))
    if world == 1:
        return list(local_results)
    import torch.distributed as dist
    parts = [None] * world
    dist.all_gather_object(parts, (start, list(local_results)), group=group)
    out = [None] * n_frames
    for s, res in parts:
        out[s:s + len(res)] = res
    if any(r is None for r in out):
        raise RuntimeError("frame results missing after the gather")
    return out
