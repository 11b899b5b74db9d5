"""Multi-GPU driver logic: independent prompts are the unit of work.  One process per
GPU; rank r takes prompts r::world (config 3: 64 prompts over 8 GPUs = 8 each); there is NO
data-path collective -- torch.distributed is used only for the init barrier, the max-over-ranks
timing reduction and an optional gather of the finished images to rank 0.  The reference itself is
single-device (no NCCL/MPI anywhere); its Swift pipeline batches `imageCount` latents through one
model (Unet.swift:106-122), which is what each rank does with its shard."""
from __future__ import annotations

from typing import List, Sequence


def shard_indices(n_items: int, rank: int, world: int) -> List[int]:
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    return list(range(rank, n_items, world))


def shard_prompts(prompts: Sequence, rank: int, world: int) -> List:
    return [prompts[i] for i in shard_indices(len(prompts), rank, world)]


def chunks(items: Sequence, size: int) -> List[List]:
    """Splits a rank's shard into per-call batches of `size`, padding the last one by repetition
    (results for padded slots are dropped by `unpad`)."""
    out = []
    for i in range(0, len(items), size):
        b = list(items[i:i + size])
        while len(b) < size:
            b.append(b[-1])
        out.append(b)
    return out


def gather_in_order(local_results: Sequence, n_items: int, rank: int, world: int, dist=None):
    """All ranks contribute their shard's results; rank 0 receives them in original prompt order."""
    if world == 1 or dist is None:
        return list(local_results)
    bucket = [None] * world
    dist.all_gather_object(bucket, list(local_results))
    if rank != 0:
        return None
    out = [None] * n_items
    for r, res in enumerate(bucket):
        for i, v in zip(shard_indices(n_items, r, world), res):
            out[i] = v
    return out


def max_over_ranks(value: float, dist=None, device=None) -> float:
    """Timing reduction: the job time is the slowest rank's device time."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
