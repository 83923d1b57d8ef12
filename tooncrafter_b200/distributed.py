"""Multi-GPU plumbing for the clip-parallel path: the reference shards prompts across ranks with no
collective on the data path (scripts/evaluation/inference.py:314-320, scripts/evaluation/ddp_wrapper.py:8-46).

One process per GPU; the only collectives are ONE broadcast of the weights at init and the max-over-ranks of the
timed region.  Backend "nccl" on GPUs (NVLink 5 / NVSwitch), "gloo" in the CPU tests.
"""
from __future__ import annotations

from typing import List

import torch
import torch.distributed as dist


def shard_clips(n_clips: int, rank: int, world: int) -> List[int]:
    """Clip indices of `rank`: contiguous blocks like the reference (inference.py:314-320) when n_clips divides
    evenly, and the remainder spread one-per-rank from rank 0 (the reference silently drops it)."""
    base, rem = divmod(n_clips, world)
    start = rank * base + min(rank, rem)
    return list(range(start, start + base + (1 if rank < rem else 0)))


def clip_seed(base_seed: int, clip_index: int) -> int:
    """Per-clip RNG seed: results do not depend on the world size."""
    return base_seed + clip_index


def broadcast_parameters(module: torch.nn.Module, src: int = 0, bucket_bytes: int = 1 << 30) -> int:
    """Broadcast every parameter and buffer of `module` from `src` in flat buckets; returns bytes sent."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return 0
    tensors = [p.data for p in module.parameters()] + [b.data for b in module.buffers()]
    total = 0
    by_dtype = {}
    for t in tensors:
        by_dtype.setdefault((t.dtype, t.device), []).append(t)
    for (_, _), group in by_dtype.items():
        bucket, size = [], 0
        def flush():
            nonlocal bucket, size, total
            if not bucket:
                return
            flat = torch._utils._flatten_dense_tensors(bucket)
            dist.broadcast(flat, src=src)
            for t, s in zip(bucket, torch._utils._unflatten_dense_tensors(flat, bucket)):
                t.copy_(s)
            total += flat.numel() * flat.element_size()
            bucket, size = [], 0
        for t in group:
            bucket.append(t)
            size += t.numel() * t.element_size()
            if size >= bucket_bytes:
                flush()
        flush()
    return total


def max_over_ranks(value: float, device="cpu") -> float:
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
