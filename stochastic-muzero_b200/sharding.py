"""Multi-GPU plumbing for the search path: trees are independent, so the batch is sharded across ranks
(one process per GPU) and the ONLY collective on the path is the broadcast of the packed weight blob
after a weight update.  Replaces the reference's CPU `ray` fan-out of whole games
(self_play.py:240-256) and the per-forward DataParallel replicate/gather (muzero_model.py:360-367).

Works with any torch.distributed backend: NCCL over NVLink on the GPU box, gloo in the CPU tests.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.distributed as dist

from .weights import ModelShape, blob_layout


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced partition of `total` trees: rank r owns [lo, hi)."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def broadcast_weights(blob: Optional[torch.Tensor], shape: ModelShape, src: int = 0,
                      device: Optional[torch.device] = None) -> torch.Tensor:
    """Every rank returns the src rank's fp32 blob.  Non-src ranks may pass None."""
    _, total = blob_layout(shape)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        assert blob is not None
        return blob
    if dist.get_rank() == src:
        assert blob is not None and blob.numel() == total, "blob does not match the model shape"
        buf = blob.to(device=device or blob.device, dtype=torch.float32).contiguous()
    else:
        buf = torch.empty(total, dtype=torch.float32, device=device or "cpu")
    dist.broadcast(buf, src=src)
    return buf


def gather_roots(local: Dict[str, torch.Tensor], total: int) -> Dict[str, torch.Tensor]:
    """Optional: assemble the per-rank root statistics (visits [b,A], root_values [b], ...) into the
    global tree order on every rank.  Shards may be ragged (shard_range), so pad to the largest."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    sizes = [shard_range(total, r, world) for r in range(world)]
    widest = max(hi - lo for lo, hi in sizes)
    out = {}
    for key, t in local.items():
        pad = torch.zeros((widest,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        pad[:t.shape[0]] = t
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad)
        out[key] = torch.cat([p[:hi - lo] for p, (lo, hi) in zip(parts, sizes)], dim=0)
    return out
