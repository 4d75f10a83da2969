"""Multi-GPU plumbing: reference views shard across ranks, one final point all-gather.

The path partitions naturally -- reference views are independent units once the sampler's RNG is keyed per
view (Philox stream = global reference index) -- so each rank processes a contiguous slice of ``refs_local``
and the only exchange is the final gather of the packed points.  Concatenating the ranks' outputs
in rank order reproduces the single-GPU output order exactly.

Works with any ``torch.distributed`` backend (NCCL over NVLink on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_refs: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of the reference list for ``rank``: floor(rank*n/world) .. floor((rank+1)*n/world)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return (rank * n_refs) // world, ((rank + 1) * n_refs) // world


def shard_refs(refs: Sequence[int], rank: Optional[int] = None, world: Optional[int] = None) -> List[int]:
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    lo, hi = shard_bounds(len(refs), rank, world)
    return list(refs[lo:hi])


def all_gather_points(xyz: torch.Tensor, rgb: torch.Tensor, err: torch.Tensor, n_valid: Optional[int] = None,
                      group=None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """Order-preserving all-gather of per-rank packed points.

    ``xyz`` [cap,3], ``rgb`` [cap,3], ``err`` [cap] hold ``n_valid`` points (default: all rows).  Two collectives:
    the counts, then one padded gather of a fused [max_count, 7] f32 buffer (28 B per point).  Returns the
    concatenation in rank order plus the per-rank counts.
    """
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        n = xyz.shape[0] if n_valid is None else int(n_valid)
        return xyz[:n], rgb[:n], err[:n], torch.tensor([n], dtype=torch.int64, device=xyz.device)
    world = dist.get_world_size(group)
    n = xyz.shape[0] if n_valid is None else int(n_valid)
    dev = xyz.device
    counts = torch.zeros(world, dtype=torch.int64, device=dev)
    mine = torch.tensor([n], dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(counts, mine, group=group)
    cmax = int(counts.max().item())
    fused = torch.zeros((max(cmax, 1), 7), dtype=torch.float32, device=dev)
    fused[:n, 0:3] = xyz[:n]
    fused[:n, 3:6] = rgb[:n]
    fused[:n, 6] = err[:n]
    gathered = torch.empty((world * max(cmax, 1) * 7,), dtype=torch.float32, device=dev)
    dist.all_gather_into_tensor(gathered, fused.reshape(-1), group=group)
    gathered = gathered.view(world, max(cmax, 1), 7)
    parts = [gathered[r, : int(counts[r].item())] for r in range(world)]
    allp = torch.cat(parts, dim=0) if parts else fused[:0]
    return allp[:, 0:3].contiguous(), allp[:, 3:6].contiguous(), allp[:, 6].contiguous(), counts
