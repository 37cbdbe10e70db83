"""Multi-GPU plumbing for the brick-sharded remesh (one process per GPU, torch.distributed).

The data path has no collective: every rank converts the replicated mesh and keeps its slab of the sorted brick list
(`bs_mesh_to_volume_sharded`). The one exchange is the final all-gather of the per-rank triangle buffers; ranks hold
different counts, so it is an all-gather of counts followed by a padded all-gather that is trimmed on arrival.
Concatenating the per-rank buffers in rank order reproduces the single-GPU output order exactly (slabs are contiguous
in the reference's leaf visit order).
"""
import torch
import torch.distributed as dist


def slab_bounds(n_bricks, rank, world):
    """Same arithmetic as bs_convert.cu: rank r owns bricks [n*r/world, n*(r+1)/world) of the sorted list."""
    return n_bricks * rank // world, n_bricks * (rank + 1) // world


def all_gather_varlen(local, group=None):
    """local: 1-D tensor (any length, same dtype/device on all ranks) -> (concatenation in rank order, counts list)."""
    world = dist.get_world_size(group)
    n = torch.tensor([local.numel()], dtype=torch.int64, device=local.device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    m = max(counts) if counts else 0
    padded = torch.zeros(max(m, 1), dtype=local.dtype, device=local.device)
    padded[: local.numel()] = local
    bufs = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(bufs, padded, group=group)
    return torch.cat([b[:c] for b, c in zip(bufs, counts)]), counts
