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
    """Equal-count slabs of the sorted brick list: rank r owns [n*r/world, n*(r+1)/world). bs_convert.cu cuts at equal
    cumulative WEIGHT instead (weight = sub-triangle boxes touching a brick + their mean, so dense regions get shorter
    slabs); both are contiguous in visit order, which is all the gather relies on."""
    return n_bricks * rank // world, n_bricks * (rank + 1) // world


def all_gather_varlen(local, group=None, out=None):
    """local: 1-D tensor (any length, same dtype/device on all ranks) -> (concatenation in rank order, counts list).
    One tiny all-gather of the counts, then every rank's slice is broadcast straight into its final place of `out`
    (reused if large enough): no padding, no staging copies, no concatenation pass."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n = torch.tensor([local.numel()], dtype=torch.int64, device=local.device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    total = sum(counts)
    if out is None or out.numel() < total or out.dtype != local.dtype or out.device != local.device:
        out = torch.empty(max(total, 1), dtype=local.dtype, device=local.device)
    offs = [0]
    for c in counts:
        offs.append(offs[-1] + c)
    out[offs[rank]:offs[rank + 1]].copy_(local)
    works = []
    for r in range(world):
        if counts[r]:
            src = dist.get_global_rank(group, r) if group is not None else r
            works.append(dist.broadcast(out[offs[r]:offs[r + 1]], src=src, group=group, async_op=True))
    for w in works:
        w.wait()
    return out[:total], counts
