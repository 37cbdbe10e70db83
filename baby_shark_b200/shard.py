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


def all_gather_varlen(local, group=None, out=None, counts=None):
    """local: 1-D tensor (any length, same dtype/device on all ranks) -> (concatenation in rank order, counts list).
    `counts` (per-rank lengths) may be passed when the caller already knows them (a steady workload: the sizes of the
    previous step) -- then there is no count exchange and no host synchronisation at all. Every rank's slice travels
    straight into its final place of `out` (reused if large enough): NCCL takes the uneven all-gather as ONE grouped launch
    of broadcasts; gloo (CPU tests) gets one broadcast per rank. No padding, no staging copies, no concatenation pass."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if counts is None:
        n = torch.tensor([local.numel()], dtype=torch.int64, device=local.device)
        got = [torch.zeros_like(n) for _ in range(world)]
        dist.all_gather(got, n, group=group)
        counts = [int(c.item()) for c in got]
    assert counts[rank] == local.numel(), "stale counts: %d expected, %d local" % (counts[rank], local.numel())
    total = sum(counts)
    if out is None or out.numel() < total or out.dtype != local.dtype or out.device != local.device:
        out = torch.empty(max(total, 1), dtype=local.dtype, device=local.device)
    offs = [0]
    for c in counts:
        offs.append(offs[-1] + c)
    views = [out[offs[r]:offs[r + 1]] for r in range(world)]
    if dist.get_backend(group) == "nccl" and all(counts):
        dist.all_gather(views, local, group=group)
    else:
        views[rank].copy_(local)
        works = []
        for r in range(world):
            if counts[r]:
                src = dist.get_global_rank(group, r) if group is not None else r
                works.append(dist.broadcast(views[r], src=src, group=group, async_op=True))
        for w in works:
            w.wait()
    return out[:total], counts
