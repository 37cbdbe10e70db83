"""CPU tests of the multi-GPU host plumbing: world_size-2 gloo all-gather of ragged triangle buffers, slab bounds."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from baby_shark_b200.shard import all_gather_varlen, slab_bounds


def test_slab_bounds_cover_and_are_disjoint():
    for n in (0, 1, 7, 281195):
        for world in (1, 2, 4, 8):
            b = [slab_bounds(n, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            assert max(hi - lo for lo, hi in b) - min(hi - lo for lo, hi in b) <= 1


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    local = torch.arange(3 * (rank + 2) + (5 if rank else 0), dtype=torch.float32) + 100 * rank  # ragged lengths
    cat, counts = all_gather_varlen(local)
    expect = torch.cat([torch.arange(3 * (r + 2) + (5 if r else 0), dtype=torch.float32) + 100 * r for r in range(world)])
    ok = torch.equal(cat, expect) and counts == [3 * (r + 2) + (5 if r else 0) for r in range(world)]
    empty, _ = all_gather_varlen(torch.zeros(0) if rank == 0 else torch.ones(4))
    ok = ok and torch.equal(empty, torch.ones(4))
    out[rank] = ok
    dist.destroy_process_group()


def test_all_gather_varlen_gloo_world2():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, 29731, out), nprocs=world, join=True)
    assert all(out[r] for r in range(world))
