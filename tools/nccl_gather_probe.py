#!/usr/bin/env python3
"""Micro-benchmark of the output all-gather (N ranks, ~1.1 GB total like config 5): grouped broadcasts into uneven views vs
NCCL's native all-gather on padded chunks (+ the compaction copies). torchrun --nproc-per-node N tools/nccl_gather_probe.py"""
import os, torch, torch.distributed as dist
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
total = 276_628_284  # floats of the config-5 soup
counts = [total // world + (37 * r) % 1000 for r in range(world)]
offs = [sum(counts[:r]) for r in range(world + 1)]
local = torch.randn(counts[rank], device="cuda")
out = torch.empty(offs[world], device="cuda")
pad = max(counts)
padded = torch.empty(world * pad, device="cuda")
send = torch.empty(pad, device="cuda")
def t(fn, n=10):
    for _ in range(3): fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
views = [out[offs[r]:offs[r + 1]] for r in range(world)]
a = t(lambda: dist.all_gather(views, local))
def padded_ag():
    send[:counts[rank]].copy_(local)
    dist.all_gather_into_tensor(padded, send)
    for r in range(world):
        out[offs[r]:offs[r + 1]].copy_(padded[r * pad:r * pad + counts[r]])
b = t(padded_ag)
c = t(lambda: dist.all_gather_into_tensor(padded, send))
if rank == 0:
    print("world %d: uneven all_gather %.2f ms | padded all_gather + copies %.2f ms | native all_gather alone %.2f ms (%.0f GB/s received per rank)" % (world, a, b, c, (world - 1) * pad * 4 / c / 1e6))
dist.destroy_process_group()
