"""torchrun --nproc-per-node N tools/d2h_probe.py — aggregate device-to-host copy rate of N ranks copying at once
(pinned destination, 2 GiB per copy), the ceiling of bench.py's e2e arm at N GPUs.  Measurement tool."""
import os
import time

import torch
import torch.distributed as dist

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 1 << 28  # doubles: 2 GiB
src = torch.zeros(n, dtype=torch.float64, device="cuda")
dst = torch.zeros(n, dtype=torch.float64).pin_memory()
for _ in range(2):
    dst.copy_(src, non_blocking=True)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
reps = 5
for _ in range(reps):
    dst.copy_(src, non_blocking=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
rate = torch.tensor([reps * n * 8 / dt / 1e9], device="cuda")
lo = rate.clone()
if world > 1:
    dist.all_reduce(rate)
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"d2h_probe: {world} rank(s) copying concurrently: {rate.item():.1f} GB/s aggregate, slowest rank {lo.item():.1f} GB/s")
if world > 1:
    dist.destroy_process_group()
