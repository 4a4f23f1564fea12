"""Partition of independent links over ranks (one process per GPU) and the bench's reductions.

The reference scatters contiguous row chunks from MPI rank 0 to ranks 1..size-1
(main.cpp:275-307: baseChunk = n / workers, the first `remainder` workers get one more row) and never
gathers: each rank writes its own files (main.cpp:796-797).  Links are independent (no equation reads
next_stream), so the partition needs no data-path collective; the only collectives are the
bench's MAX over ranks of the device time and SUM of the step counts.
"""
from __future__ import annotations


def shard_range(n_total: int, world: int, rank: int) -> tuple[int, int]:
    """Rows [lo, hi) of rank `rank` under the reference's chunk rule (main.cpp:275-307)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def reduce_timing(ms_local: float, sums_local, dist=None, device=None):
    """(max over ranks of ms, element-wise sum over ranks of sums).  dist=None: single process."""
    import torch
    t = torch.tensor([ms_local], dtype=torch.float64, device=device)
    s = torch.tensor([float(x) for x in sums_local], dtype=torch.float64, device=device)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
    return float(t.item()), [float(x) for x in s.tolist()]


def bind_to_gpu_numa_node(gpu_index: int) -> str:
    """Pin this process to the CPUs of the NUMA node its GPU hangs off, so that pinned host buffers are
    allocated next to the GPU's PCIe root and the D2H stream of one rank does not cross sockets (with several
    ranks per box the e2e path moves ~10 GB per step per GPU).  Best effort: returns what was done."""
    import os
    import subprocess
    try:
        bus = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(gpu_index)],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        if not bus:
            return "numa: no bus id"
        if len(bus.split(":")[0]) == 8:      # nvidia-smi prints an 8-digit PCI domain, sysfs uses 4
            bus = bus[4:]
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip())
        if node < 0:
            return "numa: node unknown"
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return f"numa: node {node} has no allowed cpu"
        os.sched_setaffinity(0, allowed)
        return f"numa: gpu {gpu_index} ({bus}) -> node {node}, {len(allowed)} cpus"
    except Exception as ex:  # a missing sysfs entry or tool must never stop a run
        return f"numa: not bound ({type(ex).__name__})"
