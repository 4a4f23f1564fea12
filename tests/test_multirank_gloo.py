"""The N > 1 host logic on CPU: world_size-2 gloo processes shard links with the reference's chunk
rule, integrate their shard (the CPU oracle stands in for the device here) and reduce the bench's
timing/step counters.  Sharded results must equal the unsharded ones bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tiger_hlm_gpu_b200 import synthetic
from tiger_hlm_gpu_b200.sharding import reduce_timing, shard_range


def test_shard_range_follows_reference_chunk_rule():
    # main.cpp:275-307: 10 rows over 3 workers -> 4,3,3
    assert [shard_range(10, 3, r) for r in range(3)] == [(0, 4), (4, 7), (7, 10)]
    assert [shard_range(8, 8, r) for r in range(8)] == [(r, r + 1) for r in range(8)]
    assert shard_range(3, 4, 3) == (3, 3)  # more ranks than rows: empty shard
    for n, w in ((41274, 8), (10_000_000, 8), (7, 2)):
        cuts = [shard_range(n, w, r) for r in range(w)]
        assert cuts[0][0] == 0 and cuts[-1][1] == n and all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))


def _worker(rank, world, port, ns, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    sp = synthetic.make_spatial_params(ns)
    col, ncells = synthetic.make_cells(ns, 16)
    pr, t2m = synthetic.make_forcing_grid(ncells, 1)
    y0 = synthetic.make_y0(ns, 0.3)
    lo, hi = shard_range(ns, world, rank)
    tq = synthetic.hourly_queries(0.0, 1440.0)
    r = O.run_rk45(204, O.Params.make(initialStep=1e-6), y0[lo:hi], 0.0, 1440.0, tq, sp=sp[lo:hi],
                   forcing=O.Forcing([pr, t2m], [1.0, 24.0], col=col[lo:hi]))
    ms, sums = reduce_timing(10.0 + rank, [r["n_accept"].sum(), hi - lo], dist)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), final=r["final"], dense=r["dense"], n_accept=r["n_accept"],
             ms=ms, sums=sums)
    dist.barrier()
    dist.destroy_process_group()


def test_two_gloo_ranks_reproduce_the_unsharded_run(tmp_path):
    ns, world = 75, 2
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(world, port, ns, str(tmp_path)), nprocs=world, join=True)
    from oracle import oracle as O
    sp = synthetic.make_spatial_params(ns)
    col, ncells = synthetic.make_cells(ns, 16)
    pr, t2m = synthetic.make_forcing_grid(ncells, 1)
    whole = O.run_rk45(204, O.Params.make(initialStep=1e-6), synthetic.make_y0(ns, 0.3), 0.0, 1440.0,
                       synthetic.hourly_queries(0.0, 1440.0), sp=sp, forcing=O.Forcing([pr, t2m], [1.0, 24.0], col=col))
    parts = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    for key in ("final", "dense", "n_accept"):
        assert np.array_equal(np.concatenate([p[key] for p in parts]), whole[key])
    for p in parts:
        assert float(p["ms"]) == 11.0                       # MAX over ranks
        assert p["sums"].tolist() == [float(whole["n_accept"].sum()), float(ns)]  # SUM over ranks
