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


# ---- routed runs: partition by sub-basin, boundary links exchanged with one all-gather per interval ----

def _routed_case(ns=120, seed=5):
    from tests.test_oracle_model200 import network_case
    sp, down, rain, temp, pr, t2m, y0 = network_case(ns=ns, seed=seed)
    return sp, pr, t2m, y0


def _routed_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    from tests import routed_ref
    from tiger_hlm_gpu_b200 import routing
    sp, pr, t2m, y0 = _routed_case()
    p = routing.plan(sp["stream"], sp["next_stream"], world, subbasin_links=10)
    topo = p.ranks[rank]
    sel = p.order[topo.lo:topo.hi]
    rk = routed_ref.OracleRank(topo, sp[sel], O.Forcing([pr[:, sel], t2m[:, sel]], [1.0, 24.0]), y0[sel],
                               O.Params.make(initialStep=1e-6), 0.0, threads=1)
    halo = torch.zeros(p.halo_len, dtype=torch.float64)
    for b in np.arange(15.0, 120.0 + 1e-9, 15.0):
        dist.all_gather_into_tensor(halo, torch.from_numpy(rk.send(p.max_send)))   # the exchange of RoutedSolver
        rk.gather(halo.numpy())
        rk.advance(b, np.array([b]))
    np.savez(os.path.join(out_dir, f"routed{rank}.npz"), final=rk.y, n_accept=rk.na, sel=sel, cut=p.n_cut_edges)
    dist.barrier()
    dist.destroy_process_group()


def test_two_gloo_ranks_reproduce_the_single_rank_routed_run(tmp_path):
    world = 2
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_routed_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    from oracle import oracle as O
    from tests import routed_ref
    from tiger_hlm_gpu_b200 import routing
    sp, pr, t2m, y0 = _routed_case()
    p1 = routing.plan(sp["stream"], sp["next_stream"], 1, subbasin_links=10)
    fin, _, _, na = routed_ref.run_single(sp, O.Forcing([pr, t2m], [1.0, 24.0]), y0, O.Params.make(initialStep=1e-6), p1,
                                          0.0, 120.0, 15.0, threads=1)
    parts = [np.load(tmp_path / f"routed{r}.npz") for r in range(world)]
    assert int(parts[0]["cut"]) > 0                       # the partition really cuts the network
    got_fin, got_na = np.zeros_like(fin), np.zeros_like(na)
    for q in parts:
        got_fin[q["sel"]], got_na[q["sel"]] = q["final"], q["n_accept"]
    assert np.array_equal(got_fin, fin) and np.array_equal(got_na, na)
