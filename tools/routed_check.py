"""torchrun --nproc-per-node N tools/routed_check.py — the routed run over N GPUs with the NCCL all-gather,
checked against the CPU oracle's single-rank run (tests/routed_ref.py) bit for bit.  Test tool."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from tests import routed_ref  # noqa: E402
from tests.test_gpu_model200 import OPRM, routed_inputs, upload  # noqa: E402
from tiger_hlm_gpu_b200 import Solver, routing  # noqa: E402


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ns, tf, dt, sub = 20000, 240.0, 15.0, 512
    sp, col, pr, t2m, y0 = routed_inputs(ns, sub, seed=21)
    p = routing.plan(sp["stream"], sp["next_stream"], world, subbasin_links=sub)
    topo = p.ranks[rank]
    sel = p.order[topo.lo:topo.hi]
    s = Solver(local)
    upload(s, sp[sel], col[sel], pr, t2m)
    exchange = "peer" if "--peer" in sys.argv else "nccl"
    rs = routing.RoutedSolver(s, 200, topo, world, p.max_send, dist if world > 1 else None, exchange=exchange)
    edges = np.arange(0.0, tf + 1e-9, dt)
    for i, (a, b) in enumerate(zip(edges[:-1], edges[1:])):
        tq = np.array([b])
        if i == 0:
            rs.begin(y0[sel], a, b, tq)
        rs.advance(b, tq, want_dense=False)
    r = rs.end()
    fin = torch.zeros(ns, 5, dtype=torch.float64, device="cuda")
    na = torch.zeros(ns, dtype=torch.float64, device="cuda")
    fin[torch.as_tensor(sel, device="cuda")] = torch.as_tensor(r["final"], device="cuda")
    na[torch.as_tensor(sel, device="cuda")] = torch.as_tensor(r["n_accept"].astype(np.float64), device="cuda")
    if world > 1:
        dist.all_reduce(fin)
        dist.all_reduce(na)
    if rank == 0:
        p1 = routing.plan(sp["stream"], sp["next_stream"], 1, subbasin_links=sub)
        F = O.Forcing([pr, t2m], [1.0, 24.0], col=col)
        fin_o, _, _, na_o = routed_ref.run_single(sp, F, y0, OPRM, p1, 0.0, tf, dt, threads=os.cpu_count() or 4)
        na_g = na.cpu().numpy().astype(np.int64)
        if not np.array_equal(na_g, na_o):
            bad = np.flatnonzero(na_g != na_o)
            inv = p.inverse_order()
            print("mismatching links", bad.size, bad[:10], "planned pos", inv[bad[:10]], "ranges", p.ranges,
                  "gpu", na_g[bad[:10]], "oracle", na_o[bad[:10]], "stiff codes", np.unique(r["stiff"], return_counts=True))
        assert np.array_equal(na_g, na_o), "accepted-step counts differ"
        assert np.array_equal(fin.cpu().numpy(), fin_o), "final states differ"
        print(f"routed_check ok: world {world}, {ns} links, {p.n_subbasins} sub-basins, {p.n_cut_edges} cut edges, "
              f"{rs.exchanges} x " + ("barrier (boundary discharge stored into peer memory by the kernels)" if rs.peer else
                                      f"all-gather of {p.halo_len} doubles") + ", bit-identical to the single-rank CPU oracle")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
