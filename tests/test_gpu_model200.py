"""Model 200 and routed runs on the GPU against the CPU oracle: bit for bit, like Model204.

Model 200 is project-defined (the reference names it, README.md:95, and ships no definition) and the routed
scheme is this project's design (parity unpinned against the reference); the pins are
the CPU restatement (oracle/oracle_rk45.c rhs_200, tests/routed_ref.py) and, through it, SciPy
(tests/test_oracle_model200.py)."""
import numpy as np
import pytest
import torch

from oracle import oracle as O
from tiger_hlm_gpu_b200 import Parameters, routing, synthetic
from tests import routed_ref

pytestmark = pytest.mark.gpu

PRM = Parameters(initialStep=1e-6)
OPRM = O.Params.make(initialStep=1e-6)


def inputs(ns, days=1, seed=1, wet_fraction=0.3, links_per_cell=97):
    sp = synthetic.make_spatial_params(ns)
    col, ncells = synthetic.make_cells(ns, links_per_cell=links_per_cell)
    pr, t2m = synthetic.make_forcing_grid(ncells, days)
    rng = np.random.default_rng(seed)
    y0 = np.tile(synthetic.Y0_200, (ns, 1))
    y0[:, 0] = rng.uniform(0.01, 20.0, ns)
    y0[rng.random(ns) < wet_fraction, 2] = 0.01
    return sp, col, pr, t2m, y0


def upload(solver, sp, col, pr, t2m):
    solver.set_model_parameters(200, PRM)
    solver.set_max_attempts(2_000_000)
    solver.set_precision(64)
    solver.upload_spatial_params(sp)
    solver.clear_forcings()
    solver.upload_forcing(0, 1.0, pr)
    solver.upload_forcing(1, 24.0, t2m)
    solver.set_forcing_columns(col)


@pytest.mark.parametrize("ns", [1, 33, 1000])
def test_model200_unrouted_equals_oracle_bit_for_bit(solver, ns):
    sp, col, pr, t2m, y0 = inputs(ns)
    upload(solver, sp, col, pr, t2m)
    solver.route_clear()
    tq = synthetic.hourly_queries(0.0, 1440.0)
    g = solver.run_rk45(200, y0, 0.0, 1440.0, tq)
    o = O.run_rk45(200, OPRM, y0, 0.0, 1440.0, tq, sp=sp, forcing=O.Forcing([pr, t2m], [1.0, 24.0], col=col),
                   device_pow=True, threads=8, max_attempts=2_000_000)
    assert g["n_accept"].sum() > 20 * ns
    for k in ("stiff", "n_accept", "n_reject", "n_jump", "final", "dense"):
        assert np.array_equal(g[k], o[k]), k


def test_model200_days_in_a_resident_session_under_every_schedule(solver):
    """Three day-sized intervals chained on the device (hlm_solve_restart: what chained run_rk45 calls do).  From the
    second day on AUTO deals the links as sorted tiles — 32 links that took the same number of attempts the day
    before — over all links (unrouted: no blocks); the first day, and the "lanes"/"tiles" runs, take the other kernels.
    Same arithmetic per link: the same bits as the CPU oracle chained the same way."""
    ns, days = 3000, 3
    sp, col, pr, t2m, y0 = inputs(ns, days=days, seed=5)
    F = O.Forcing([pr, t2m], [1.0, 24.0], col=col)
    y, acc_o, dense_o = y0, np.zeros(ns, np.int64), []
    for d in range(days):
        tq = synthetic.hourly_queries(d * 1440.0, (d + 1) * 1440.0)
        o = O.run_rk45(200, OPRM, y, d * 1440.0, (d + 1) * 1440.0, tq, sp=sp, forcing=F, device_pow=True, threads=8,
                       max_attempts=2_000_000, reject_limit=routing.ROUTED_REJECT_LIMIT)
        assert (o["stiff"] == 0).all()  # (with the reference's limit of 5 the first day's kinks flag a few links)
        y, acc_o = o["final"], acc_o + o["n_accept"]
        dense_o.append(o["dense"])
    upload(solver, sp, col, pr, t2m)
    solver.route_clear()
    try:
        solver.set_reject_limit(routing.ROUTED_REJECT_LIMIT)
        for schedule in ("auto", "sorted", "lanes", "tiles"):
            solver.set_schedule(schedule)
            for d in range(days):
                tq = synthetic.hourly_queries(d * 1440.0, (d + 1) * 1440.0)
                if d == 0:
                    solver.solve_begin(200, y0, 0.0, 1440.0, tq)
                else:
                    solver.solve_restart(d * 1440.0, (d + 1) * 1440.0, tq)
                solver.solve_window(len(tq), True)
                win = np.zeros((ns, len(tq), 5))
                solver.solve_wait_copy(solver.solve_fetch_window_packed(win))
                assert np.array_equal(win, dense_o[d]), (schedule, d)
            r = solver.solve_end()
            assert np.array_equal(r["final"], y), schedule
            assert np.array_equal(r["n_accept"], acc_o), schedule
    finally:
        solver.set_schedule("auto")
        solver.set_reject_limit(5)


def test_model200_fp32_mode_close_to_fp64(solver):
    sp, col, pr, t2m, y0 = inputs(256)
    upload(solver, sp, col, pr, t2m)
    solver.route_clear()
    solver.set_model_parameters(200, Parameters(initialStep=1e-6, rtol=1e-4, atol=1e-7))
    tq = synthetic.hourly_queries(0.0, 1440.0)
    a = solver.run_rk45(200, y0, 0.0, 1440.0, tq)
    try:
        solver.set_precision(32)
        b = solver.run_rk45(200, y0, 0.0, 1440.0, tq)
    finally:
        solver.set_precision(64)
        solver.set_model_parameters(200, PRM)
    ok = (a["stiff"] == 0) & (b["stiff"] == 0)
    assert ok.mean() > 0.9
    # FP32 has no reference counterpart; the static store sits near a kink of its min() term, where single
    # precision moves a few links visibly: nearly all elements within 0.5 %, every one within 10 %
    close = np.isclose(b["final"][ok], a["final"][ok], rtol=5e-3, atol=5e-6)
    assert close.mean() > 0.97
    np.testing.assert_allclose(b["final"][ok], a["final"][ok], rtol=0.1, atol=5e-6)


def routed_inputs(ns, subbasin, seed=4):
    sp, col, pr, t2m, y0 = inputs(ns, seed=seed, wet_fraction=0.1)
    sp = synthetic.apply_network(sp, synthetic.make_network(ns, subbasin_links=subbasin, seed=seed))
    return sp, col, pr, t2m, y0


@pytest.mark.parametrize("schedule", ["auto", "sorted", "tiles"])
def test_routed_single_rank_equals_oracle_bit_for_bit(solver, schedule):
    """Every way of dealing links to lanes (auto = lane refill; sorted = tiles of links with equal attempt counts in the
    previous interval; tiles = 32 consecutive links) integrates each link with the same arithmetic: same bits."""
    ns, tf, dt, qpi = 3000, 360.0, 15.0, 2
    sp, col, pr, t2m, y0 = routed_inputs(ns, 256)
    p1 = routing.plan(sp["stream"], sp["next_stream"], 1, subbasin_links=256)
    F = O.Forcing([pr, t2m], [1.0, 24.0], col=col)
    fin_o, dense_o, tq_o, na_o = routed_ref.run_single(sp, F, y0, OPRM, p1, 0.0, tf, dt, queries_per_interval=qpi, threads=8)

    upload(solver, sp, col, pr, t2m)
    edges = np.arange(0.0, tf + 1e-9, dt)
    dense_g = np.zeros_like(dense_o)
    try:
        solver.set_schedule(schedule)
        rs = routing.RoutedSolver(solver, 200, p1.ranks[0], 1, 0)
        for i, (a, b) in enumerate(zip(edges[:-1], edges[1:])):
            tq = a + (b - a) * np.arange(1, qpi + 1) / qpi
            if i == 0:
                rs.begin(y0, a, b, tq)
            rs.advance(b, tq)
            win = np.zeros((ns, qpi, 5))
            solver.solve_wait_copy(solver.solve_fetch_window_packed(win))
            dense_g[:, i * qpi:(i + 1) * qpi] = win
        qin, _ = solver.route_peek()
        r = rs.end()
    finally:
        solver.set_schedule("auto")
        solver.set_stream(None)
        solver.set_stiff_fallback(False)
        solver.route_clear()
    assert np.isin(r["stiff"], (0, 3)).all()
    assert np.array_equal(r["n_accept"], na_o)
    assert np.array_equal(r["final"], fin_o)
    assert np.array_equal(dense_g, dense_o)
    assert (qin > 0).sum() > ns // 3          # the inflow really reached the kernel


def test_routing_errors_are_loud(solver):
    from tiger_hlm_gpu_b200 import HlmError
    sp, col, pr, t2m, y0 = inputs(64)
    upload(solver, sp, col, pr, t2m)
    up_ptr = np.arange(65, dtype=np.int64)
    try:
        with pytest.raises(HlmError, match="upstream index out of range"):
            solver.route_set_topology(up_ptr, np.full(64, 64, np.int32))
        with pytest.raises(HlmError, match="not ascending|bad CSR"):
            solver.route_set_topology(up_ptr[::-1].copy(), np.zeros(64, np.int32))
        solver.route_set_topology(up_ptr, np.full(64, -3, np.int32))          # every link fed by halo slot 2
        solver.solve_begin(200, y0, 0.0, 60.0, [60.0])
        with pytest.raises(HlmError, match="3 halo elements"):
            solver.route_gather(None)
        solver.solve_end()
        solver.route_set_topology(np.zeros(11, np.int64), np.zeros(0, np.int32))   # a topology for 10 links
        solver.solve_begin(200, y0, 0.0, 60.0, [60.0])
        with pytest.raises(HlmError, match="another link count"):
            solver.solve_window(1)
        solver.solve_end()
    finally:
        solver.route_clear()


def test_two_ranks_on_one_gpu_equal_one_rank_bit_for_bit():
    """Two device contexts play two ranks; the all-gather is done by hand through a device buffer.  The
    boundary discharge comes from the window kernel's epilogue (send buffer), not from a pack kernel."""
    from tiger_hlm_gpu_b200 import Solver
    ns, tf, dt = 2000, 180.0, 20.0
    sp, col, pr, t2m, y0 = routed_inputs(ns, 128, seed=9)
    F = O.Forcing([pr, t2m], [1.0, 24.0], col=col)
    p1 = routing.plan(sp["stream"], sp["next_stream"], 1, subbasin_links=128)
    fin_o, _, _, na_o = routed_ref.run_single(sp, F, y0, OPRM, p1, 0.0, tf, dt, threads=8)
    p2 = routing.plan(sp["stream"], sp["next_stream"], 2, subbasin_links=128)
    assert p2.n_cut_edges > 0 and p2.max_send > 0
    dev = torch.device("cuda", 0)
    halo = torch.zeros(p2.halo_len, dtype=torch.float64, device=dev)
    ctxs = []
    try:
        for topo in p2.ranks:
            sel = p2.order[topo.lo:topo.hi]
            s = Solver(0)
            upload(s, sp[sel], col[sel], pr, t2m)
            s.route_set_topology(topo.up_ptr, topo.up_idx, topo.send_idx)
            s.set_stiff_fallback(True)
            s.set_reject_limit(routing.ROUTED_REJECT_LIMIT)  # what RoutedSolver sets, and routed_ref uses
            # each context writes straight into its segment of the "gathered" vector
            s.route_set_send_buffer(halo.data_ptr() + 8 * topo.rank * p2.max_send)
            ctxs.append((s, topo, sel))
        edges = np.arange(0.0, tf + 1e-9, dt)
        for i, (a, b) in enumerate(zip(edges[:-1], edges[1:])):
            tq = np.array([b])
            for s, topo, sel in ctxs:
                if i == 0:
                    s.solve_begin(200, y0[sel], a, b, tq)
                    s.route_pack()
                s.synchronize()
            for s, topo, sel in ctxs:
                s.route_gather(halo.data_ptr())
                if i > 0:
                    s.solve_advance(b, tq)
                s.synchronize()                      # every rank has read the halo before anyone overwrites it
            for s, topo, sel in ctxs:
                s.solve_window(1, False)
        fin = np.zeros_like(fin_o)
        na = np.zeros_like(na_o)
        for s, topo, sel in ctxs:
            r = s.solve_end()
            fin[sel], na[sel] = r["final"], r["n_accept"]
    finally:
        for s, _, _ in ctxs:
            s.close()
    assert np.array_equal(na, na_o)
    assert np.array_equal(fin, fin_o)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (the exchange itself); run by tools/routed_check.py under torchrun")
@pytest.mark.parametrize("mode", ["nccl", "peer"])
def test_exchange_two_gpus(mode):
    """Two processes, two GPUs: boundary discharge all-gathered over NCCL, or stored into the peers' halo vectors by
    the integration kernels (CUDA IPC) with a one-element all-reduce as the barrier — bit-identical to one rank."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29641" if mode == "nccl" else "29642", os.path.join(root, "tools", "routed_check.py")]
    out = subprocess.run(cmd + (["--peer"] if mode == "peer" else []), capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "routed_check ok" in out.stdout and ("peer memory" in out.stdout) == (mode == "peer")
