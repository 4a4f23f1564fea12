"""CPU pins of Model 200 (project-defined: the reference names it, README.md:95, but ships no definition —
"parity unpinned" against the reference) and of the routed scheme.

1. The C restatement (oracle/oracle_rk45.c rhs_200) against the same equations written in plain Python and
   integrated by SciPy at much tighter tolerances: within the solver tolerance 10 * (atol + rtol |y|).
2. The routed scheme (inflow held over a coupling interval) against SciPy on the fully coupled network: the
   error falls about linearly with the coupling interval (first-order coupling).
3. Partition planning invariants, and a 3-rank run with a simulated exchange equal to the 1-rank run bit for bit.
"""
import numpy as np
from scipy.integrate import solve_ivp

from oracle import oracle as O
from tiger_hlm_gpu_b200 import routing, synthetic
from tests import routed_ref

RTOL, ATOL = 1e-6, 1e-9
PRM = O.Params.make(initialStep=1e-6, rtol=RTOL, atol=ATOL)


def rhs200_py(y, p, rain, T, q_in):
    q, h_stat, h_surf, h_grav, h_aq = y
    x2 = max(0.0, rain + h_stat - p["Hu"])
    d1 = rain - x2 - (h_stat / p["Hu"]) * min(0.1 * T, h_stat)
    x3 = min(x2, p["infil"])
    w = min(1.0, (1.0 / p["n_mann"]) * max(h_surf, 0.0) ** (2.0 / 3.0) * np.sqrt(p["slope"]) * p["L"] / p["A_h"] * 60.0)
    out_surf = h_surf * w
    x4 = min(x3, p["perco"])
    out_grav = h_grav / p["alpha3"] if p["alpha3"] >= 1 else 0.0
    out_aq = h_aq / p["alpha4"] if p["alpha4"] >= 1 else 0.0
    runoff = out_surf + out_grav + out_aq
    invtau = 19.8 / (800.0 * p["L"] * p["A_h"] ** 0.1)
    dq = invtau * max(q, 1e-6) ** 0.2 * (runoff * p["A_h"] * 1e6 / 60.0 + q_in - q)
    return [dq, d1, x2 - x3 - out_surf, x3 - x4 - out_grav, x4 - out_aq]


def small_case(ns=12, seed=3, wet=True):
    rng = np.random.default_rng(seed)
    sp = synthetic.make_spatial_params(ns)
    rain = (rng.uniform(0.0, 3.0, ns) * synthetic.C1).astype(np.float32)
    temp = rng.uniform(2, 15, ns).astype(np.float32)
    pr, t2m = np.tile(rain, (24, 1)), np.tile(temp, (1, 1))
    y0 = np.tile(synthetic.Y0_200, (ns, 1))
    y0[:, 0] = rng.uniform(0.05, 5.0, ns)
    if wet:
        y0[::2, 2] = rng.uniform(1e-4, 2e-3, (ns + 1) // 2)
    return sp, rain, temp, pr, t2m, y0


def test_model200_unrouted_within_tolerance_of_scipy():
    """Global error against LSODA at 1e-12: bounded by 200 x the local tolerance (atol + rtol |y|) the
    controller enforces per step — the stores decay, so relative errors made early do not shrink — and it
    falls with the tolerance (checked at rtol 1e-6 and 1e-9)."""
    sp, rain, temp, pr, t2m, y0 = small_case()
    ns = len(sp)
    tq = 60.0 * np.arange(1, 25)
    qin = np.linspace(0.0, 2.0, ns)
    refs = []
    for s in range(ns):
        sol = solve_ivp(lambda t, y: rhs200_py(y, sp[s], float(rain[s]), float(temp[s]), qin[s]), (0.0, 1440.0), y0[s],
                        method="LSODA", rtol=1e-12, atol=1e-15, t_eval=tq)
        refs.append(sol.y.T)
    ref = np.stack(refs)
    for rtol, atol in ((1e-6, 1e-9), (1e-9, 1e-12)):
        prm = O.Params.make(initialStep=1e-6, rtol=rtol, atol=atol)
        for dev in (False, True):
            r = O.run_rk45(200, prm, y0, 0.0, 1440.0, tq, sp=sp, forcing=O.Forcing([pr, t2m], [1.0, 24.0]), inflow=qin,
                           device_pow=dev, max_attempts=1_000_000)
            # at the tight tolerance the reference's stiffness floor (h < (tf - t0) * 1e-6 after a rejection,
            # rk45_kernel.cu:160) abandons links whose stores cross a kink of the min/max terms: skip those
            ok = r["stiff"] == 0
            assert ok.all() if rtol == 1e-6 else ok.sum() >= 3
            assert (r["n_accept"][ok] > 10).all()
            bound = 200.0 * (atol + rtol * np.abs(ref))
            assert (np.abs(r["dense"] - ref) <= bound)[ok].all()
            assert (np.abs(r["final"] - ref[:, -1]) <= bound[:, -1])[ok].all()


def test_inflow_enters_only_the_channel_equation():
    sp, rain, temp, pr, t2m, y0 = small_case(ns=4)
    tq = np.array([30.0, 60.0])
    F = O.Forcing([pr, t2m], [1.0, 24.0])
    a = O.run_rk45(200, PRM, y0, 0.0, 60.0, tq, sp=sp, forcing=F)
    b = O.run_rk45(200, PRM, y0, 0.0, 60.0, tq, sp=sp, forcing=F, inflow=np.full(4, 3.0))
    z = O.run_rk45(200, PRM, y0, 0.0, 60.0, tq, sp=sp, forcing=F, inflow=np.zeros(4))
    assert np.array_equal(a["final"], z["final"])                    # no inflow array == zero inflow
    assert (b["final"][:, 0] > a["final"][:, 0]).all()               # more water in, more discharge
    np.testing.assert_allclose(b["final"][:, 1:], a["final"][:, 1:], rtol=1e-5, atol=1e-8)  # hillslope does not see the channel


def network_case(ns=40, seed=11):
    sp, rain, temp, pr, t2m, y0 = small_case(ns=ns, seed=seed, wet=False)
    down = synthetic.make_network(ns, subbasin_links=10, seed=seed)
    sp = synthetic.apply_network(sp, down)
    return sp, down, rain, temp, pr, t2m, y0


def test_routed_scheme_converges_to_the_coupled_solution():
    sp, down, rain, temp, pr, t2m, y0 = network_case()
    ns = len(sp)
    tf = 240.0

    def coupled(t, Y):
        y = Y.reshape(ns, 5)
        qin = np.zeros(ns)
        np.add.at(qin, down[down >= 0], y[down >= 0, 0])
        return np.concatenate([rhs200_py(y[s], sp[s], float(rain[s]), float(temp[s]), qin[s]) for s in range(ns)])

    ref = solve_ivp(coupled, (0.0, tf), y0.ravel(), method="LSODA", rtol=1e-10, atol=1e-13).y[:, -1].reshape(ns, 5)
    F = O.Forcing([pr, t2m], [1.0, 24.0])
    errs = []
    for dt in (30.0, 7.5, 1.875):
        p1 = routing.plan(sp["stream"], sp["next_stream"], 1, subbasin_links=10)
        final, _, _, _ = routed_ref.run_single(sp, F, y0, PRM, p1, 0.0, tf, dt, device_pow=False)
        errs.append(np.max(np.abs(final[:, 0] - ref[:, 0]) / np.abs(ref[:, 0])))
        np.testing.assert_allclose(final[:, 1:], ref[:, 1:], rtol=1e-4)  # hillslope stores are uncoupled
    assert errs[0] > errs[1] > errs[2]
    assert errs[2] < 0.02 and errs[2] < errs[0] / 6.0   # ~first order: 16x shorter interval, > 6x smaller error


def test_plan_invariants():
    ns = 5000
    sp = synthetic.apply_network(synthetic.make_spatial_params(ns), synthetic.make_network(ns, subbasin_links=250))
    down = routing.downstream_index(sp["stream"], sp["next_stream"])
    for world in (1, 2, 3, 8):
        p = routing.plan(sp["stream"], sp["next_stream"], world, subbasin_links=250)
        assert sorted(p.order.tolist()) == list(range(ns))
        assert p.ranges[0][0] == 0 and p.ranges[-1][1] == ns and all(a[1] == b[0] for a, b in zip(p.ranges, p.ranges[1:]))
        sizes = np.array([hi - lo for lo, hi in p.ranges])
        assert sizes.max() - sizes.min() <= 2 * 2 * 250          # balanced to about a sub-basin
        inv = p.inverse_order()
        owner = np.searchsorted([hi for _, hi in p.ranges], inv, side="right")
        n_up = 0
        for r, topo in enumerate(p.ranks):
            assert topo.up_ptr[0] == 0 and topo.up_ptr[-1] == topo.up_idx.size
            for i in range(topo.n_local):
                ups = topo.up_idx[topo.up_ptr[i]:topo.up_ptr[i + 1]]
                orig = []
                for e in ups:
                    if e >= 0:
                        orig.append(p.order[topo.lo + e])
                    else:
                        slot = -(e + 1)
                        src, k = divmod(slot, p.max_send)
                        assert src != r
                        orig.append(p.order[p.ranks[src].lo + p.ranks[src].send_idx[k]])
                assert orig == sorted(orig)                                     # ascending original index
                assert all(down[o] == p.order[topo.lo + i] for o in orig)       # every entry really drains here
                n_up += len(orig)
            for k in topo.send_idx:                                            # boundary links drain off-rank
                assert owner[down[p.order[topo.lo + k]]] != r
        assert n_up == (down >= 0).sum()
        if world == 1:
            assert p.max_send == 0 and p.n_cut_edges == 0


def test_plan_leaves_no_rank_empty_and_rejects_too_few_subbasins():
    import pytest
    # three sub-basins of very unlike size over three ranks: a share-based cut alone would leave rank 2 empty
    down = np.concatenate([[-1], np.arange(9), [-1], np.arange(10, 19), [-1], np.arange(20, 99)])
    stream = 1000 + np.arange(100)
    nxt = np.where(down >= 0, stream[np.maximum(down, 0)], 0)
    p = routing.plan(stream, nxt, 3, subbasin_links=1000)
    assert p.n_subbasins == 3 and [hi - lo for lo, hi in p.ranges] == [10, 10, 80]
    with pytest.raises(ValueError, match="cannot be dealt"):
        routing.plan(stream, nxt, 4, subbasin_links=1000)
    # a network without a single edge: every link is its own sub-basin, nothing to exchange
    p = routing.plan(stream, np.zeros(100, np.int64), 2)
    assert p.n_cut_edges == 0 and p.max_send == 0 and all(t.up_idx.size == 0 for t in p.ranks)
    assert [hi - lo for lo, hi in p.ranges] == [50, 50]


def test_three_ranks_with_a_simulated_exchange_equal_one_rank_bit_for_bit():
    sp, down, rain, temp, pr, t2m, y0 = network_case(ns=60, seed=5)
    F = O.Forcing([pr, t2m], [1.0, 24.0])
    tf, dt = 120.0, 15.0
    p1 = routing.plan(sp["stream"], sp["next_stream"], 1, subbasin_links=10)
    final1, dense1, tq1, na1 = routed_ref.run_single(sp, F, y0, PRM, p1, 0.0, tf, dt, queries_per_interval=2)
    p3 = routing.plan(sp["stream"], sp["next_stream"], 3, subbasin_links=10)
    assert p3.n_cut_edges > 0
    ranks = []
    for topo in p3.ranks:
        sel = p3.order[topo.lo:topo.hi]
        Fr = O.Forcing([pr[:, sel], t2m[:, sel]], [1.0, 24.0])
        ranks.append(routed_ref.OracleRank(topo, sp[sel], Fr, y0[sel], PRM, 0.0))
    edges = np.arange(0.0, tf + 1e-9, dt)
    dense3 = np.zeros_like(dense1)
    qi = 0
    for a, b in zip(edges[:-1], edges[1:]):
        halo = np.concatenate([rk.send(p3.max_send) for rk in ranks])      # what the all-gather produces
        tq = a + (b - a) * np.arange(1, 3) / 2
        for rk in ranks:
            rk.gather(halo)
        for topo, rk in zip(p3.ranks, ranks):
            r = rk.advance(b, tq)
            dense3[p3.order[topo.lo:topo.hi], qi:qi + 2] = r["dense"]
        qi += 2
    final3 = np.zeros_like(final1)
    na3 = np.zeros_like(na1)
    for topo, rk in zip(p3.ranks, ranks):
        final3[p3.order[topo.lo:topo.hi]] = rk.y
        na3[p3.order[topo.lo:topo.hi]] = rk.na
    assert np.array_equal(final3, final1) and np.array_equal(dense3, dense1) and np.array_equal(na3, na1)


def test_model200_regression_fixture():
    """tests/golden/model200_routed.npz (written by tests/golden/make_model200_golden.py): the oracle with the
    restated libdevice pow reproduces it bit for bit — a self-pin over time, Model 200 has no reference counterpart."""
    import os
    from tests.golden.make_model200_golden import case
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "model200_routed.npz"))
    sp, col, pr, t2m, y0 = case()
    F = O.Forcing([pr, t2m], [1.0, 24.0], col=col)
    un = O.run_rk45(200, PRM, y0, 0.0, 720.0, 60.0 * np.arange(1, 13), sp=sp, forcing=F, device_pow=True, max_attempts=1_000_000)
    for k in ("final", "dense", "n_accept", "n_reject", "stiff"):
        assert np.array_equal(un[k], g["unrouted_" + k]), k
    p1 = routing.plan(sp["stream"], sp["next_stream"], 1, subbasin_links=12)
    fin, dense, tqr, na = routed_ref.run_single(sp, F, y0, PRM, p1, 0.0, 360.0, 20.0, threads=2)
    assert np.array_equal(fin, g["routed_final"]) and np.array_equal(dense, g["routed_dense"])
    assert np.array_equal(tqr, g["routed_tq"]) and np.array_equal(na, g["routed_n_accept"])
    assert (g["unrouted_dense"][:, :, 2] > 0).any()      # some links pond water: both pow() call sites are covered
