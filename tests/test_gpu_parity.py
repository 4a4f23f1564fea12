"""Parity of the CUDA path (through the C ABI) with the CPU oracle and the golden vectors.

Bar (north star): FP64 final and dense states within the solver's own tolerance, attempt counts
exactly equal.  Here the bar is stricter: BIT FOR BIT in every state, every dense record and every
counter, because add/mul/fma/div/sqrt are IEEE correctly rounded on both sides, the contraction
pattern is pinned (fp_exact.cuh) and the oracle restates libdevice's pow (oracle/devpow.h).
Against the reference's committed goldens (produced with a correctly rounded pow, see
oracle_rk45.c) the comparison is at the solver tolerance.
"""
import os

import numpy as np
import pytest

from oracle import oracle as O
import tiger_hlm_gpu_b200 as hlm
from tiger_hlm_gpu_b200 import Parameters, synthetic

pytestmark = pytest.mark.gpu

PRM = Parameters(initialStep=1e-6)  # main.cpp:633-640
OPRM = O.Params.make(initialStep=1e-6)


def orun(*a, **kw):
    """The oracle with CUDA libdevice's pow (oracle/devpow.h): bit-comparable with any CUDA build."""
    kw.setdefault("device_pow", True)
    return O.run_rk45(*a, **kw)


def assert_same_result(g, o, exact=True, rtol=1e-11):
    assert np.array_equal(g["stiff"], o["stiff"])
    assert np.array_equal(g["n_accept"], o["n_accept"])
    assert np.array_equal(g["n_reject"], o["n_reject"])
    assert np.array_equal(g["n_jump"], o["n_jump"])
    if exact:
        assert np.array_equal(g["final"], o["final"])
        if o.get("dense") is not None and g.get("dense") is not None:
            assert np.array_equal(g["dense"], o["dense"])
    else:
        np.testing.assert_allclose(g["final"], o["final"], rtol=rtol, atol=1e-300)
        if o.get("dense") is not None and g.get("dense") is not None:
            np.testing.assert_allclose(g["dense"], o["dense"], rtol=rtol, atol=1e-300)


def setup_synth(solver, ns, days, wet_fraction=0.0, grid=True):
    sp = synthetic.make_spatial_params(ns)
    col, ncells = synthetic.make_cells(ns, links_per_cell=97)
    pr, t2m = synthetic.make_forcing_grid(ncells, days)
    y0 = synthetic.make_y0(ns, wet_fraction)
    solver.set_model_parameters(204, PRM)
    solver.set_max_attempts(2_000_000)
    solver.upload_spatial_params(sp)
    solver.clear_forcings()
    if grid:
        solver.upload_forcing(0, 1.0, pr)
        solver.upload_forcing(1, 24.0, t2m)
        solver.set_forcing_columns(col)
    else:
        solver.upload_forcing(0, 1.0, synthetic.expand_forcing_per_link(pr, col))
        solver.upload_forcing(1, 24.0, synthetic.expand_forcing_per_link(t2m, col))
        solver.set_forcing_columns(None)
    forcing = O.Forcing([pr, t2m], [1.0, 24.0], col=col)
    return sp, y0, forcing


# ---- golden vectors --------------------------------------------------------------------------------

def test_model204_golden_example(solver, small_test_params, golden204):
    """C2: 10 links of data/small_test.csv, constant forcing, 2881 queries (src/*_example.nc)."""
    sp = small_test_params
    ns = len(sp)
    y0 = np.tile(synthetic.Y0_204, (ns, 1))
    tq = golden204["query_times"]
    solver.set_model_parameters(204, PRM)
    solver.set_max_attempts(100000)
    solver.upload_spatial_params(sp)
    solver.clear_forcings()
    # the golden used double literals 0.001 / 1.0; the operator takes float forcings (model_204.hpp:82-83)
    pr = np.full((48, ns), 0.001, np.float32)
    t2m = np.full((2, ns), 1.0, np.float32)
    solver.upload_forcing(0, 1.0, pr)
    solver.upload_forcing(1, 24.0, t2m)
    solver.set_forcing_columns(None)
    g = solver.run_rk45(204, y0, 0.0, 2880.0, tq)
    o = orun(204, OPRM, y0, 0.0, 2880.0, tq, sp=sp, forcing=O.Forcing([pr, t2m], [1.0, 24.0]))
    assert_same_result(g, o, exact=True)
    # against the reference's own output: float(0.001) vs the double literal shifts the snow store by
    # 4.7e-8 relative, well inside the solver tolerance
    tol = 10 * (1e-9 + 1e-6 * np.abs(golden204["final"]))
    assert np.all(np.abs(g["final"] - golden204["final"]) <= tol)
    gd = golden204["dense_sys0"]
    assert np.all(np.abs(g["dense"] - gd[None]) <= 10 * (1e-9 + 1e-6 * np.abs(gd[None])))
    assert np.all(g["dense"][:, 0, :] == 0.0)  # tq == t0 is never written


def test_dummy_golden(solver, golden_dummy):
    """C1: DummyModel, y0 = ones, t in [0,5], 10 000 queries (src/final.csv, src/dense.csv)."""
    tq = golden_dummy["query_times"]
    solver.set_model_parameters(0, Parameters())
    solver.set_max_attempts(100000)
    y0 = np.ones((4, 5))
    g = solver.run_rk45(0, y0, 0.0, 5.0, tq)
    o = orun(0, O.Params.make(), y0, 0.0, 5.0, tq)
    assert_same_result(g, o, exact=True)
    np.testing.assert_allclose(g["final"], golden_dummy["final_csv"], rtol=3e-6)
    np.testing.assert_allclose(g["dense"][0], golden_dummy["dense_csv_sys0"], rtol=2e-5, atol=1e-6)
    tol = 10 * (1e-9 + 1e-6 * np.abs(golden_dummy["scipy_final"]))
    assert np.all(np.abs(g["final"][0] - golden_dummy["scipy_final"]) < tol)


# ---- seeded synthetic inputs -----------------------------------------------------------------------

@pytest.mark.parametrize("ns,days", [(1, 2), (31, 2), (33, 2), (1000, 3)])
def test_synthetic_dry_bit_exact(solver, ns, days):
    sp, y0, forcing = setup_synth(solver, ns, days)
    tf = days * 1440.0
    tq = synthetic.hourly_queries(0.0, tf)
    g = solver.run_rk45(204, y0, 0.0, tf, tq)
    o = orun(204, OPRM, y0, 0.0, tf, tq, sp=sp, forcing=forcing, threads=8)
    assert g["n_accept"].min() > 100 * days
    assert_same_result(g, o, exact=True)


def test_synthetic_wet_bit_exact(solver):
    ns, days = 600, 2
    sp, y0, forcing = setup_synth(solver, ns, days, wet_fraction=0.5)
    tf = days * 1440.0
    tq = synthetic.hourly_queries(0.0, tf)
    g = solver.run_rk45(204, y0, 0.0, tf, tq)
    o = orun(204, OPRM, y0, 0.0, tf, tq, sp=sp, forcing=forcing, threads=8)
    assert (y0[:, 2] > 0).sum() > 100
    assert_same_result(g, o, exact=True)


def test_grid_forcing_equals_per_link_expanded_forcing(solver):
    """The grid + column map is the same operator as the reference's per-link expansion (main.cpp:543-548)."""
    ns, days = 500, 2
    tf = days * 1440.0
    tq = synthetic.hourly_queries(0.0, tf)
    _, y0, _ = setup_synth(solver, ns, days, grid=True)
    a = solver.run_rk45(204, y0, 0.0, tf, tq)
    setup_synth(solver, ns, days, grid=False)
    b = solver.run_rk45(204, y0, 0.0, tf, tq)
    assert_same_result(a, b, exact=True)


def test_windowed_run_is_bit_identical_to_single_window(solver):
    """Cutting the run into query windows must not change a single bit or count (no clipping of h)."""
    ns, days = 300, 3
    sp, y0, forcing = setup_synth(solver, ns, days, wet_fraction=0.3)
    tf = days * 1440.0
    tq = synthetic.hourly_queries(0.0, tf)
    solver.set_dense_window_bytes(8 << 30)
    a = solver.run_rk45(204, y0, 0.0, tf, tq)
    solver.set_dense_window_bytes(ns * 5 * 8 * 7)  # 7 queries per window -> 11 windows
    n0 = solver.launch_count()
    b = solver.run_rk45(204, y0, 0.0, tf, tq)
    assert solver.launch_count() - n0 >= 11
    solver.set_dense_window_bytes(ns * 5 * 8)      # 1 query per window
    c = solver.run_rk45(204, y0, 0.0, tf, tq)
    solver.set_dense_window_bytes(8 << 30)
    assert_same_result(a, b, exact=True)
    assert_same_result(a, c, exact=True)


def test_link_chunked_run_is_bit_identical_to_single_window(solver):
    """hlm_run_rk45 cuts a large output into chunks of links (each one contiguous block of the caller's
    array, copied while the next chunk integrates); ragged last chunk, ns not a multiple of 32."""
    ns, days = 3011, 1
    sp, y0, forcing = setup_synth(solver, ns, days, wet_fraction=0.3)
    tq = synthetic.hourly_queries(0.0, 1440.0)
    solver.set_dense_window_bytes(8 << 30)
    n0 = solver.launch_count()
    a = solver.run_rk45(204, y0, 0.0, 1440.0, tq)
    one = solver.launch_count() - n0
    solver.set_dense_window_bytes(len(tq) * 5 * 8 * 32 * 10)      # at most 10 tiles = 320 links per chunk
    n0 = solver.launch_count()
    b = solver.run_rk45(204, y0, 0.0, 1440.0, tq)
    assert solver.launch_count() - n0 >= one + 8                  # >= 10 window kernels (parameters are prepared once)
    solver.set_dense_window_bytes(8 << 30)
    assert_same_result(a, b, exact=True)
    o = orun(204, OPRM, y0, 0.0, 1440.0, tq, sp=sp, forcing=forcing, threads=8)
    assert_same_result(b, o, exact=True)


def test_link_chunked_run_with_optional_outputs_absent(solver):
    """hlm_run_rk45's outputs other than the dense records may be NULL; in the chunked path each absent one just
    drops its per-chunk copy."""
    import ctypes as C
    ns = 700
    sp, y0, forcing = setup_synth(solver, ns, 1)
    tq = synthetic.hourly_queries(0.0, 1440.0)
    a = solver.run_rk45(204, y0, 0.0, 1440.0, tq)
    dense = np.zeros((ns, len(tq), 5))
    acc = np.zeros(ns, np.int64)
    lib = hlm.load_library()
    try:
        solver.set_dense_window_bytes(len(tq) * 5 * 8 * 32 * 4)
        p = lambda x: x.ctypes.data_as(C.c_void_p)
        rc = lib.hlm_run_rk45(solver._h, 204, p(y0), ns, 0.0, 1440.0, p(tq), len(tq), None, p(dense), None, p(acc), None, None)
        assert rc == 0, lib.hlm_last_error().decode()
    finally:
        solver.set_dense_window_bytes(8 << 30)
    assert np.array_equal(dense, a["dense"]) and np.array_equal(acc, a["n_accept"])


def test_lane_refill_schedule_is_bit_identical_to_tiles(solver, golden_dummy):
    """hlm_set_schedule: a lane takes the next unclaimed link when it is done with its own instead of waiting for
    its tile.  Same per-link arithmetic, so the same bits — whole run, windowed, link-chunked, DummyModel."""
    ns, days = 1237, 2
    sp, y0, forcing = setup_synth(solver, ns, days, wet_fraction=0.4)
    tf = days * 1440.0
    tq = synthetic.hourly_queries(0.0, tf)
    a = solver.run_rk45(204, y0, 0.0, tf, tq)
    try:
        solver.set_schedule("lanes")
        b = solver.run_rk45(204, y0, 0.0, tf, tq)
        solver.set_dense_window_bytes(ns * 5 * 8 * 5)                   # windows of 5 queries
        c = solver.run_rk45(204, y0, 0.0, tf, tq)
        solver.set_dense_window_bytes(len(tq) * 5 * 8 * 32 * 7)          # chunks of 7 tiles
        d = solver.run_rk45(204, y0, 0.0, tf, tq)
        solver.set_dense_window_bytes(8 << 30)
        solver.set_model_parameters(0, Parameters())
        yd = np.ones((70, 5)) * np.linspace(0.5, 1.5, 70)[:, None]
        e = solver.run_rk45(0, yd, 0.0, 5.0, golden_dummy["query_times"])
    finally:
        solver.set_schedule("auto")
        solver.set_dense_window_bytes(8 << 30)
    for other in (b, c, d):
        assert_same_result(a, other, exact=True)
    assert_same_result(e, orun(0, O.Params.make(), yd, 0.0, 5.0, golden_dummy["query_times"]), exact=True)
    solver.set_model_parameters(204, PRM)


def test_sorted_tiles_schedule_is_bit_identical_to_tiles(solver):
    """HLM_SCHEDULE_SORTED_TILES for a model without an inflow term (its second tile-kernel instance): three days chained
    in a resident session — the first launch has no attempt counts to sort by and takes lane refill, the next two run as
    tiles of links with equal counts — give the bits of the plain tile schedule, dense records and counters included."""
    ns, days = 2500, 3
    sp, y0, forcing = setup_synth(solver, ns, days, wet_fraction=0.4)
    runs = {}
    try:
        for schedule in ("tiles", "sorted"):
            solver.set_schedule(schedule)
            dense = []
            for d in range(days):
                tq = synthetic.hourly_queries(d * 1440.0, (d + 1) * 1440.0)
                if d == 0:
                    solver.solve_begin(204, y0, 0.0, 1440.0, tq)
                else:
                    solver.solve_restart(d * 1440.0, (d + 1) * 1440.0, tq)
                solver.solve_window(len(tq), True)
                win = np.zeros((ns, len(tq), 5))
                solver.solve_wait_copy(solver.solve_fetch_window_packed(win))
                dense.append(win)
            r = solver.solve_end()
            r["dense"] = np.concatenate(dense, axis=1)
            runs[schedule] = r
    finally:
        solver.set_schedule("auto")
    a, b = runs["tiles"], runs["sorted"]
    assert a["n_accept"].sum() > 100 * ns
    for k in ("final", "dense", "stiff", "n_accept", "n_reject", "n_jump"):
        assert np.array_equal(a[k], b[k]), k


def test_windows_with_steps_longer_than_the_query_spacing(solver, golden_dummy):
    """DummyModel takes 12 steps for 10 000 queries: every window boundary falls inside a step, which
    exercises the leave-uncommitted-and-redo path."""
    tq = golden_dummy["query_times"]
    solver.set_model_parameters(0, Parameters())
    y0 = np.ones((40, 5)) * np.linspace(0.5, 1.5, 40)[:, None]
    a = solver.run_rk45(0, y0, 0.0, 5.0, tq)
    solver.set_dense_window_bytes(40 * 5 * 8 * 333)
    b = solver.run_rk45(0, y0, 0.0, 5.0, tq)
    solver.set_dense_window_bytes(8 << 30)
    assert_same_result(a, b, exact=True)
    o = orun(0, O.Params.make(), y0, 0.0, 5.0, tq)
    assert_same_result(a, o, exact=True)


# ---- edge cases ------------------------------------------------------------------------------------

def test_no_queries_and_final_only(solver):
    ns, days = 200, 1
    sp, y0, forcing = setup_synth(solver, ns, days)
    tf = 1440.0
    g = solver.run_rk45(204, y0, 0.0, tf, None)
    o = orun(204, OPRM, y0, 0.0, tf, np.zeros(0), sp=sp, forcing=forcing)
    assert g["dense"] is None
    assert_same_result(g, o, exact=True)
    tq = synthetic.hourly_queries(0.0, tf)
    g2 = solver.run_rk45(204, y0, 0.0, tf, tq, want_dense=False)
    assert np.array_equal(g2["final"], g["final"]) and np.array_equal(g2["n_accept"], g["n_accept"])


def test_queries_outside_the_interval(solver):
    """tq <= t0 slots stay zero (F10); tq > tf slots are never reached."""
    ns = 64
    sp, y0, forcing = setup_synth(solver, ns, 1)
    tq = np.array([-5.0, 0.0, 1e-3, 30.0, 600.0, 1440.0, 1440.0 + 1e-9, 2000.0])
    g = solver.run_rk45(204, y0, 0.0, 1440.0, tq)
    o = orun(204, OPRM, y0, 0.0, 1440.0, tq, sp=sp, forcing=forcing)
    assert_same_result(g, o, exact=True)
    assert np.all(g["dense"][:, :2] == 0) and np.all(g["dense"][:, 6:] == 0) and np.all(g["dense"][:, 2:6, 1] != 0)


def test_empty_interval(solver):
    ns = 40
    sp, y0, forcing = setup_synth(solver, ns, 1)
    g = solver.run_rk45(204, y0, 10.0, 10.0, np.array([10.0, 11.0]))
    assert np.array_equal(g["final"], y0) and not g["n_accept"].any() and not g["stiff"].any()
    assert np.all(g["dense"] == 0)


def test_stiff_flag_matches_reference_semantics(solver):
    """Raw-unit rain (the reference feeds pr unscaled) drives links stiff; flagged links
    report no final state and the same flags/counters as the oracle."""
    ns, days = 128, 2
    sp = synthetic.make_spatial_params(ns)
    col, ncells = synthetic.make_cells(ns, links_per_cell=16)
    pr, t2m = synthetic.make_forcing_grid(ncells, days)
    pr = (pr / synthetic.C1).astype(np.float32)  # back to mm/h magnitudes, i.e. 6e4 x too large
    y0 = synthetic.make_y0(ns)
    solver.set_model_parameters(204, PRM)
    solver.upload_spatial_params(sp)
    solver.clear_forcings()
    solver.upload_forcing(0, 1.0, pr)
    solver.upload_forcing(1, 24.0, t2m)
    solver.set_forcing_columns(col)
    tf = days * 1440.0 * 15  # long horizon raises the min-step threshold (tf - t0)*1e-6
    tq = synthetic.hourly_queries(0.0, 2880.0)
    g = solver.run_rk45(204, y0, 0.0, tf, tq)
    o = orun(204, OPRM, y0, 0.0, tf, tq, sp=sp, forcing=O.Forcing([pr, t2m], [1.0, 24.0], col=col), threads=8)
    assert o["stiff"].sum() > 0, "test input should drive some links stiff"
    assert_same_result(g, o, exact=True)
    assert np.all(g["final"][g["stiff"] == 1] == 0.0)


def test_attempt_budget_reports_stalled(solver):
    ns = 32
    sp, y0, forcing = setup_synth(solver, ns, 1)
    solver.set_max_attempts(50)
    g = solver.run_rk45(204, y0, 0.0, 1440.0, None)
    solver.set_max_attempts(2_000_000)
    assert np.all(g["stiff"] == 2) and np.all(g["n_accept"] + g["n_reject"] + g["n_jump"] == 50)
    assert np.all(g["final"] == 0)


def test_errors_are_loud(solver):
    from tiger_hlm_gpu_b200 import HlmError
    with pytest.raises(HlmError, match="unknown model uid"):
        solver.set_model_parameters(190, Parameters())
    with pytest.raises(HlmError, match="unknown model uid"):
        solver.run_rk45(190, np.ones((4, 5)), 0.0, 1.0, None)
    solver.upload_spatial_params(synthetic.make_spatial_params(10))
    solver.clear_forcings()
    with pytest.raises(HlmError, match="exactly ns SpatialParams"):
        solver.run_rk45(204, np.ones((11, 5)), 0.0, 1.0, None)
    # a forcing column outside the arrays would be an out-of-bounds device read: refused on the host
    solver.upload_forcing(0, 1.0, np.zeros((3, 4), np.float32))
    with pytest.raises(HlmError, match="column 4"):
        solver.set_forcing_columns(np.array([0, 1, 4, 2, 0, 0, 0, 0, 0, 0], np.int32))
    with pytest.raises(HlmError, match="negative column"):
        solver.set_forcing_columns(np.array([0, -1], np.int32))
    solver.set_forcing_columns(np.array([0, 1, 3, 2, 0, 0, 0, 0, 0, 0], np.int32))
    with pytest.raises(HlmError, match="refers to column 3"):
        solver.upload_forcing(0, 1.0, np.zeros((3, 3), np.float32))
    solver.clear_forcings()


# ---- resident session, partition invariance, FP32 ---------------------------------------------------

def test_session_windows_and_totals(solver):
    ns, days = 257, 2
    sp, y0, forcing = setup_synth(solver, ns, days)
    tf = days * 1440.0
    tq = synthetic.hourly_queries(0.0, tf)
    full = solver.run_rk45(204, y0, 0.0, tf, tq)
    solver.solve_begin(204, y0, 0.0, tf, tq)
    dense = np.zeros((ns, len(tq), 5))
    for q in (10, 25, 48, len(tq)):
        solver.solve_window(q)
        solver.solve_fetch_window(dense)
    tot = solver.solve_totals()
    r = solver.solve_end()
    assert np.array_equal(dense, full["dense"]) and np.array_equal(r["final"], full["final"])
    assert tot["n_accept"] == full["n_accept"].sum() and tot["n_reject"] == full["n_reject"].sum()
    assert tot["done"] == ns and tot["active"] == 0 and tot["stiff"] == 0


def test_partition_invariance(solver):
    """Sharding links over devices = running disjoint slices; per-link results must not depend on it."""
    ns, days = 400, 2
    sp = synthetic.make_spatial_params(ns)
    col, ncells = synthetic.make_cells(ns, links_per_cell=97)
    pr, t2m = synthetic.make_forcing_grid(ncells, days)
    y0 = synthetic.make_y0(ns, 0.25)
    tf = days * 1440.0
    tq = synthetic.hourly_queries(0.0, tf)
    solver.set_model_parameters(204, PRM)
    solver.clear_forcings()
    solver.upload_forcing(0, 1.0, pr)
    solver.upload_forcing(1, 24.0, t2m)

    def run(lo, hi):
        solver.upload_spatial_params(sp[lo:hi])
        solver.set_forcing_columns(col[lo:hi])
        return solver.run_rk45(204, y0[lo:hi], 0.0, tf, tq)

    whole = run(0, ns)
    for cuts in ([0, 200, 400], [0, 100, 200, 300, 400], [0, 57, 400]):
        parts = [run(a, b) for a, b in zip(cuts[:-1], cuts[1:])]
        for key in ("final", "dense", "n_accept", "n_reject", "n_jump", "stiff"):
            assert np.array_equal(np.concatenate([p[key] for p in parts]), whole[key])


def test_fp32_mode_tracks_fp64(solver):
    """FP32 has no reference counterpart; it must stay within a stated looser tolerance of FP64."""
    ns, days = 256, 1
    sp, y0, forcing = setup_synth(solver, ns, days)
    tf = days * 1440.0
    tq = synthetic.hourly_queries(0.0, tf)
    solver.set_model_parameters(204, Parameters(initialStep=1e-4, rtol=1e-4, atol=1e-7))
    d = solver.run_rk45(204, y0, 0.0, tf, tq)
    solver.set_precision(32)
    try:
        s = solver.run_rk45(204, y0, 0.0, tf, tq)
    finally:
        solver.set_precision(64)
        solver.set_model_parameters(204, PRM)
    assert not s["stiff"].any()
    np.testing.assert_allclose(s["final"], d["final"], rtol=2e-3, atol=2e-6)
    np.testing.assert_allclose(s["dense"], d["dense"], rtol=2e-3, atol=2e-6)


def test_restart_equals_chained_run_rk45(solver):
    """hlm_solve_restart = a second run_rk45 from the previous final state, without the host round trip."""
    ns = 300
    sp, y0, forcing = setup_synth(solver, ns, 3, wet_fraction=0.3)
    days = 3
    chained, y = [], y0
    for k in range(days):
        tq = k * 1440.0 + 60.0 * np.arange(1, 25)
        r = solver.run_rk45(204, y, k * 1440.0, (k + 1) * 1440.0, tq)
        o = orun(204, OPRM, y, k * 1440.0, (k + 1) * 1440.0, tq, sp=sp, forcing=forcing, threads=8)
        assert_same_result(r, o, exact=True)
        chained.append(r)
        y = r["final"]
    solver.solve_begin(204, y0, 0.0, 1440.0, 60.0 * np.arange(1, 25))
    for k in range(days):
        if k:
            solver.solve_restart(k * 1440.0, (k + 1) * 1440.0, k * 1440.0 + 60.0 * np.arange(1, 25))
        solver.solve_window(24)
        dense = np.zeros((ns, 24, 5))
        solver.solve_fetch_window(dense)
        solver.synchronize()
        assert np.array_equal(dense, chained[k]["dense"])
    r = solver.solve_end()
    assert np.array_equal(r["final"], chained[-1]["final"])
    assert np.array_equal(r["n_accept"], sum(c["n_accept"] for c in chained))


def test_unusual_parameters_take_the_exact_fallback(solver):
    """Links whose divisors or states leave the fast paths' ranges (alpha < 1, alpha huge, zero or
    negative storages, Hu tiny) must still match the oracle bit for bit via the exact redo."""
    ns, days = 96, 1
    sp = synthetic.make_spatial_params(ns)
    sp["alpha3"][0:8] = 0.5          # < 1: term is 0 (model_204.hpp:109)
    sp["alpha4"][8:16] = 0.0
    sp["alpha3"][16:24] = 1e300      # reciprocal out of range -> exact division
    sp["Hu"][24:32] = 1e-30
    sp["alpha4"][32:40] = 1.0
    sp["Hu"][40:48] = 3.5            # saturation excess: h_surf becomes positive on its own
    col, ncells = synthetic.make_cells(ns, links_per_cell=16)
    pr, t2m = synthetic.make_forcing_grid(ncells, days)
    y0 = synthetic.make_y0(ns, 0.5)
    y0[48:56, 3] = 0.0               # zero numerator of h_grav / alpha3
    y0[56:64, 4] = 0.0
    y0[64:72, 1] = 1e-320            # subnormal numerator of h_stat / Hu
    y0[72:80, 3] = -1e-3             # negative storage
    solver.set_model_parameters(204, PRM)
    solver.set_max_attempts(2_000_000)
    solver.upload_spatial_params(sp)
    solver.clear_forcings()
    solver.upload_forcing(0, 1.0, pr)
    solver.upload_forcing(1, 24.0, t2m)
    solver.set_forcing_columns(col)
    tf = days * 1440.0
    tq = synthetic.hourly_queries(0.0, tf)
    g = solver.run_rk45(204, y0, 0.0, tf, tq)
    o = orun(204, OPRM, y0, 0.0, tf, tq, sp=sp, forcing=O.Forcing([pr, t2m], [1.0, 24.0], col=col), threads=8)
    assert_same_result(g, o, exact=True)


def test_time_chunked_forcing_equals_resident_record(solver):
    """hlm_upload_forcing_chunk: only the samples an interval needs are resident; indexing and the
    clamp at the end of the record stay those of the whole record (solver/rk45_kernel.cu:90-98)."""
    ns, days = 200, 4
    sp, y0, forcing = setup_synth(solver, ns, days, wet_fraction=0.2)
    pr, t2m = forcing.blocks
    col = forcing.col
    # reference result: whole record resident, one interval per day, the last one running past the record
    def drive(upload):
        out = []
        for k in range(days + 1):
            tq = k * 1440.0 + 120.0 * np.arange(1, 13)
            upload(k)
            if k == 0:
                solver.solve_begin(204, y0, 0.0, 1440.0, tq)
            else:
                solver.solve_restart(k * 1440.0, (k + 1) * 1440.0, tq)
            solver.solve_window(12)
            win = np.zeros((ns, 12, 5))
            t = solver.solve_fetch_window_packed(win)
            solver.solve_wait_copy(t)
            out.append(win)
        r = solver.solve_end()
        return np.concatenate(out, axis=1), r
    d_full, r_full = drive(lambda k: None)

    def upload_chunk(k):  # samples of day k (+ the one the interval's end time indexes), clamped to the record
        lo = min(k * 24, pr.shape[0] - 1)
        hi = min((k + 1) * 24, pr.shape[0] - 1)
        solver.upload_forcing_chunk(0, 1.0, pr.shape[0], lo, pr[lo:hi + 1])
        lo2 = min(k, t2m.shape[0] - 1)
        hi2 = min(k + 1, t2m.shape[0] - 1)
        solver.upload_forcing_chunk(1, 24.0, t2m.shape[0], lo2, t2m[lo2:hi2 + 1])
    solver.set_forcing_columns(col)
    d_chunk, r_chunk = drive(upload_chunk)
    assert np.array_equal(d_chunk, d_full)
    for key in ("final", "stiff", "n_accept", "n_reject", "n_jump"):
        assert np.array_equal(r_chunk[key], r_full[key]), key
    # a chunk that does not cover the interval is refused, not silently clamped
    solver.upload_forcing_chunk(0, 1.0, pr.shape[0], 30, pr[30:40])
    solver.solve_begin(204, y0, 0.0, 1440.0, 120.0 * np.arange(1, 13))
    with pytest.raises(hlm.HlmError, match="does not cover"):
        solver.solve_window(12)
    solver.solve_end()
    solver.clear_forcings()


# ---- output selection on the device (hlm_set_output_states / hlm_set_output_precision) --------------

def test_output_states_and_precision_on_device(solver):
    """config.yaml's output.states applied where the records are produced: the selected columns of the full
    run, bit for bit, through all three ways a dense record leaves (one window, chunks of links, session
    windows); as float32 they are the full run's values rounded once."""
    ns, days = 300, 2
    sp, y0, forcing = setup_synth(solver, ns, days, wet_fraction=0.3)
    tf = days * 1440.0
    tq = synthetic.hourly_queries(0.0, tf)
    full = solver.run_rk45(204, y0, 0.0, tf, tq)
    sel = [0, 3]
    try:
        solver.set_output_states(sel)
        assert solver.output_layout(204) == (2, np.float64)
        g = solver.run_rk45(204, y0, 0.0, tf, tq)
        assert g["dense"].shape == (ns, tq.size, 2)
        assert np.array_equal(g["dense"], full["dense"][:, :, sel])
        for key in ("final", "n_accept", "n_reject", "n_jump", "stiff"):
            assert np.array_equal(g[key], full[key])
        solver.set_output_precision(32)
        assert solver.output_layout(204) == (2, np.float32)
        want32 = full["dense"][:, :, sel].astype(np.float32)
        g32 = solver.run_rk45(204, y0, 0.0, tf, tq)
        assert g32["dense"].dtype == np.float32 and np.array_equal(g32["dense"], want32)
        # chunks of links (run_link_chunks): 64 KiB window buffers
        solver.set_dense_window_bytes(64 << 10)
        c32 = solver.run_rk45(204, y0, 0.0, tf, tq)
        assert np.array_equal(c32["dense"], want32) and np.array_equal(c32["final"], full["final"])
        solver.set_dense_window_bytes(8 << 30)
        # session windows of 7 queries into the full host array
        host = np.full((ns, tq.size, 2), np.nan, np.float32)
        solver.solve_begin(204, y0, 0.0, tf, tq)
        for q in list(range(7, tq.size, 7)) + [tq.size]:
            solver.solve_window(q)
            solver.solve_fetch_window(host)
        solver.synchronize()
        r = solver.solve_end()
        assert np.array_equal(host, want32) and np.array_equal(r["final"], full["final"])
        # one state, double
        solver.set_output_precision(64)
        solver.set_output_states([4])
        g1 = solver.run_rk45(204, y0, 0.0, tf, tq)
        assert np.array_equal(g1["dense"][:, :, 0], full["dense"][:, :, 4])
    finally:
        solver.set_output_states(None)
        solver.set_output_precision(64)
        solver.set_dense_window_bytes(8 << 30)
    again = solver.run_rk45(204, y0, 0.0, tf, tq)
    assert np.array_equal(again["dense"], full["dense"])


def test_every_dense_slot_is_written_by_the_kernel(solver):
    """No memset precedes a window launch: the kernel itself writes zeros for queries a link never reaches
    (tq <= t0, after a stall, after tf).  The buffers are first filled with other values by a run of the same
    shape; what the second run leaves must equal the oracle's array, zeros included."""
    ns = 200
    sp, y0, forcing = setup_synth(solver, ns, 1, wet_fraction=0.5)
    tq = np.concatenate([[-30.0, 0.0], 60.0 * np.arange(1, 25), [1500.0, 1600.0]])  # before t0, at t0, inside, after tf
    for _ in range(2):  # both window buffers hold non-zero records of this shape
        first = solver.run_rk45(204, y0 + 1.0, 0.0, 1440.0, tq)
    assert np.count_nonzero(first["dense"][:, 2:26]) > 0
    solver.set_max_attempts(150)  # most links run out of attempts before the last queries
    try:
        g = solver.run_rk45(204, y0, 0.0, 1440.0, tq)
        o = orun(204, OPRM, y0, 0.0, 1440.0, tq, sp=sp, forcing=forcing, max_attempts=150, threads=8)
    finally:
        solver.set_max_attempts(2_000_000)
    stalled = g["stiff"] == 2  # HLM_LINK_STALLED (the oracle has no such code: it stops at the same attempt and reports the state)
    assert stalled.sum() > ns // 2
    for key in ("n_accept", "n_reject", "n_jump", "dense"):
        assert np.array_equal(g[key], o[key]), key
    assert np.all(g["final"][stalled] == 0.0) and np.array_equal(g["final"][~stalled], o["final"][~stalled])
    assert np.all(g["dense"][:, :2] == 0.0) and np.all(g["dense"][:, 26:] == 0.0)
    assert np.all(g["dense"][stalled, 25] == 0.0)  # the last hour lies beyond what 150 attempts reach
    # the same through session windows, where a stalled link is met again by later launches
    solver.set_max_attempts(150)
    try:
        host = np.full((ns, tq.size, 5), np.nan)
        solver.solve_begin(204, y0, 0.0, 1440.0, tq)
        solver.solve_window(2)
        solver.solve_fetch_window(host)
        for q in (9, 20, tq.size):
            solver.solve_window(q)
            solver.solve_fetch_window(host)
        solver.synchronize()
        solver.solve_end()
    finally:
        solver.set_max_attempts(2_000_000)
    assert not np.isnan(host).any()
    assert np.all(host[:, :2] == 0.0) and np.all(host[:, 26:] == 0.0)


def test_fp32_mode_at_the_bench_tolerances(solver):
    """BASELINE configs[3] "FP32 vs FP64": the bench runs the FP32 mode at the FP64 settings (rtol 1e-6, atol 1e-9,
    initialStep 1e-6) on the bench workload.  Two stated bounds, both at exactly those settings:

    (a) forcing constant in time: every link finishes and final + hourly dense states are within
        25 * (atol + rtol*|y|) of the FP64 run (measured worst case 12.3, h_stat; the other states <= 4.5) — the
        solver's own tolerance with the factor a global error over ~470 steps takes;
    (b) the bench workload (hourly changing rain): the reference samples forcings at the step START and holds
        them over the step, so a state depends on where steps fall relative to the hour marks — at
        the 1e-3 level for h_stat, for ANY change of the step sequence.  FP64 itself moves by that much between
        rtol 1e-6 and 1e-5.  The bound is therefore FP64's own sensitivity S = |FP64(10x tolerances) - FP64|:
        |FP32 - FP64| <= 8 * (atol + rtol*|y|) + 2 * S per state (measured: 0.9e-3 vs S = 0.87e-3 for h_stat;
        the states without that sensitivity stay within 5 * (atol + rtol*|y|))."""
    ns, days = 4096, 1
    sp = synthetic.make_spatial_params(ns)
    col, ncells = synthetic.make_cells(ns)
    pr, t2m = synthetic.make_forcing_grid(ncells, days)
    y0 = synthetic.make_y0(ns, wet_fraction=0.5)
    tq = 60.0 * np.arange(1, 25)
    solver.set_max_attempts(2_000_000)
    solver.upload_spatial_params(sp)

    def run(forc_pr, forc_t, bits, prm):
        solver.set_model_parameters(204, prm)
        solver.clear_forcings()
        solver.upload_forcing(0, 1.0, forc_pr)
        solver.upload_forcing(1, 24.0, forc_t)
        solver.set_forcing_columns(col)
        solver.set_precision(bits)
        try:
            return solver.run_rk45(204, y0, 0.0, 1440.0, tq)
        finally:
            solver.set_precision(64)
            solver.set_model_parameters(204, PRM)

    # (a) constant in time, varying over cells
    pr_c = np.repeat(pr[5:6], pr.shape[0], axis=0)
    pr_c[:, ::2] = pr.max() * 0.5
    t_c = np.repeat(t2m[0:1], t2m.shape[0], axis=0)
    d, s = run(pr_c, t_c, 64, PRM), run(pr_c, t_c, 32, PRM)
    assert not d["stiff"].any()
    assert not s["stiff"].any(), f"{(s['stiff'] != 0).sum()} of {ns} links did not finish in FP32"
    for key in ("final", "dense"):
        excess = np.abs(s[key] - d[key]) / (1e-9 + 1e-6 * np.abs(d[key]))
        assert excess.max() <= 25.0, f"FP32 {key} (constant forcing): {excess.max():.1f} x (atol + rtol|y|)"
    # (b) the bench workload
    d, s = run(pr, t2m, 64, PRM), run(pr, t2m, 32, PRM)
    loose = run(pr, t2m, 64, Parameters(initialStep=1e-6, rtol=1e-5, atol=1e-8))
    assert not d["stiff"].any()
    assert not s["stiff"].any(), f"{(s['stiff'] != 0).sum()} of {ns} links did not finish in FP32"
    for key in ("final", "dense"):
        for c in range(5):
            err = np.abs(s[key][..., c] - d[key][..., c])
            sens = np.abs(loose[key][..., c] - d[key][..., c]).max()
            bound = 8.0 * (1e-9 + 1e-6 * np.abs(d[key][..., c])) + 2.0 * sens
            assert np.all(err <= bound), (f"FP32 {key} state {c}: worst {np.max(err / bound):.2f} x the bound "
                                          f"(FP64 sensitivity to a 10x tolerance change: {sens:.3e})")
    ratio = s["n_accept"].sum() / d["n_accept"].sum()
    assert 0.9 < ratio < 1.1, ratio  # measured 1.00003: the controller is not yet at its noise floor


# ---- BASELINE configs[0] / configs[1] at the size of the reference's (absent) small_example_params.csv ----------

N_SMALL_EXAMPLE = 41_274  # rows of data/small_example_pr_lookup.csv: the presumed size of small_example_params.csv


def test_dummy_model_at_the_small_example_size(solver, golden_dummy):
    """C1 at 41 274 systems (the reference's DummyModel run on data/small_example_params.csv): every system is the same
    ODE from the same y0, so every row must carry the bits of the oracle's single system — and the golden's values."""
    tq = golden_dummy["query_times"][::50]  # 200 of the 10 000 queries: the dense array stays at 330 MB
    solver.set_model_parameters(0, Parameters())
    solver.set_max_attempts(100000)
    y0 = np.ones((N_SMALL_EXAMPLE, 5))
    g = solver.run_rk45(0, y0, 0.0, 5.0, tq)
    o = orun(0, O.Params.make(), y0[:1], 0.0, 5.0, tq)
    assert not g["stiff"].any()
    assert np.all(g["n_accept"] == o["n_accept"][0]) and np.all(g["n_reject"] == o["n_reject"][0])
    assert np.array_equal(g["final"], np.repeat(o["final"], N_SMALL_EXAMPLE, axis=0))
    assert np.array_equal(g["dense"], np.repeat(o["dense"], N_SMALL_EXAMPLE, axis=0))
    np.testing.assert_allclose(g["final"][-1], golden_dummy["final_csv"][0], rtol=3e-6)


def test_model204_at_the_small_example_size(solver):
    """C2 at 41 274 links: the reference's run shape (t in [0, 2880] min, 49 hourly queries, main.cpp:610-657) on synthetic
    parameters and a forcing grid of 88 cells (the cell count of the reference's lookup files), bit for bit against
    the oracle in every state, dense record and counter."""
    ns = N_SMALL_EXAMPLE
    sp = synthetic.make_spatial_params(ns, seed=41)
    col, ncells = synthetic.make_cells(ns, links_per_cell=470)
    assert ncells == 88
    pr, t2m = synthetic.make_forcing_grid(ncells, 2, seed=274)
    y0 = synthetic.make_y0(ns, wet_fraction=0.2, seed=3)
    solver.set_model_parameters(204, PRM)
    solver.set_max_attempts(2_000_000)
    solver.upload_spatial_params(sp)
    solver.clear_forcings()
    solver.upload_forcing(0, 1.0, pr)
    solver.upload_forcing(1, 24.0, t2m)
    solver.set_forcing_columns(col)
    tq = synthetic.hourly_queries(0.0, 2880.0)
    assert tq.size == 49
    g = solver.run_rk45(204, y0, 0.0, 2880.0, tq)
    o = orun(204, OPRM, y0, 0.0, 2880.0, tq, sp=sp, forcing=O.Forcing([pr, t2m], [1.0, 24.0], col=col), threads=os.cpu_count() or 8)
    assert_same_result(g, o, exact=True)
    assert (g["stiff"] == 0).sum() > 0.9 * ns
