"""The C++ host path end to end: hlm_example (loadSpatialParams -> run_rk45<Model204> shim -> CSV writers)
against the oracle and the reference's golden, through the same files a reference user would read."""
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as O
from tiger_hlm_gpu_b200 import synthetic

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXAMPLE = os.path.join(ROOT, "tiger_hlm_gpu_b200", "host", "build", "hlm_example")
GOLDEN = os.path.join(ROOT, "tests", "golden")
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not os.path.exists(EXAMPLE), reason="hlm_example not built")]


def test_cpp_example_writes_reference_format_csvs(tmp_path, small_test_params, golden204):
    r = subprocess.run([EXAMPLE, os.path.join(GOLDEN, "small_test.csv"), str(tmp_path), "2", "60"], capture_output=True,
                       text=True)
    assert r.returncode == 0, r.stderr
    final = np.loadtxt(tmp_path / "final.csv", delimiter=",", skiprows=1)
    hdr = open(tmp_path / "dense.csv").readline().strip().split(",")
    assert hdr[0] == "time" and hdr[1] == "var0_sys0" and hdr[-1] == "var4_sys9" and len(hdr) == 51
    assert open(tmp_path / "final.csv").readline().strip() == "h_snow,var1,var2,var3,var4"
    dense = np.loadtxt(tmp_path / "dense.csv", delimiter=",", skiprows=1)
    assert dense.shape == (49, 51) and np.array_equal(dense[:, 0], np.arange(0, 2881, 60.0))
    sp = small_test_params
    ns = len(sp)
    y0 = np.tile(synthetic.Y0_204, (ns, 1))
    pr = np.full((48, ns), 0.001, np.float32)
    t2m = np.full((2, ns), 1.0, np.float32)
    o = O.run_rk45(204, O.Params.make(initialStep=1e-6), y0, 0.0, 2880.0, np.arange(0, 2881, 60.0), sp=sp,
                   forcing=O.Forcing([pr, t2m], [1.0, 24.0]), device_pow=True)
    np.testing.assert_allclose(final, o["final"], rtol=2e-6)            # final.csv holds 6 significant digits
    got = dense[:, 1:].reshape(49, ns, 5).transpose(1, 0, 2)
    np.testing.assert_allclose(got, o["dense"], rtol=2e-9, atol=5e-10)   # dense.csv: setprecision(9) fixed
    tol = 10 * (1e-9 + 1e-6 * np.abs(golden204["final"]))
    assert np.all(np.abs(final - golden204["final"]) <= tol + 2e-6 * np.abs(golden204["final"]))


def test_cpp_routed_example_matches_the_routed_oracle(tmp_path):
    """hlm_routed_example: parameter CSV with a river network -> hlm_b200::plan_routes -> RoutedRun (Model 200,
    coupling intervals, implicit fallback) -> final.csv/dense.csv, against the CPU routed run."""
    from tests import routed_ref
    from tiger_hlm_gpu_b200 import routing
    from tiger_hlm_gpu_b200.hostio import load_spatial_params, write_spatial_params_csv
    exe = os.path.join(ROOT, "tiger_hlm_gpu_b200", "host", "build", "hlm_routed_example")
    ns, hours, dt, sub = 600, 3.0, 15.0, 64
    sp = synthetic.apply_network(synthetic.make_spatial_params(ns), synthetic.make_network(ns, subbasin_links=sub, seed=3))
    csv = str(tmp_path / "params.csv")
    write_spatial_params_csv(csv, sp)
    sp = load_spatial_params(csv)                       # what the C++ loader sees (unit conversions round-tripped)
    r = subprocess.run([exe, csv, str(tmp_path), str(hours), str(dt), str(sub), "4", "2e-5", "8.0"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    p4 = routing.plan(sp["stream"], sp["next_stream"], 4, subbasin_links=sub)
    assert f"{p4.n_subbasins} sub-basins, {p4.n_cut_edges} cut edges, halo vector of {p4.halo_len} doubles" in r.stdout
    assert " 0 lost" in r.stdout
    final = np.loadtxt(tmp_path / "final.csv", delimiter=",", skiprows=1)
    dense = np.loadtxt(tmp_path / "dense.csv", delimiter=",", skiprows=1)
    n_int = int(hours * 60 / dt)
    assert dense.shape == (n_int, 1 + ns * 5)
    pr = np.full((int(hours + 1.5), ns), 2e-5, np.float32)
    t2m = np.full((1, ns), 8.0, np.float32)
    y0 = np.tile(synthetic.Y0_200, (ns, 1))
    p1 = routing.plan(sp["stream"], sp["next_stream"], 1, subbasin_links=sub)
    fin_o, dense_o, tq_o, _ = routed_ref.run_single(sp, O.Forcing([pr, t2m], [1.0, 24.0]), y0, O.Params.make(initialStep=1e-6), p1,
                                                    0.0, hours * 60.0, dt, threads=8)
    assert np.array_equal(dense[:, 0], tq_o)
    np.testing.assert_allclose(final, fin_o, rtol=6e-6)                                    # 6 significant digits in final.csv
    np.testing.assert_allclose(dense[:, 1:].reshape(n_int, ns, 5).transpose(1, 0, 2), dense_o, rtol=2e-9, atol=5e-10)
    assert fin_o[:, 0].max() > 5 * fin_o[:, 0].min()                                        # discharge accumulates downstream


def test_cpp_routed_nccl_two_processes(tmp_path):
    """hlm_routed_nccl: the C++ host of INTEGRATION.md section 6 — one process per GPU, hlm_b200::plan_routes,
    RoutedRun with ncclAllGather as the exchange — against the single-rank CPU routed run, bit for bit
    (final_rank_*.csv carry 17 significant digits)."""
    import torch
    from tests import routed_ref
    from tiger_hlm_gpu_b200 import routing
    from tiger_hlm_gpu_b200.hostio import load_spatial_params, write_spatial_params_csv
    exe = os.path.join(ROOT, "tiger_hlm_gpu_b200", "host", "build", "hlm_routed_nccl")
    if torch.cuda.device_count() < 2 or not os.path.exists(exe):
        pytest.skip("needs 2 GPUs and the NCCL example")
    ns, hours, dt, sub, world = 4000, 2.0, 15.0, 128, 2
    sp = synthetic.apply_network(synthetic.make_spatial_params(ns), synthetic.make_network(ns, subbasin_links=sub, seed=8))
    csv = str(tmp_path / "params.csv")
    write_spatial_params_csv(csv, sp)
    sp = load_spatial_params(csv)
    procs = [subprocess.Popen([exe, csv, str(tmp_path), str(hours), str(dt), str(sub), "2e-5", "8.0"],
                              env=dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), LOCAL_RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for r in range(world)]
    outs = [p.communicate(timeout=300) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert all(" 0 links lost" in o[0] for o in outs) and "all-gathers" in outs[0][0]
    got = {}
    for r in range(world):
        rows = np.loadtxt(tmp_path / f"final_rank_{r}.csv", delimiter=",", skiprows=1)
        for row in rows:
            got[int(row[0])] = row[1:]
    assert len(got) == ns
    pr = np.full((int(hours + 1.5), ns), 2e-5, np.float32)
    t2m = np.full((1, ns), 8.0, np.float32)
    y0 = np.tile(synthetic.Y0_200, (ns, 1))
    p1 = routing.plan(sp["stream"], sp["next_stream"], 1, subbasin_links=sub)
    fin_o, _, _, _ = routed_ref.run_single(sp, O.Forcing([pr, t2m], [1.0, 24.0]), y0, O.Params.make(initialStep=1e-6), p1, 0.0,
                                          hours * 60.0, dt, threads=8)
    fin_g = np.array([got[int(s)] for s in sp["stream"]])
    assert np.array_equal(fin_g, fin_o)
    assert routing.plan(sp["stream"], sp["next_stream"], world, subbasin_links=sub).n_cut_edges > 0


def test_reference_call_sequence_compiles_and_runs_against_the_shim(golden204):
    """host/refcall_selftest.cpp: setup_gpu_buffers -> launch_rk45_kernel -> retrieve_and_free with the reference's
    names, tuple shape and argument order (main.cpp:666-732, solver/rk45_api.hpp:63-270), then run_rk45 with the
    reference's signature on a cudaMalloc'ed d_sp (main.cpp:392-404); the program itself checks that both give the
    same bits, here the printed final state is checked against the reference's golden."""
    exe = os.path.join(ROOT, "tiger_hlm_gpu_b200", "host", "build", "hlm_refcall_selftest")
    r = subprocess.run([exe, os.path.join(GOLDEN, "small_test.csv")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert r.stdout.startswith("refcall ok: 10 systems, 49 queries")
    got = np.array([float(x) for x in r.stdout.split("=")[1].split()])
    want = golden204["final"][0]
    assert np.all(np.abs(got - want) <= 10 * (1e-9 + 1e-6 * np.abs(want)) + 1e-8 * np.abs(want))


def test_shim_bench_program_reports_throughput():
    """host/bench_shim.cpp (bench.py's e2e_shim record) on a small case: pooled page-locked result vectors."""
    import json
    exe = os.path.join(ROOT, "tiger_hlm_gpu_b200", "host", "build", "hlm_bench_shim")
    r = subprocess.run([exe, "20000", "2", "1"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["links"] == 20000 and d["steps"] == 2 and d["value"] > 1e6 and np.isfinite(d["checksum"])
