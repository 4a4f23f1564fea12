"""End-to-end run driver (tiger_hlm_gpu_b200/host/hlm_run.cpp) on a B200: config.yaml + parameters CSV +
lookup CSV + forcing NetCDF in, final/dense NetCDF out — the reference's main() flow (main.cpp:314-823)
for this path.  Outputs must equal, bit for bit, the same run driven through the Python mirror of the
C ABI and the CPU oracle chained over the same day-sized intervals."""
import os
import subprocess

import numpy as np
import pytest
from scipy.io import netcdf_file

import tiger_hlm_gpu_b200 as hlm
from tiger_hlm_gpu_b200 import hostio
from oracle import oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "tiger_hlm_gpu_b200", "host")
RUN = os.path.join(HOST, "build", "hlm_run")
COLS = ["stream", "next_stream", "i2", "i3", "hu", "centroid_lat", "sw", "ss", "n", "slope", "length_km",
        "drainage_area_km2", "melt", "t_thres", "res_ss", "res_gw"]
NLAT, NLON, DAYS, NS = 3, 4, 3, 203


def write_case(d, world_dirs=True):
    rng = np.random.default_rng(11)
    os.makedirs(d / "params"), os.makedirs(d / "forc"), os.makedirs(d / "out")
    stream = 420000000 + np.arange(NS)
    rows = np.column_stack([
        stream, stream + 1, np.full(NS, 4.0), np.full(NS, 1.6), rng.uniform(150, 200, NS), rng.uniform(38, 42, NS),
        np.full(NS, 0.11), np.full(NS, 0.33), np.full(NS, 0.1), np.full(NS, 0.02), rng.uniform(0.09, 2.1, NS),
        np.exp(rng.uniform(np.log(0.13), np.log(1.6), NS)), np.full(NS, 3.7), np.zeros(NS), np.full(NS, 2.0), np.full(NS, 55.0)])
    with open(d / "params" / "links.csv", "w") as f:
        f.write(",".join(COLS) + "\n")
        for r in rows:
            f.write(",".join([str(int(r[0])), str(int(r[1]))] + [repr(float(x)) for x in r[2:]]) + "\n")
    lat_i, lon_i = rng.integers(0, NLAT, NS), rng.integers(0, NLON, NS)
    with open(d / "forc" / "lookup.csv", "w") as f:
        f.write("stream,lat_index,lon_index\n")
        for s, a, b in zip(stream, lat_i, lon_i):
            f.write(f"{s},{a},{b}\n")
    # heavy enough rain that some links pond water (the pow() branch) without going stiff
    wet = rng.random((DAYS * 24, NLAT, NLON)) < 0.3
    pr = (np.where(wet, rng.exponential(2.0, wet.shape), 0.0) * (0.001 / 60.0) * 8.0).astype(np.float32)
    t2m = (5.0 + 6.0 * rng.standard_normal((DAYS, NLAT, NLON))).astype(np.float32)
    with netcdf_file(str(d / "forc" / "a_pr_hourly.nc"), "w", version=2) as f:
        f.createDimension("time", None), f.createDimension("latitude", NLAT), f.createDimension("longitude", NLON)
        tv = f.createVariable("time", "i4", ("time",))
        tv.units = "hours since 2021-01-01 00:00:00"
        tv[:] = np.arange(DAYS * 24)
        f.createVariable("PRCP", "f4", ("time", "latitude", "longitude"))[:] = pr
    with netcdf_file(str(d / "forc" / "b_t2m_daily.nc"), "w", version=1) as f:  # no time coordinate: dt falls back to 24 h
        f.createDimension("time", DAYS), f.createDimension("latitude", NLAT), f.createDimension("longitude", NLON)
        f.createVariable("Tair", "f4", ("time", "latitude", "longitude"))[:] = t2m
    return pr.reshape(DAYS * 24, -1), t2m.reshape(DAYS, -1), (lat_i * NLON + lon_i).astype(np.int32)


def config_text(start, end, origin=None, mode="cold", init_file="", states="[0, 2, 4]", prefix=""):
    origin_line = f'  origin: "{origin}"\n' if origin else ""
    init_line = f'  file: "{init_file}"\n' if mode == "hot" else ""
    return f"""\
model:
  uid: 204
  name: Model204
time:
  start: "{start}"
  end:   "{end}"
{origin_line}initial:
  mode: {mode}
{init_line}local_params:
  file: "params/links.csv"
forcings:
  type:    folder_nc
  path:    "forc"
  lookup:  "lookup.csv"
  vars:
    precipitation: "PRCP"
    temperature:   "Tair"
output:
  print_interval: "1h"
  states: {states}
  dir: "out"
  prefix: "{prefix}"
solver:
  method: RK45
  tolerances:
    rtol:      1e-6
    atol:      1e-9
    safety:    0.9
    min_scale: 0.2
    max_scale: 10.0
  initial_step: null
  max_attempts: 2000000
mpi:
  step_storage: 30
  transfer_buffer: 10
  discontinuity_buf: 0
"""


def run_driver(d, cfg_name, world=1):
    subprocess.check_call(["make", "-C", HOST, "build/hlm_run"], stdout=subprocess.DEVNULL)
    for r in range(world):
        out = subprocess.run([RUN, str(d / cfg_name), "--rank", str(r), "--world", str(world), "--device", "0"],
                             capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stdout + out.stderr
        assert "done:" in out.stdout


from tests.ncread import is_hdf5, read_nc  # noqa: E402  (NetCDF-4 through the repo's reader, classic through SciPy)


def expected(sp, col, pr, t2m, day_lo, day_hi, y0):
    """The same run through the Python mirror of the C ABI, one interval per day (hlm_solve_restart keeps
    links that were flagged stiff flagged, exactly what the driver's session does)."""
    ns = len(sp)
    dense = np.zeros((ns, (day_hi - day_lo) * 24 + 1, 5))
    with hlm.Solver(0) as s:
        s.set_model_parameters(204, hlm.Parameters(initialStep=1e-6))
        s.set_max_attempts(2_000_000)
        s.upload_spatial_params(sp)
        s.upload_forcing(0, 1.0, pr)
        s.upload_forcing(1, 24.0, t2m)
        s.set_forcing_columns(col)
        for k in range(day_lo, day_hi):
            tq = k * 1440.0 + 60.0 * np.arange(1, 25)
            if k == day_lo:  # the first interval also owns the (never written, F10) query at its start
                tq = np.concatenate([[k * 1440.0], tq])
                s.solve_begin(204, y0, k * 1440.0, (k + 1) * 1440.0, tq)
            else:
                s.solve_restart(k * 1440.0, (k + 1) * 1440.0, tq)
            s.solve_window(len(tq))
            win = np.zeros((ns, len(tq), 5))
            s.solve_fetch_window(win)
            s.synchronize()
            q0 = (k - day_lo) * 24 + (0 if k == day_lo else 1)
            dense[:, q0:q0 + len(tq)] = win
        r = s.solve_end()
    return r["final"], dense, r["stiff"]


def test_driver_two_ranks_matches_python_api_and_oracle(tmp_path):
    pr, t2m, col = write_case(tmp_path)
    (tmp_path / "config.yaml").write_text(config_text("2021-01-01T00:00:00", "2021-01-04T00:00:00"))
    run_driver(tmp_path, "config.yaml", world=2)
    sp_all = hostio.load_spatial_params(str(tmp_path / "params" / "links.csv"))
    from tiger_hlm_gpu_b200.sharding import shard_range
    y0c = np.tile([0.01, 3.0, 0.0, 5.0, 0.2], (NS, 1))
    saw_wet = False
    for rank in range(2):
        lo, hi = shard_range(NS, 2, rank)
        fin = read_nc(tmp_path / "out" / f"final_rank_{rank}.nc")
        den = read_nc(tmp_path / "out" / f"dense_rank_{rank}.nc")
        # the default container is the reference's: NetCDF-4 / HDF5, deflate + shuffle (I_O/output_series.cpp:31,56)
        assert is_hdf5(tmp_path / "out" / f"final_rank_{rank}.nc") and is_hdf5(tmp_path / "out" / f"dense_rank_{rank}.nc")
        assert np.array_equal(fin["system"], sp_all["stream"][lo:hi])      # real link ids
        assert np.array_equal(den["variable"], [0, 2, 4])                  # output.states
        assert np.array_equal(den["time"], 60.0 * np.arange(DAYS * 24 + 1))
        y, dense, stiff = expected(sp_all[lo:hi], col[lo:hi], pr, t2m, 0, DAYS, y0c[lo:hi])
        assert np.array_equal(fin["outputs"], y)
        assert np.array_equal(den["outputs"], dense[:, :, [0, 2, 4]])
        assert not den["outputs"][:, 0].any()                              # t = t0 is never written
        saw_wet |= bool((dense[:, :, 2] > 0).any())
        if rank == 0:  # and the CPU oracle, chained the same way (links it never flags stiff)
            yo = y0c[lo:hi]
            ok = np.ones(hi - lo, bool)
            for k in range(DAYS):
                o = O.run_rk45(204, O.Params.make(initialStep=1e-6), yo, k * 1440.0, (k + 1) * 1440.0,
                               k * 1440.0 + 60.0 * np.arange(1, 25), sp=sp_all[lo:hi],
                               forcing=O.Forcing([pr, t2m], [1.0, 24.0], col=col[lo:hi]), threads=os.cpu_count() or 1,
                               device_pow=True)
                ok &= o["stiff"] == 0
                assert np.array_equal(o["dense"][ok], dense[ok, 1 + 24 * k:25 + 24 * k])
                yo = o["final"]
            assert np.array_equal(yo[ok], y[ok]) and np.array_equal(~ok, stiff != 0) and ok.sum() > 50
    assert saw_wet, "the case was meant to exercise the surface-storage branch"


def test_driver_hot_start_continues_a_run(tmp_path):
    """Days 1-2, then day 3 restarted from the final-state file: same states as the uninterrupted run."""
    pr, t2m, col = write_case(tmp_path)
    (tmp_path / "full.yaml").write_text(config_text("2021-01-01T00:00:00", "2021-01-04T00:00:00", states="[0, 1, 2, 3, 4]", prefix="full_"))
    # days 1-2 in the classic container (output.format: netcdf3), the rest in the default NetCDF-4: hot start reads either
    (tmp_path / "a.yaml").write_text(config_text("2021-01-01T00:00:00", "2021-01-03T00:00:00", states="[0, 1, 2, 3, 4]", prefix="a_")
                                     .replace('  dir: "out"', '  dir: "out"\n  format: netcdf3'))
    (tmp_path / "b.yaml").write_text(config_text("2021-01-03T00:00:00", "2021-01-04T00:00:00", origin="2021-01-01T00:00:00", mode="hot",
                                                 init_file="out/a_final_rank_0.nc", states="[0, 1, 2, 3, 4]", prefix="b_"))
    for c in ("full.yaml", "a.yaml", "b.yaml"):
        run_driver(tmp_path, c)
    full_f, full_d = read_nc(tmp_path / "out" / "full_final_rank_0.nc"), read_nc(tmp_path / "out" / "full_dense_rank_0.nc")
    b_f, b_d = read_nc(tmp_path / "out" / "b_final_rank_0.nc"), read_nc(tmp_path / "out" / "b_dense_rank_0.nc")
    # links flagged stiff during days 1-2 have no final state (zero rows, solver/rk45_kernel.cu:167-170):
    # the uninterrupted run keeps them flagged, a restart would start them from zeros
    a_f = read_nc(tmp_path / "out" / "a_final_rank_0.nc")
    assert not is_hdf5(tmp_path / "out" / "a_final_rank_0.nc") and is_hdf5(tmp_path / "out" / "b_final_rank_0.nc")
    ok = a_f["outputs"].any(axis=1)
    assert 50 < ok.sum() < NS
    assert np.array_equal(b_f["system"], full_f["system"])
    assert np.array_equal(b_f["outputs"][ok], full_f["outputs"][ok])
    assert np.array_equal(b_d["time"], 2880.0 + 60.0 * np.arange(25))
    assert np.array_equal(b_d["outputs"][ok, 1:], full_d["outputs"][ok, 49:])


def test_driver_reports_errors_like_main(tmp_path):
    write_case(tmp_path)
    (tmp_path / "bad.yaml").write_text(config_text("2021-01-01T00:00:00", "2021-01-02T00:00:00").replace('"PRCP"', '"nope"'))
    subprocess.check_call(["make", "-C", HOST, "build/hlm_run"], stdout=subprocess.DEVNULL)
    out = subprocess.run([RUN, str(tmp_path / "bad.yaml")], capture_output=True, text=True, timeout=120)
    assert out.returncode == 1 and "nope" in out.stderr


def test_driver_routed_model200(tmp_path):
    """routing.enabled: Model 200 links coupled along next_stream (here one chain of 203 links), coupling
    interval 30 min, implicit fallback on; against the Python mirror's RoutedSolver, bit for bit."""
    from tiger_hlm_gpu_b200 import routing
    pr, t2m, col = write_case(tmp_path)
    cfg = config_text("2021-01-01T00:00:00", "2021-01-02T00:00:00", states="[0, 1, 2, 3, 4]", prefix="r_")
    cfg = cfg.replace("uid: 204", "uid: 200").replace("name: Model204", "name: Model200")
    cfg += 'routing:\n  enabled: true\n  couple: "30m"\n  subbasin_links: 64\n'
    (tmp_path / "routed.yaml").write_text(cfg)
    run_driver(tmp_path, "routed.yaml")
    fin, den = read_nc(tmp_path / "out" / "r_final_rank_0.nc"), read_nc(tmp_path / "out" / "r_dense_rank_0.nc")
    sp = hostio.load_spatial_params(str(tmp_path / "params" / "links.csv"))
    p1 = routing.plan(sp["stream"], sp["next_stream"], 1, subbasin_links=64)
    y0 = np.tile([0.5, 3.0, 0.0, 5.0, 0.2], (NS, 1))
    tq_all = 60.0 * np.arange(25)
    dense = np.zeros((NS, 25, 5))
    with hlm.Solver(0) as s:
        s.set_model_parameters(200, hlm.Parameters(initialStep=1e-6))
        s.set_max_attempts(2_000_000)
        s.upload_spatial_params(sp)
        s.upload_forcing(0, 1.0, pr)
        s.upload_forcing(1, 24.0, t2m)
        s.set_forcing_columns(col)
        rs = routing.RoutedSolver(s, 200, p1.ranks[0], 1, 0)
        for k in range(48):
            ta, tb = 30.0 * k, 30.0 * (k + 1)
            sel = np.flatnonzero((tq_all > ta) & (tq_all <= tb)) if k else np.flatnonzero(tq_all <= tb)
            tq = tq_all[sel]
            if k == 0:
                rs.begin(y0, ta, tb, tq)
            rs.advance(tb, tq)
            if len(tq):
                win = np.zeros((NS, len(tq), 5))
                s.solve_wait_copy(s.solve_fetch_window_packed(win))
                dense[:, sel] = win
        r = rs.end()
    assert np.isin(r["stiff"], (0, 3)).all()
    assert np.array_equal(fin["outputs"], r["final"])
    assert np.array_equal(den["outputs"], dense)
    q = r["final"][:, 0]
    assert q[-1] > 20 * q[0]                     # the chain outlet carries the discharge of everything above it


def test_driver_routed_two_ranks_equal_one_rank(tmp_path):
    """routing.enabled with WORLD_SIZE=2: two hlm_run processes, one per GPU, sub-basins dealt to the ranks, the boundary
    links' discharge all-gathered over NCCL once per coupling interval (hlm_nccl.hpp).  Per link id the two ranks'
    files must hold the bits of the single-rank run."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (one hlm_run process per GPU)")
    write_case(tmp_path)
    cfg = config_text("2021-01-01T00:00:00", "2021-01-01T06:00:00", states="[0, 1, 2, 3, 4]", prefix="r_")
    cfg = cfg.replace("uid: 204", "uid: 200").replace("name: Model204", "name: Model200")
    cfg += 'routing:\n  enabled: true\n  couple: "30m"\n  subbasin_links: 16\n'
    (tmp_path / "routed.yaml").write_text(cfg)
    run_driver(tmp_path, "routed.yaml")
    one_f, one_d = read_nc(tmp_path / "out" / "r_final_rank_0.nc"), read_nc(tmp_path / "out" / "r_dense_rank_0.nc")
    os.rename(tmp_path / "out", tmp_path / "out_one")
    os.makedirs(tmp_path / "out")
    env = dict(os.environ, WORLD_SIZE="2", HLM_RUN_TOKEN=f"t{os.getpid()}")
    procs = [subprocess.Popen([RUN, str(tmp_path / "routed.yaml"), "--rank", str(r), "--world", "2", "--device", str(r)],
                              env=dict(env, RANK=str(r), LOCAL_RANK=str(r)), stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=300) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "boundary links" in outs[0][0] and "cut edges" in outs[0][0]
    ids, fin, den = [], [], []
    for r in range(2):
        f, d = read_nc(tmp_path / "out" / f"r_final_rank_{r}.nc"), read_nc(tmp_path / "out" / f"r_dense_rank_{r}.nc")
        ids.append(f["system"]); fin.append(f["outputs"]); den.append(d["outputs"])
    ids, fin, den = np.concatenate(ids), np.concatenate(fin), np.concatenate(den)
    assert len(ids) == NS and len(set(ids.tolist())) == NS
    order, order_one = np.argsort(ids), np.argsort(one_f["system"])
    assert np.array_equal(fin[order], one_f["outputs"][order_one])
    assert np.array_equal(den[order], one_d["outputs"][order_one])
    assert not list((tmp_path / "out").glob("nccl_id*"))  # the id file is gone once the communicator exists


def test_driver_single_process_owning_several_contexts(tmp_path):
    """--devices D0,D1: one process, one host thread and one device context per entry, thread r = rank r (the
    reference's one MPI rank per GPU folded into one process).  Here both entries name GPU 0 — two contexts at once on
    one device — and the files must equal those of two separate processes."""
    write_case(tmp_path)
    (tmp_path / "config.yaml").write_text(config_text("2021-01-01T00:00:00", "2021-01-02T00:00:00"))
    run_driver(tmp_path, "config.yaml", world=2)
    two = {r: (read_nc(tmp_path / "out" / f"final_rank_{r}.nc"), read_nc(tmp_path / "out" / f"dense_rank_{r}.nc")) for r in range(2)}
    os.rename(tmp_path / "out", tmp_path / "out_two_processes")
    os.makedirs(tmp_path / "out")
    out = subprocess.run([RUN, str(tmp_path / "config.yaml"), "--devices", "0,0"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    for r in range(2):
        f, d = read_nc(tmp_path / "out" / f"final_rank_{r}.nc"), read_nc(tmp_path / "out" / f"dense_rank_{r}.nc")
        for a, b in ((f, two[r][0]), (d, two[r][1])):
            assert a.keys() == b.keys()
            for k in a:
                assert np.array_equal(a[k], b[k]), (r, k)
    # a routed run cannot be folded this way: the ranks need their collective
    cfg = config_text("2021-01-01T00:00:00", "2021-01-01T02:00:00").replace("uid: 204", "uid: 200") + 'routing:\n  enabled: true\n'
    (tmp_path / "routed.yaml").write_text(cfg)
    out = subprocess.run([RUN, str(tmp_path / "routed.yaml"), "--devices", "0,0"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 1 and "one process per GPU" in out.stderr
