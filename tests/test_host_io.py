"""Host-side I/O contracts of the path (no GPU): NetCDF reading/writing and the config.yaml loader.

What pins them
  * writer (reference: I_O/output_series.cpp:18-124) — files are read back with SciPy's independent
    NetCDF implementation and must carry the reference's dimension/variable/attribute names;
  * classic reader (reference: I_O/forcing_loader.cpp:67-218 over netcdf-c) — files written by SciPy
    (CDF-1 and CDF-2, fixed and record variables, short/int/float/double) must read back exactly;
  * NetCDF-4/HDF5 reader — the reference's own committed outputs src/final_example.nc and
    src/dense_example.nc (copied as data fixtures into tests/golden/) must decode to the values
    tests/golden/make_golden.py extracted from their raw chunks;
  * load_config (reference: I_O/config_loader.cpp:19-84) — the reference's data/config.yaml
    (restated below, and read from /root/reference when that tree is present).
"""
import ctypes as C
import json
import os
import subprocess

import numpy as np
import pytest
from scipy.io import netcdf_file

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "tiger_hlm_gpu_b200", "host")
GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def io():
    so = os.path.join(HOST, "build", "libhlm_hostio.so")
    subprocess.check_call(["make", "-C", HOST, "build/libhlm_hostio.so"], stdout=subprocess.DEVNULL)
    lib = C.CDLL(so)
    lib.hlmio_last_error.restype = C.c_char_p
    lib.hlmio_variables.restype = C.c_char_p
    lib.hlmio_load_config_json.restype = C.c_char_p
    lib.hlmio_load_time_chunk.argtypes = [C.c_char_p, C.c_char_p, C.c_longlong, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.hlmio_read_double.argtypes = [C.c_char_p, C.c_char_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    lib.hlmio_inquire.argtypes = [C.c_char_p, C.c_char_p, C.c_void_p, C.c_void_p]
    return lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


NC4, CLASSIC = 0, 1  # the `format` argument of the hlmio_write_* entry points


def read_var(io, path, name, start=None, count=None):
    shape = (C.c_longlong * 8)()
    es = C.c_int()
    rank = io.hlmio_inquire(path.encode(), name.encode(), shape, C.byref(es))
    assert rank >= 0, io.hlmio_last_error().decode()
    full = [shape[k] for k in range(rank)]
    start = list(start) if start is not None else [0] * rank
    count = list(count) if count is not None else [f - s for f, s in zip(full, start)]
    out = np.zeros(count, np.float64)
    s = (C.c_longlong * max(rank, 1))(*start)
    c = (C.c_longlong * max(rank, 1))(*count)
    rc = io.hlmio_read_double(path.encode(), name.encode(), s, c, rank, _p(out))
    assert rc == 0, io.hlmio_last_error().decode()
    return out


# ---------------------------------------------------------------------------------- writer
def test_write_dense_and_final_netcdf_read_back_by_scipy(io, tmp_path):
    rng = np.random.default_rng(1)
    ns, nq, n = 7, 11, 5
    dense = rng.standard_normal((ns, nq, n))
    final = rng.standard_normal((ns, n))
    t = 60.0 * np.arange(nq)
    ids = (420000000 + np.arange(ns)).astype(np.int32)
    states = np.arange(n, dtype=np.int32)
    pd, pf = str(tmp_path / "dense.nc"), str(tmp_path / "final.nc")
    io.hlmio_write_dense_netcdf(pd.encode(), _p(dense), _p(t), _p(ids), _p(states), nq, ns, n, CLASSIC, 0)
    io.hlmio_write_final_netcdf(pf.encode(), _p(final), _p(ids), _p(states), ns, n, CLASSIC, 0)
    with netcdf_file(pd, "r", mmap=False) as f:
        assert f.version_byte == 2
        assert {k: v for k, v in f.dimensions.items()} == {"system": ns, "time": nq, "variable": n}
        assert f.variables["outputs"].dimensions == ("system", "time", "variable")
        assert np.array_equal(f.variables["outputs"][:], dense)
        assert np.array_equal(f.variables["system"][:], ids)
        assert np.array_equal(f.variables["time"][:], t)
        assert np.array_equal(f.variables["variable"][:], states)
        # attributes of I_O/output_series.cpp:42-47
        assert f.variables["system"].long_name == b"LinkID"
        assert f.variables["time"].long_name == b"Time"
        assert f.variables["time"].units == b"minutes since start of simulation"
        assert f.variables["variable"].long_name == b"state variable"
        assert f.variables["variable"].units == b"various units"
    with netcdf_file(pf, "r", mmap=False) as f:
        assert f.variables["outputs"].dimensions == ("system", "variable")
        assert np.array_equal(f.variables["outputs"][:], final)
        assert f.variables["system"].long_name == b"LinkID"
    # and our own reader agrees
    assert np.array_equal(read_var(io, pd, "outputs"), dense)
    assert np.array_equal(read_var(io, pf, "outputs", start=[2, 1], count=[3, 2]), final[2:5, 1:3])


@pytest.mark.parametrize("qw", [1, 4, 11, 50])
def test_windowed_writer_equals_whole_array_writer(io, tmp_path, qw):
    rng = np.random.default_rng(2)
    ns, nq, n = 33, 11, 5
    dense = rng.standard_normal((ns, nq, n))
    t = 15.0 * np.arange(nq)
    ids = np.arange(100, 100 + ns, dtype=np.int32)
    states = np.array([4, 0, 2], np.int32)  # output.states subset, in the caller's order
    p = str(tmp_path / f"win{qw}.nc")
    rc = io.hlmio_write_dense_windows(p.encode(), _p(dense), _p(t), _p(ids), _p(states), len(states), nq, ns, n, qw, CLASSIC, 0, 0, 0)
    assert rc == 0, io.hlmio_last_error().decode()
    with netcdf_file(p, "r", mmap=False) as f:
        assert f.variables["outputs"].shape == (ns, nq, 3)
        assert np.array_equal(f.variables["outputs"][:], dense[:, :, states])
        assert np.array_equal(f.variables["variable"][:], states)
        assert np.array_equal(f.variables["time"][:], t)


def test_windowed_writer_rejects_bad_state(io, tmp_path):
    d = np.zeros((2, 2, 5))
    t = np.arange(2.0)
    ids = np.arange(2, dtype=np.int32)
    states = np.array([5], np.int32)
    rc = io.hlmio_write_dense_windows(str(tmp_path / "x.nc").encode(), _p(d), _p(t), _p(ids), _p(states), 1, 2, 2, 5, 1, NC4, 4, 0, 0)
    assert rc != 0 and b"state index" in io.hlmio_last_error()


# ---------------------------------------------------------------------------------- classic reader
@pytest.mark.parametrize("version,record", [(1, False), (2, False), (1, True), (2, True)])
def test_classic_reader_matches_scipy_written_files(io, tmp_path, version, record):
    rng = np.random.default_rng(3)
    nt, nlat, nlon = 30, 4, 6
    p = str(tmp_path / f"f{version}{record}.nc")
    pr = rng.random((nt, nlat, nlon)).astype(np.float32)
    packed = rng.integers(-3000, 3000, (nt, nlat, nlon)).astype(np.int16)
    with netcdf_file(p, "w", version=version) as f:
        f.createDimension("time", None if record else nt)
        f.createDimension("latitude", nlat)
        f.createDimension("longitude", nlon)
        tv = f.createVariable("time", "i4", ("time",))
        tv.units = "hours since 2019-01-01 00:00:00"
        tv[:] = np.arange(nt) * 3
        lat = f.createVariable("latitude", "f8", ("latitude",))
        lat[:] = np.linspace(40, 41, nlat)
        v = f.createVariable("pr", "f4", ("time", "latitude", "longitude"))
        v[:] = pr
        s = f.createVariable("t2m_packed", "i2", ("time", "latitude", "longitude"))
        s.scale_factor = 0.01
        s[:] = packed
    assert set(io.hlmio_variables(p.encode()).decode().split()) == {"time", "latitude", "pr", "t2m_packed"}
    assert np.array_equal(read_var(io, p, "pr"), pr.astype(np.float64))
    assert np.array_equal(read_var(io, p, "latitude"), np.linspace(40, 41, nlat))
    assert np.array_equal(read_var(io, p, "pr", start=[7, 1, 2], count=[5, 2, 3]), pr[7:12, 1:3, 2:5].astype(np.float64))
    # integer storage converts like nc_get_vara_float: the stored numbers, no scale_factor applied
    assert np.array_equal(read_var(io, p, "t2m_packed"), packed.astype(np.float64))
    # NetCDFLoader (forcing_loader.cpp): dims, chunk, dt from the time coordinate
    dims = (C.c_longlong * 3)()
    dt = C.c_double()
    chunk = np.zeros((6, nlat, nlon), np.float32)
    rc = io.hlmio_load_time_chunk(p.encode(), b"pr", 10, 6, _p(chunk), dims, C.byref(dt))
    assert rc == 0, io.hlmio_last_error().decode()
    assert list(dims) == [nt, nlat, nlon] and dt.value == 3.0
    assert np.array_equal(chunk, pr[10:16])


def test_netcdf_loader_error_behaviour(io, tmp_path):
    """Same failure modes and messages as forcing_loader.cpp:77-107,165-176."""
    p = str(tmp_path / "g.nc")
    with netcdf_file(p, "w") as f:
        f.createDimension("time", 4)
        f.createDimension("y", 2)
        f.createVariable("flat", "f4", ("time", "y"))[:] = np.zeros((4, 2), np.float32)
        f.createDimension("x", 3)
        f.createVariable("pr", "f4", ("time", "y", "x"))[:] = np.zeros((4, 2, 3), np.float32)
    buf = np.zeros(1000, np.float32)
    assert io.hlmio_load_time_chunk(p.encode(), b"nope", 0, 1, _p(buf), None, None) != 0
    assert io.hlmio_last_error() == b"Variable nope not found in file"
    assert io.hlmio_load_time_chunk(p.encode(), b"flat", 0, 1, _p(buf), None, None) != 0
    assert io.hlmio_last_error() == b"Expected 3D variable (time, lat, lon), got 2D"
    assert io.hlmio_load_time_chunk(p.encode(), b"pr", 0, 0, _p(buf), None, None) != 0
    assert io.hlmio_last_error() == b"Size of time chunk must be greater than zero"
    assert io.hlmio_load_time_chunk(p.encode(), b"pr", 4, 1, _p(buf), None, None) != 0
    assert io.hlmio_last_error() == b"Start time index out of range"
    assert io.hlmio_load_time_chunk(p.encode(), b"pr", 2, 3, _p(buf), None, None) != 0
    assert io.hlmio_last_error() == b"Requested time steps exceed available data"
    assert io.hlmio_load_time_chunk(str(tmp_path / "missing.nc").encode(), b"pr", 0, 1, _p(buf), None, None) != 0
    assert io.hlmio_last_error().startswith(b"Opening file")
    junk = tmp_path / "junk.nc"
    junk.write_bytes(b"not a netcdf file at all")
    assert io.hlmio_load_time_chunk(str(junk).encode(), b"pr", 0, 1, _p(buf), None, None) != 0


# ---------------------------------------------------------------------------------- NetCDF-4 / HDF5 reader
def test_hdf5_reader_decodes_the_reference_outputs(io):
    g = np.load(os.path.join(GOLD, "model204_example.npz"))
    pf, pd = os.path.join(GOLD, "final_example.nc"), os.path.join(GOLD, "dense_example.nc")
    assert set(io.hlmio_variables(pf.encode()).decode().split()) == {"system", "variable", "outputs"}
    assert set(io.hlmio_variables(pd.encode()).decode().split()) == {"system", "time", "variable", "outputs"}
    final = read_var(io, pf, "outputs")
    assert final.shape == (10, 5) and np.array_equal(final, g["final"])
    dense = read_var(io, pd, "outputs")
    assert dense.shape == (10, 2881, 5)
    for s in range(10):
        assert np.array_equal(dense[s], g["dense_sys0"])
    assert np.array_equal(read_var(io, pd, "time"), np.arange(2881.0))
    assert np.array_equal(read_var(io, pd, "system"), np.arange(10.0))  # the reference writes 0..ns-1 (main.cpp:788-793)
    # hyperslabs cut through the single compressed chunk
    assert np.array_equal(read_var(io, pd, "outputs", start=[3, 100, 1], count=[2, 50, 3]), dense[3:5, 100:150, 1:4])


# ---------------------------------------------------------------------------------- NetCDF-4 / HDF5 writer
class H5:
    """A minimal walk over an HDF5 file of the kind netcdf-c writes (superblock v2, version-2 object headers, compact
    links): just enough to lay the reference's own files and this repo's side by side, message by message."""
    NAMES = {1: "dataspace", 2: "linkinfo", 3: "datatype", 5: "fill", 6: "link", 8: "layout", 0xa: "groupinfo", 0xb: "filter",
             0xc: "attr", 0x10: "cont", 0x15: "attrinfo", 0: "nil"}

    def __init__(self, path):
        self.b = open(path, "rb").read()
        b = self.b
        assert b[:8] == b"\x89HDF\r\n\x1a\n"
        self.superblock_version, self.size_of_offsets, self.size_of_lengths = b[8], b[9], b[10]
        self.base, self.ext, self.eof, self.root = np.frombuffer(b[12:44], "<u8").tolist()
        self.superblock_checksum = int.from_bytes(b[44:48], "little")
        self.objects = {}
        for t, body in self.messages(self.root):
            if t == 6:
                fl, q = body[1], 2
                q += (1 if fl & 8 else 0) + (8 if fl & 4 else 0) + (1 if fl & 0x10 else 0)
                ls = 1 << (fl & 3)
                ln = int.from_bytes(body[q:q + ls], "little")
                q += ls
                self.objects[body[q:q + ln].decode()] = int.from_bytes(body[q + ln:q + ln + 8], "little")

    def chunks_of(self, addr):
        """(start, end) of every header chunk of the object at addr, checksum excluded: [(lo, hi, stored_checksum)]"""
        b = self.b
        assert b[addr:addr + 4] == b"OHDR" and b[addr + 4] == 2
        fl = b[addr + 5]
        p = addr + 6 + (16 if fl & 0x20 else 0) + (4 if fl & 0x10 else 0)
        n = 1 << (fl & 3)
        size0 = int.from_bytes(b[p:p + n], "little")
        p += n
        return fl, p, size0

    def messages(self, addr):
        b = self.b
        fl, p, size0 = self.chunks_of(addr)
        blocks, out = [(p, p + size0)], []
        self.checks = getattr(self, "checks", [])
        self.checks.append((addr, p + size0))
        while blocks:
            p, end = blocks.pop(0)
            while p + 4 <= end:
                t, sz = b[p], int.from_bytes(b[p + 1:p + 3], "little")
                hp = p + 4 + (2 if fl & 4 else 0)
                body = b[hp:hp + sz]
                out.append((t, body))
                if t == 0x10:
                    ca, cl = int.from_bytes(body[:8], "little"), int.from_bytes(body[8:16], "little")
                    assert b[ca:ca + 4] == b"OCHK"
                    blocks.append((ca + 4, ca + cl - 4))
                    self.checks.append((ca, ca + cl - 4))
                p = hp + sz
        return out

    def by_type(self, name):
        msgs = self.messages(self.objects[name])
        return {self.NAMES.get(t, hex(t)): body for t, body in msgs if t not in (0, 0xc, 0x10)}

    def attr_names(self, name):
        out = []
        for t, body in self.messages(self.objects[name]):
            if t == 0xc:
                n = int.from_bytes(body[2:4], "little")
                out.append(body[9:9 + n - 1].decode())
        return out


def test_lookup3_matches_the_checksums_in_the_reference_files(io):
    """HDF5's metadata checksum (Jenkins lookup3), as implemented for the writer, against the checksums netcdf-c/HDF5
    stored in the reference's own outputs: the superblock's and every object-header chunk's."""
    io.hlmio_lookup3.restype = C.c_uint
    io.hlmio_lookup3.argtypes = [C.c_char_p, C.c_longlong]
    for f in ("final_example.nc", "dense_example.nc"):
        h = H5(os.path.join(GOLD, f))
        assert io.hlmio_lookup3(h.b[:44], 44) == h.superblock_checksum
        for name in h.objects:
            h.messages(h.objects[name])
        assert len(h.checks) >= 6
        for lo, hi in h.checks:
            assert io.hlmio_lookup3(h.b[lo:hi], hi - lo) == int.from_bytes(h.b[hi:hi + 4], "little"), (f, lo)


def test_netcdf4_writer_lays_the_file_out_like_the_reference(io, tmp_path):
    """The reference's outputs are NetCDF-4 with shuffle + deflate on `outputs` (I_O/output_series.cpp:31,56,88,109).  The
    values of src/final_example.nc and src/dense_example.nc written by this repo's NetCDF-4 writer give files with the same
    superblock version, the same dataspace / datatype / fill / filter-pipeline messages byte for byte, the same layout class,
    version and chunk shape (one chunk), the same attribute names in the same order; they decode to the same values through
    the reader that decodes the reference's, and every checksum in them verifies."""
    io.hlmio_lookup3.restype = C.c_uint
    io.hlmio_lookup3.argtypes = [C.c_char_p, C.c_longlong]
    g = np.load(os.path.join(GOLD, "model204_example.npz"))
    final = np.ascontiguousarray(g["final"])
    dense = np.ascontiguousarray(np.repeat(g["dense_sys0"][None], 10, axis=0))
    ids, states, t = np.arange(10, dtype=np.int32), np.arange(5, dtype=np.int32), np.arange(2881.0)
    pf, pd = str(tmp_path / "final.nc"), str(tmp_path / "dense.nc")
    io.hlmio_write_final_netcdf(pf.encode(), _p(final), _p(ids), _p(states), 10, 5, NC4, 4)
    io.hlmio_write_dense_netcdf(pd.encode(), _p(dense), _p(t), _p(ids), _p(states), 2881, 10, 5, NC4, 4)
    for mine, ref in ((pf, "final_example.nc"), (pd, "dense_example.nc")):
        a, r = H5(mine), H5(os.path.join(GOLD, ref))
        assert (a.superblock_version, a.size_of_offsets, a.size_of_lengths) == (r.superblock_version, r.size_of_offsets, r.size_of_lengths) == (2, 8, 8)
        assert a.eof == len(a.b) and a.base == 0
        assert list(a.objects) == list(r.objects) or sorted(a.objects) == sorted(r.objects)
        for name in r.objects:
            ma, mr = a.by_type(name), r.by_type(name)
            for kind in ("dataspace", "datatype", "fill", "filter"):
                assert ma.get(kind) == mr.get(kind), (ref, name, kind)
            la, lr = ma["layout"], mr["layout"]
            assert la[:2] == lr[:2]                      # version 3, class (1 contiguous / 2 chunked)
            if la[1] == 2:
                assert la[2] == lr[2] and la[11:] == lr[11:]  # dimensionality, chunk shape + element size (the B-tree address differs)
            else:
                assert la[10:] == lr[10:]                # size in bytes
            assert a.attr_names(name) == r.attr_names(name), (ref, name)
        a.checks = []
        for name in a.objects:
            a.messages(a.objects[name])
        a.messages(a.root)
        for lo, hi in a.checks:
            assert io.hlmio_lookup3(a.b[lo:hi], hi - lo) == int.from_bytes(a.b[hi:hi + 4], "little")
        assert io.hlmio_lookup3(a.b[:44], 44) == a.superblock_checksum
    assert np.array_equal(read_var(io, pf, "outputs"), final)
    assert np.array_equal(read_var(io, pd, "outputs"), dense)
    assert np.array_equal(read_var(io, pd, "time"), t) and np.array_equal(read_var(io, pd, "system"), ids)
    assert os.path.getsize(pd) < dense.nbytes // 4  # deflate + shuffle did their work (the reference's file: 114 009 bytes)
    # level 0: no filter, contiguous (output_series.cpp:56 only deflates when the level is positive)
    p0 = str(tmp_path / "final0.nc")
    io.hlmio_write_final_netcdf(p0.encode(), _p(final), _p(ids), _p(states), 10, 5, NC4, 0)
    m0 = H5(p0).by_type("outputs")
    assert "filter" not in m0 and m0["layout"][1] == 1 and np.array_equal(read_var(io, p0, "outputs"), final)


@pytest.mark.parametrize("ns,nq,qw,staging,f32", [(33, 11, 4, 0, 0), (33, 11, 1, 2000, 0), (700, 40, 7, 9000, 0), (700, 40, 40, 0, 1),
                                                   (5, 300, 13, 600, 1), (3000, 130, 50, 100000, 0)])
def test_netcdf4_windowed_writer_many_chunks(io, tmp_path, ns, nq, qw, staging, f32):
    """DenseSeriesWriter's NetCDF-4 container: windows of any length into chunk-rows of queries (a small staging slab forces
    many rows, partial last rows and system chunks that overhang), a multi-level chunk B-tree, float or double values,
    output.states in the caller's order; read back whole and by hyperslab through the HDF5 reader."""
    rng = np.random.default_rng(ns + nq)
    n = 5
    dense = rng.standard_normal((ns, nq, n))
    t = 15.0 * np.arange(nq)
    ids = np.arange(100, 100 + ns, dtype=np.int32)
    states = np.array([4, 0, 2], np.int32)
    p = str(tmp_path / "win.nc")
    rc = io.hlmio_write_dense_windows(p.encode(), _p(dense), _p(t), _p(ids), _p(states), len(states), nq, ns, n, qw, NC4, 4, f32, staging)
    assert rc == 0, io.hlmio_last_error().decode()
    want = dense[:, :, states]
    if f32:
        want = want.astype(np.float32).astype(np.float64)
    got = read_var(io, p, "outputs")
    assert got.shape == (ns, nq, 3) and np.array_equal(got, want)
    assert np.array_equal(read_var(io, p, "variable"), states) and np.array_equal(read_var(io, p, "time"), t)
    s0, q0 = ns // 3, nq // 4
    assert np.array_equal(read_var(io, p, "outputs", start=[s0, q0, 1], count=[ns - s0, nq - q0, 2]), want[s0:, q0:, 1:])
    h = H5(p)
    lay = h.by_type("outputs")["layout"]
    assert lay[1] == 2 and h.by_type("outputs")["datatype"][4] == (4 if f32 else 8)


# ---------------------------------------------------------------------------------- config.yaml
REFERENCE_CONFIG = """\
# simulation.yaml
model:
  uid: 204                        # Model UID
  name: Model204                  # Optional human-readable name
time:
  start: "2021-01-01T00:00:00"    # ISO8601 for t=0
  end:   "2021-10-01T00:00:00"     # for output control only
initial:
  mode: hot                       # "cold" or "hot"
  file: "inits/inicond_204.uini"  # only if mode: hot
global_params:
  - name: foo
    value: 0.0
local_params:
  file: "params/stream_params.csv"
  columns:
    stream_id:      0
    next_stream_id: 1
    params_start:   2
    num_params:    15
forcings:
  type:    folder_nc
  path:    "/data/forcings/2021"
  lookup:  "forcings_lookup.csv"
  vars:
    precipitation: "PRCP"
    temperature:   "Tair"
    # doy is computed internally from time.start and t
output:
  print_interval: "1h"           # e.g. "15m", "1h", "1d"
  states:                        # explicitly list if you want only a subset
    - 0   # snow
    - 1   # static
    - 2   # surface
    - 3   # grav
    - 4   # aquifer
solver:
  method: RK45
  tolerances:
    rtol:      1e-6
    atol:      1e-9
    safety:    0.9
    min_scale: 0.2
    max_scale: 10.0
  initial_step: null             # null -> auto-scaled
mpi:
  step_storage:     30           # max RK steps to keep per link
  transfer_buffer: 10            # max steps sent between procs
  discontinuity_buf: 0
flags:
  uses_dam:     false
  convert_area: false
"""

EXPECTED = {
    "model.uid": 204, "model.name": "Model204", "time.start": "2021-01-01T00:00:00", "time.end": "2021-10-01T00:00:00",
    "time.minutes": 273 * 1440.0, "initial.mode": "hot", "initial.file": "inits/inicond_204.uini", "global_params": 1,
    "local_params.file": "params/stream_params.csv", "local_params.num_params": 15, "forcings.type": "folder_nc",
    "forcings.path": "/data/forcings/2021", "forcings.lookup": "forcings_lookup.csv", "forcings.precipitation": "PRCP",
    "forcings.temperature": "Tair", "output.print_interval": "1h", "output.print_minutes": 60.0,
    "output.states": [0, 1, 2, 3, 4], "solver.override_tolerances": True, "solver.rtol": 1e-6, "solver.atol": 1e-9,
    "solver.safety": 0.9, "solver.min_scale": 0.2, "solver.max_scale": 10.0, "solver.override_initial_step": False,
    "mpi.step_storage": 30, "mpi.transfer_buffer": 10, "mpi.discontinuity_buf": 0, "flags.uses_dam": False,
    "flags.convert_area": False,
}


def load_cfg(io, path):
    s = io.hlmio_load_config_json(str(path).encode())
    assert s is not None, io.hlmio_last_error().decode()
    return json.loads(s.decode())


def test_load_config_reference_schema(io, tmp_path):
    p = tmp_path / "config.yaml"
    p.write_text(REFERENCE_CONFIG)
    cfg = load_cfg(io, p)
    for k, v in EXPECTED.items():
        assert cfg[k] == v, k
    # extension keys default so that the reference's file needs no change
    assert cfg["solver.interval"] == "1d" and cfg["output.format"] == "netcdf" and cfg["output.dense"] is True
    assert cfg["solver.stiff_fallback"] is False and cfg["routing.enabled"] is False and cfg["routing.couple_minutes"] == 15.0


@pytest.mark.skipif(not os.path.exists("/root/reference/data/config.yaml"), reason="reference tree absent")
def test_load_config_reads_the_reference_file_itself(io):
    cfg = load_cfg(io, "/root/reference/data/config.yaml")
    for k, v in EXPECTED.items():
        assert cfg[k] == v, k


def test_load_config_variants_and_errors(io, tmp_path):
    base = REFERENCE_CONFIG
    # initial_step given, flow sequence for states, no tolerances block, extensions
    alt = (base.replace("initial_step: null", "initial_step: 0.01")
               .replace("  states:                        # explicitly list if you want only a subset\n"
                        "    - 0   # snow\n    - 1   # static\n    - 2   # surface\n    - 3   # grav\n    - 4   # aquifer\n",
                        "  states: [1, 3]\n  format: csv\n  dense: false\n")
               .replace("mode: hot ", "mode: cold ")
               .replace("  method: RK45\n", "  method: RK45\n  interval: 12h\n  max_attempts: 500000\n"))
    p = tmp_path / "alt.yaml"
    p.write_text(alt)
    cfg = load_cfg(io, p)
    assert cfg["solver.override_initial_step"] is True and cfg["solver.initial_step"] == 0.01
    assert cfg["output.states"] == [1, 3] and cfg["output.format"] == "csv" and cfg["output.dense"] is False
    assert cfg["initial.mode"] == "cold" and cfg["initial.file"] == ""
    assert cfg["solver.interval"] == "12h" and cfg["solver.max_attempts"] == 500000
    # routed runs and the implicit fallback (project extensions)
    (tmp_path / "routed.yaml").write_text(base.replace("  method: RK45\n", "  method: RK45\n  stiff_fallback: true\n")
                                          + 'routing:\n  enabled: true\n  couple: "30m"\n  subbasin_links: 512\n')
    cfg = load_cfg(io, tmp_path / "routed.yaml")
    assert cfg["solver.stiff_fallback"] is True and cfg["routing.enabled"] is True
    assert cfg["routing.couple_minutes"] == 30.0 and cfg["routing.subbasin_links"] == 512
    # a missing required key and a malformed time raise (std::runtime_error in the reference)
    (tmp_path / "bad1.yaml").write_text(base.replace("  uid: 204", "  id: 204"))
    assert io.hlmio_load_config_json(str(tmp_path / "bad1.yaml").encode()) is None
    assert b"model.uid" in io.hlmio_last_error()
    (tmp_path / "bad2.yaml").write_text(base.replace("2021-01-01T00:00:00", "January first"))
    assert io.hlmio_load_config_json(str(tmp_path / "bad2.yaml").encode()) is None
    assert io.hlmio_last_error() == b"Failed to parse time: January first"
    assert io.hlmio_load_config_json(str(tmp_path / "absent.yaml").encode()) is None


def test_interval_boundaries_always_advance(io):
    """hlm_run's interval loop (hlm_run.cpp) with intervals that are not a whole number of minutes: the boundary
    after k*interval must be (k+1)*interval, never k*interval again (floor(ta/interval) can round just below k)."""
    io.hlmio_interval_boundaries.restype = C.c_longlong
    io.hlmio_interval_boundaries.argtypes = [C.c_double, C.c_double, C.c_char_p, C.c_void_p, C.c_longlong]
    for text, minutes in (("10s", 10.0 / 60.0), ("20s", 20.0 / 60.0), ("0.1m", 0.1), ("0.7", 0.7), ("1h", 60.0), ("15", 15.0)):
        t0, t1 = 0.0, minutes * 500.25
        out = np.zeros(2048)
        n = io.hlmio_interval_boundaries(t0, t1, text.encode(), _p(out), out.size)
        assert n == 501, (text, n, io.hlmio_last_error().decode())
        b = out[:n]
        assert np.all(np.diff(np.concatenate([[t0], b])) > 0)
        assert b[-1] == t1
        np.testing.assert_allclose(b[:-1], minutes * np.arange(1, 501), rtol=1e-12)
    # a start inside an interval, and one exactly on a boundary
    out = np.zeros(8)
    assert io.hlmio_interval_boundaries(90.0, 200.0, b"1h", _p(out), 8) == 3 and list(out[:3]) == [120.0, 180.0, 200.0]
    assert io.hlmio_interval_boundaries(120.0, 240.0, b"1h", _p(out), 8) == 2 and list(out[:2]) == [180.0, 240.0]
    assert io.hlmio_interval_boundaries(0.0, 10.0, b"0", _p(out), 8) == -1


def test_netcdf4_writer_edge_shapes(io, tmp_path):
    """One link, one query, one state; duplicated states; a window longer than the series."""
    for ns, nq, states, qw in ((1, 1, [3], 1), (1, 7, [2, 2, 0], 50), (9, 1, [0, 1, 2, 3, 4], 1)):
        rng = np.random.default_rng(ns * 10 + nq)
        dense = rng.standard_normal((ns, nq, 5))
        t = 60.0 * np.arange(nq)
        ids = np.arange(7, 7 + ns, dtype=np.int32)
        st = np.array(states, np.int32)
        p = str(tmp_path / f"e_{ns}_{nq}.nc")
        rc = io.hlmio_write_dense_windows(p.encode(), _p(dense), _p(t), _p(ids), _p(st), len(st), nq, ns, 5, qw, NC4, 4, 0, 0)
        assert rc == 0, io.hlmio_last_error().decode()
        assert np.array_equal(read_var(io, p, "outputs"), dense[:, :, st])
        assert np.array_equal(read_var(io, p, "system"), ids) and np.array_equal(read_var(io, p, "variable"), st)
        pf = str(tmp_path / f"f_{ns}.nc")
        io.hlmio_write_final_netcdf(pf.encode(), _p(np.ascontiguousarray(dense[:, 0])), _p(ids), _p(np.arange(5, dtype=np.int32)), ns, 5, NC4, 9)
        assert np.array_equal(read_var(io, pf, "outputs"), dense[:, 0])
