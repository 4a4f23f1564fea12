"""Input/output contracts of the C++ host (tiger_hlm_gpu_b200/host/hlm_host.hpp) and its Python mirror."""
import os
import subprocess

import numpy as np
import pytest

from tiger_hlm_gpu_b200 import hostio

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SELFTEST = os.path.join(ROOT, "tiger_hlm_gpu_b200", "host", "build", "hlm_host_selftest")
GOLDEN = os.path.join(ROOT, "tests", "golden")
pytestmark = pytest.mark.skipif(not os.path.exists(SELFTEST), reason="host binaries not built (run __graft_entry__.build())")


def test_cpp_loader_matches_python_mirror_and_reference_conversions():
    out = subprocess.run([SELFTEST, "params", os.path.join(GOLDEN, "small_test.csv")], capture_output=True, text=True,
                         check=True).stdout.split("\n")
    assert int(out[0]) == 10
    sp = hostio.load_spatial_params(os.path.join(GOLDEN, "small_test.csv"))
    names = list(sp.dtype.names)
    for i in range(10):
        vals = out[1 + i].split()
        assert int(vals[0]) == sp["stream"][i] and int(vals[1]) == sp["next_stream"][i]
        for n, v in zip(names[2:], vals[2:]):
            assert float(v) == sp[n][i], n
    # I_O/parameters_loader.cpp:57,94-101
    c1 = 0.001 / 60.0
    assert np.all(sp["c1"] == c1) and np.all(sp["infil"] == 4 * c1) and np.all(sp["perco"] == 1.6 * c1)
    assert np.all(sp["alpha3"] == 2 * 24.0 * 60.0) and np.all(sp["alpha4"] == 55 * 24.0 * 60.0)
    assert sp["L"][0] == 0.304 and sp["A_h"][0] == 0.158 and sp["Hu"][0] == 178


def test_loader_errors_like_the_reference(tmp_path):
    bad = tmp_path / "bad.csv"
    bad.write_text("stream,next_stream,i2\n1,2,3\n")
    r = subprocess.run([SELFTEST, "params", str(bad)], capture_output=True, text=True)
    assert r.returncode == 1 and "Missing column 'i3'" in r.stderr
    with pytest.raises(RuntimeError, match="Missing column 'i3'"):
        hostio.load_spatial_params(str(bad))
    r = subprocess.run([SELFTEST, "params", str(tmp_path / "nope.csv")], capture_output=True, text=True)
    assert r.returncode == 1 and "Failed to open parameter file" in r.stderr
    hdr = open(os.path.join(GOLDEN, "small_test.csv")).readline()
    short = tmp_path / "short.csv"
    short.write_text(hdr + "1,2,3\n")
    r = subprocess.run([SELFTEST, "params", str(short)], capture_output=True, text=True)
    assert r.returncode == 1 and "too few fields" in r.stderr
    with pytest.raises(RuntimeError, match="too few fields"):
        hostio.load_spatial_params(str(short))
    empty = tmp_path / "empty.csv"
    empty.write_text("")
    r = subprocess.run([SELFTEST, "params", str(empty)], capture_output=True, text=True)
    assert r.returncode == 1 and "Empty parameter file" in r.stderr


def test_lookup_mapper(tmp_path):
    lk = tmp_path / "lookup.csv"
    lk.write_text("stream,lat_index,lon_index\n420555774,28,39\n420552129,27,40\n")
    out = subprocess.run([SELFTEST, "lookup", str(lk), "420555774", "420552129", "5"], capture_output=True, text=True,
                         check=True).stdout.split("\n")
    assert out[0] == "2" and out[1] == "420555774 1 28 39" and out[2] == "420552129 1 27 40"
    assert out[3] == "5 0 -1 -1"  # {-1,-1} when absent, forcing_loader.cpp:54-60
    assert hostio.load_lookup(str(lk)) == {420555774: (28, 39), 420552129: (27, 40)}


def test_parameter_csv_writer_round_trips_through_the_loader(tmp_path):
    """hostio.write_spatial_params_csv (used to hand synthetic networks to the C++ programs) inverts the loader's
    unit conversions (I_O/parameters_loader.cpp:57-101) to rounding: topology exactly, parameters to 1e-15."""
    import numpy as np
    from tiger_hlm_gpu_b200 import synthetic
    from tiger_hlm_gpu_b200.hostio import load_spatial_params, write_spatial_params_csv
    sp = synthetic.apply_network(synthetic.make_spatial_params(50), synthetic.make_network(50, subbasin_links=10))
    path = str(tmp_path / "p.csv")
    write_spatial_params_csv(path, sp)
    back = load_spatial_params(path)
    assert np.array_equal(back["stream"], sp["stream"]) and np.array_equal(back["next_stream"], sp["next_stream"])
    for name in ("infil", "perco", "Hu", "n_mann", "slope", "L", "A_h", "alpha3", "alpha4", "melt_f", "temp_thr", "lat"):
        np.testing.assert_allclose(back[name], sp[name], rtol=1e-15, atol=0)
