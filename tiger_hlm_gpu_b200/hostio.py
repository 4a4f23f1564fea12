"""Host-side input contracts in Python (thin; the C++ host in tiger_hlm_gpu_b200/host mirrors them).

load_spatial_params follows I_O/parameters_loader.cpp:8-107: header-driven column lookup, required
columns, c1 = 0.001/60, infil = i2*c1, perco = i3*c1, alpha3 = res_ss*24*60, alpha4 = res_gw*24*60;
raises on a missing column or short row like the reference throws.
load_lookup follows I_O/forcing_loader.cpp:17-48 (stream,lat_index,lon_index; header skipped).
"""
from __future__ import annotations

import numpy as np

from .api import SPATIAL_PARAMS_DTYPE

REQUIRED = ["stream", "next_stream", "i2", "i3", "hu", "centroid_lat", "sw", "ss", "n", "slope", "length_km",
            "drainage_area_km2", "melt", "t_thres", "res_ss", "res_gw"]


def load_spatial_params(csv_path: str) -> np.ndarray:
    try:
        f = open(csv_path)
    except OSError:
        raise RuntimeError("Failed to open parameter file: " + csv_path)
    with f:
        header = f.readline()
        if not header:
            raise RuntimeError("Empty parameter file: " + csv_path)
        names = header.rstrip("\n").rstrip("\r").split(",")
        idx = {n: i for i, n in enumerate(names)}
        for n in REQUIRED:
            if n not in idx:
                raise RuntimeError(f"Missing column '{n}' in {csv_path}")
        rows = []
        for line in f:
            line = line.rstrip("\n").rstrip("\r")
            if not line:
                continue
            fields = line.split(",")
            if len(fields) < len(names):
                raise RuntimeError("Bad row with too few fields in " + csv_path)
            rows.append(fields)
    c1 = 0.001 / 60.0
    sp = np.zeros(len(rows), SPATIAL_PARAMS_DTYPE)
    col = lambda n, t=float: np.array([t(r[idx[n]]) for r in rows])  # noqa: E731
    sp["stream"] = col("stream", int)
    sp["next_stream"] = col("next_stream", int)
    sp["c1"] = c1
    sp["Hu"] = col("hu")
    sp["lat"] = col("centroid_lat")
    sp["sw"] = col("sw")
    sp["ss"] = col("ss")
    sp["n_mann"] = col("n")
    sp["slope"] = col("slope")
    sp["L"] = col("length_km")
    sp["A_h"] = col("drainage_area_km2")
    sp["melt_f"] = col("melt")
    sp["temp_thr"] = col("t_thres")
    sp["infil"] = col("i2") * c1
    sp["perco"] = col("i3") * c1
    sp["alpha3"] = col("res_ss") * 24.0 * 60.0
    sp["alpha4"] = col("res_gw") * 24.0 * 60.0
    return sp


def load_lookup(csv_path: str) -> dict[int, tuple[int, int]]:
    out: dict[int, tuple[int, int]] = {}
    with open(csv_path) as f:
        f.readline()
        for line in f:
            p = line.strip().split(",")
            if len(p) >= 3:
                out[int(p[0])] = (int(p[1]), int(p[2]))
    return out


def write_spatial_params_csv(path: str, sp: np.ndarray, lon=None) -> None:
    """A parameter CSV with the columns I_O/parameters_loader.cpp:35-49 requires, from SpatialParams records
    (inverse of the loader's unit conversions: i2 = infil / c1, res_ss = alpha3 / 1440, ...)."""
    c1 = 0.001 / 60.0
    cols = ["stream", "next_stream", "drainage_area_km2", "length_km", "area_sqkm", "centroid_lon", "centroid_lat", "hu",
            "i2", "i3", "sw", "ss", "n", "slope", "res_ss", "res_gw", "melt", "t_thres"]
    with open(path, "w") as f:
        f.write(",".join(cols) + "\n")
        for k, r in enumerate(sp):
            vals = [int(r["stream"]), int(r["next_stream"]), repr(float(r["A_h"])), repr(float(r["L"])), repr(float(r["A_h"])),
                    repr(float(lon[k]) if lon is not None else -75.0), repr(float(r["lat"])), repr(float(r["Hu"])),
                    repr(float(r["infil"] / c1)), repr(float(r["perco"] / c1)), repr(float(r["sw"])), repr(float(r["ss"])),
                    repr(float(r["n_mann"])), repr(float(r["slope"])), repr(float(r["alpha3"] / 1440.0)),
                    repr(float(r["alpha4"] / 1440.0)), repr(float(r["melt_f"])), repr(float(r["temp_thr"]))]
            f.write(",".join(str(v) for v in vals) + "\n")
