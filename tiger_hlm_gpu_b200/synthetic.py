"""Seeded synthetic inputs for tests and bench.py (SURVEY §8(d) "Synthetic inputs").

No dataset of the reference's scale is available offline (its 41 274-link parameter CSV and the
ERA5-Land forcing files are absent), so inputs of the same shape are generated:
  parameters  the constants of data/small_test.csv (hu=178, i2=4, i3=1.6, sw=.11, ss=.33, n=.1,
              slope=.02, res_ss=2, res_gw=55, melt=3.7, t_thres=0) with length_km ~ U(0.09, 2.1)
              and drainage_area_km2 ~ logU(0.13, 1.6), converted exactly as
              I_O/parameters_loader.cpp:57-101 does (infil = i2*c1, alpha3 = res_ss*1440, ...);
  forcing     a grid with ~537 links per cell (the median of data/small_example_pr_lookup.csv),
              links sorted by cell; pr hourly in m/min = (0.001/60)*X with X = 0 w.p. 0.85 else
              Exp(mean 2); t2m daily in deg C = 10 + 8 sin(2 pi d/365) + N(0, 2);
  y0          {0.01, 3, 0, 5, 0.2} for every link (main.cpp:376); `wet_fraction` > 0 gives that
              share of links a surface storage of 0.01 m so the pow() branch of Model204 runs.
"""
from __future__ import annotations

import numpy as np

from .api import SPATIAL_PARAMS_DTYPE

C1 = 0.001 / 60.0  # I_O/parameters_loader.cpp:57
Y0_204 = (0.01, 3.0, 0.0, 5.0, 0.2)  # main.cpp:376
LINKS_PER_CELL = 537


def make_spatial_params(ns: int, seed: int = 204) -> np.ndarray:
    rng = np.random.default_rng(seed)
    sp = np.zeros(ns, SPATIAL_PARAMS_DTYPE)
    sp["stream"] = 420000000 + np.arange(ns)
    sp["next_stream"] = 420000000 + (np.arange(ns) // 2)  # carried, read by no equation
    sp["c1"] = C1
    sp["infil"] = 4.0 * C1
    sp["perco"] = 1.6 * C1
    sp["Hu"] = 178.0
    sp["lat"] = 40.3
    sp["sw"] = 0.11
    sp["ss"] = 0.33
    sp["n_mann"] = 0.1
    sp["slope"] = 0.02
    sp["L"] = rng.uniform(0.09, 2.1, ns)
    sp["A_h"] = np.exp(rng.uniform(np.log(0.13), np.log(1.6), ns))
    sp["alpha3"] = 2.0 * 24.0 * 60.0
    sp["alpha4"] = 55.0 * 24.0 * 60.0
    sp["melt_f"] = 3.7
    sp["temp_thr"] = 0.0
    return sp


def make_cells(ns: int, links_per_cell: int = LINKS_PER_CELL) -> tuple[np.ndarray, int]:
    """Per-link forcing column (links sorted by cell) and the number of cells."""
    ncells = max(1, (ns + links_per_cell - 1) // links_per_cell)
    col = (np.arange(ns) // links_per_cell).astype(np.int32)
    return col, ncells


def make_forcing_grid(ncells: int, days: int, seed: int = 2019) -> tuple[np.ndarray, np.ndarray]:
    """(pr [24*days][ncells] f32 m/min, t2m [days][ncells] f32 deg C)."""
    rng = np.random.default_rng(seed)
    nT = 24 * days
    wet = rng.random((nT, ncells)) >= 0.85
    x = rng.exponential(2.0, (nT, ncells))
    pr = (C1 * np.where(wet, x, 0.0)).astype(np.float32)
    d = np.arange(days)[:, None]
    t2m = (10.0 + 8.0 * np.sin(2.0 * np.pi * d / 365.0) + rng.normal(0.0, 2.0, (days, ncells))).astype(np.float32)
    return pr, t2m


def make_y0(ns: int, wet_fraction: float = 0.0, seed: int = 7) -> np.ndarray:
    y0 = np.tile(np.array(Y0_204), (ns, 1))
    if wet_fraction > 0:
        rng = np.random.default_rng(seed)
        wet = rng.random(ns) < wet_fraction
        y0[wet, 2] = 0.01
    return y0


def hourly_queries(t0: float, tf: float) -> np.ndarray:
    """main.cpp:653-657: for (t = t0; t <= tf; t += 60)."""
    return np.arange(t0, tf + 0.5, 60.0)


def expand_forcing_per_link(grid: np.ndarray, col: np.ndarray) -> np.ndarray:
    """The reference's per-link expansion [time][system] (main.cpp:543-548) of a grid forcing."""
    return np.ascontiguousarray(grid[:, col])


Y0_200 = (0.5, 3.0, 0.0, 5.0, 0.2)  # channel discharge 0.5 m3/s + Model204's hillslope stores (main.cpp:376)


def make_network(ns: int, subbasin_links: int = 4096, seed: int = 200) -> np.ndarray:
    """Downstream link index per link (-1 = outlet) of a synthetic river network ("parent in a
    random binary-ish tree per sub-basin of 4096 links").

    Links come in consecutive groups of `subbasin_links`; inside a group link j > 0 drains into the link
    before it (main stem, half of the time) or into a random earlier link at most 64 back (a tributary
    joining), and the group's first link drains into a random link of one of the 8 groups before it, so the
    groups themselves form a tree whose edges are what a partition by sub-basin cuts."""
    rng = np.random.default_rng(seed)
    j = np.arange(ns, dtype=np.int64) % subbasin_links
    g0 = np.arange(ns, dtype=np.int64) - j                     # first link of the group
    back = np.where(rng.random(ns) < 0.5, 1, rng.integers(1, 65, ns))
    down = g0 + np.maximum(j - back, 0)
    heads = np.flatnonzero(j == 0)
    for h in heads:                                            # one entry per group: cheap
        lo = max(0, h - 8 * subbasin_links)
        down[h] = -1 if h == 0 else rng.integers(lo, h)
    return down


def apply_network(sp: np.ndarray, down: np.ndarray) -> np.ndarray:
    """Write the topology into the records the way the reference's CSV carries it (stream, next_stream)."""
    sp = sp.copy()
    sp["next_stream"] = np.where(down >= 0, sp["stream"][np.maximum(down, 0)], 0)
    return sp
