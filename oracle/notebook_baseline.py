"""The reference's CPU path — "the model_dummy_python notebook integrator" — as a script.

TEST/BENCH INFRASTRUCTURE ONLY (like everything under oracle/): bench.py times it as a reported baseline and a
CPU test checks it against the C oracle; nothing under tiger_hlm_gpu_b200/ imports it.

The notebook (src/model_dummy_python.ipynb) integrates one system at a time with
scipy.integrate.solve_ivp(rhs, (t0, tf), y0, method='RK45', rtol=1e-6, atol=1e-9, t_eval=...) in a Python
loop (ipynb:150-175, 935-945); its Model204 right-hand side is the cell at ipynb:860-925 with stub parameters
and forcings.  Here the same right-hand side takes a link's SpatialParams record and looks the forcing sample
up by time (rain hourly, temperature daily), and the loop over links is spread over the host cores with
multiprocessing, one process per core (SURVEY §8(d) "Reference CPU baseline").  matplotlib/jupyter are not
needed.  Accepted steps are the intervals of the dense-output solution solve_ivp returns."""
from __future__ import annotations

import multiprocessing as mp
import os
import time

import numpy as np
from scipy.integrate import solve_ivp


def model204_rhs(t, y, p, pr_col, t2m_col):
    """ipynb:860-925, with the stubs replaced by the link's parameters and the forcing samples at time t."""
    h_snow, h_stat, h_surf, h_grav, h_aq = y
    rainfall = float(pr_col[min(int(t / 60.0), len(pr_col) - 1)]) if t >= 0 else float(pr_col[0])
    temperature = float(t2m_col[min(int(t / 1440.0), len(t2m_col) - 1)]) if t >= 0 else float(t2m_col[0])
    snowmelt = (temperature >= p["temp_thr"]) * min(h_snow, temperature * p["melt_f"])
    x1 = rainfall + snowmelt
    dh_snow = rainfall - snowmelt
    x2 = max(0.0, x1 + h_stat - p["Hu"])
    d1 = x1 - x2
    Emax = min(0.1 * temperature, h_stat)
    dh_stat = d1 - (h_stat / p["Hu"]) * Emax
    x3 = min(x2, p["infil"])
    d2 = x2 - x3
    alfa2 = (1.0 / p["n_mann"]) * max(h_surf, 0.0) ** (2.0 / 3.0) * np.sqrt(p["slope"])
    w = min(1.0, alfa2 * p["L"] / p["A_h"] * 60.0)
    dh_surf = d2 - h_surf * w
    x4 = min(x3, p["perco"])
    d3 = x3 - x4
    dh_grav = d3 - ((h_grav / p["alpha3"]) if p["alpha3"] >= 1.0 else 0.0)
    dh_aq = x4 - ((h_aq / p["alpha4"]) if p["alpha4"] >= 1.0 else 0.0)
    return [dh_snow, dh_stat, dh_surf, dh_grav, dh_aq]


def _solve_links(job):
    sp, pr, t2m, col, y0, t0, tf, tq = job
    finals, steps = [], 0
    for s in range(len(sp)):
        p = {k: float(sp[s][k]) for k in sp.dtype.names[2:]}
        sol = solve_ivp(model204_rhs, (t0, tf), y0[s], method="RK45", rtol=1e-6, atol=1e-9, t_eval=tq, dense_output=True,
                        args=(p, pr[:, col[s]], t2m[:, col[s]]))
        steps += len(sol.sol.ts) - 1
        finals.append(sol.y[:, -1])
    return np.array(finals), steps


def run(sp, pr, t2m, col, y0, t0, tf, tq, processes=None):
    """Integrate every link; returns (final [ns][5] at the last query, accepted steps, wall seconds, processes)."""
    processes = processes or os.cpu_count() or 1
    ns = len(sp)
    processes = max(1, min(processes, ns))
    cuts = np.linspace(0, ns, processes + 1).astype(int)
    jobs = [(sp[a:b], pr, t2m, col[a:b], y0[a:b], t0, tf, tq) for a, b in zip(cuts[:-1], cuts[1:])]
    t_start = time.perf_counter()
    if processes == 1:
        parts = [_solve_links(jobs[0])]
    else:
        with mp.get_context("fork").Pool(processes) as pool:
            parts = pool.map(_solve_links, jobs)
    dt = time.perf_counter() - t_start
    return np.concatenate([p[0] for p in parts]), int(sum(p[1] for p in parts)), dt, processes
