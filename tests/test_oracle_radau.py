"""CPU pin of the implicit fallback (oracle/oracle_radau.inc): links the RK45 path flags stiff are carried
to tf by 3-stage Radau IIA.  The reference's own Radau code is unfinished, so the pin is
SciPy's solve_ivp(method="Radau") at much tighter tolerances; the bound is the solver's own tolerance
10 * (atol + rtol * |y|), written below."""
import numpy as np
from scipy.integrate import solve_ivp

from oracle import oracle as O
from tiger_hlm_gpu_b200 import synthetic

RTOL, ATOL = 1e-6, 1e-9


def stiff_case(ns=64, seed=5):
    """Constant heavy rain on nearly saturated static storage: the surface store fills and its fast drain
    makes the explicit path bail out (6 consecutive rejections / step below the floor)."""
    rng = np.random.default_rng(seed)
    sp = synthetic.make_spatial_params(ns)
    rain = (rng.uniform(0.5, 6.0, ns) * 0.001 / 60 * 8).astype(np.float32)  # m/min
    temp = rng.uniform(-3, 12, ns).astype(np.float32)
    pr, t2m = np.tile(rain, (48, 1)), np.tile(temp, (2, 1))
    y0 = np.tile(synthetic.Y0_204, (ns, 1))
    y0[:, 1] = rng.uniform(100, 177.9, ns)
    return sp, rain, temp, pr, t2m, y0


def rhs204(t, y, p, r, T):  # models/model_204.hpp:86-113 in plain Python
    h_snow, h_stat, h_surf, h_grav, h_aq = y
    melt = min(h_snow, T * p["melt_f"]) if T >= p["temp_thr"] else 0.0
    x1 = r + melt
    x2 = max(0.0, x1 + h_stat - p["Hu"])
    d1 = x1 - x2 - (h_stat / p["Hu"]) * min(0.1 * T, h_stat)
    x3 = min(x2, p["infil"])
    alfa2 = (1.0 / p["n_mann"]) * max(h_surf, 0.0) ** (2.0 / 3.0) * np.sqrt(p["slope"])
    w = min(1.0, alfa2 * p["L"] / p["A_h"] * 60.0)
    x4 = min(x3, p["perco"])
    return [r - melt, d1, x2 - x3 - h_surf * w, x3 - x4 - (h_grav / p["alpha3"] if p["alpha3"] >= 1 else 0.0),
            x4 - (h_aq / p["alpha4"] if p["alpha4"] >= 1 else 0.0)]


def test_fallback_solves_flagged_links_within_tolerance_of_scipy_radau():
    sp, rain, temp, pr, t2m, y0 = stiff_case()
    tq = 60.0 * np.arange(1, 25)
    P = O.Params.make(initialStep=1e-6, rtol=RTOL, atol=ATOL)
    F = O.Forcing([pr, t2m], [1.0, 24.0])
    plain = O.run_rk45(204, P, y0, 0.0, 1440.0, tq, sp=sp, forcing=F, threads=8, max_attempts=2_000_000)
    fb = O.run_rk45(204, P, y0, 0.0, 1440.0, tq, sp=sp, forcing=F, threads=8, max_attempts=2_000_000, stiff_fallback=True)
    flagged = plain["stiff"] == 1
    assert flagged.sum() >= 3
    assert np.array_equal(fb["stiff"], np.where(flagged, 3, 0))          # 3 = HLM_LINK_STIFF_SOLVED
    assert (fb["n_radau"][flagged] > 0).all() and not fb["n_radau"][~flagged].any()
    # links that never left the explicit path are untouched, bit for bit
    for k in ("final", "dense", "n_accept", "n_reject", "n_jump"):
        assert np.array_equal(fb[k][~flagged], plain[k][~flagged]), k
    # what the explicit path had already emitted for a flagged link stays as it was
    for s in np.where(flagged)[0]:
        written = plain["dense"][s].any(axis=1)
        assert np.array_equal(fb["dense"][s][written], plain["dense"][s][written])
    for s in np.where(flagged)[0]:
        sol = solve_ivp(rhs204, (0.0, 1440.0), y0[s], method="Radau", rtol=1e-9, atol=1e-12, t_eval=tq,
                        args=(sp[s], float(rain[s]), float(temp[s])))
        ref = sol.y.T
        bound = 10.0 * (ATOL + RTOL * np.abs(ref))
        assert (np.abs(fb["dense"][s] - ref) <= bound).all()
        assert (np.abs(fb["final"][s] - ref[-1]) <= bound[-1]).all()
        assert fb["n_radau"][s] < 2000  # implicit steps are few: the link is stiff, not fast


def test_fallback_off_keeps_the_reference_behaviour():
    sp, rain, temp, pr, t2m, y0 = stiff_case(ns=32, seed=6)
    tq = 60.0 * np.arange(1, 25)
    P = O.Params.make(initialStep=1e-6)
    r = O.run_rk45(204, P, y0, 0.0, 1440.0, tq, sp=sp, forcing=O.Forcing([pr, t2m], [1.0, 24.0]), max_attempts=2_000_000)
    flagged = r["stiff"] == 1
    assert flagged.any() and not r["final"][flagged].any()  # no final state written (rk45_kernel.cu:167-170)
    assert not r["n_radau"].any()
