"""Writes tests/golden/model200_routed.npz: a regression fixture for the project-defined Model 200 and the
routed scheme, produced by the CPU oracle with the restated libdevice pow (oracle/devpow.h), whose arithmetic is
plain IEEE operations and therefore reproducible on any host.  Model 200 has no counterpart in the reference
(README.md:95 names it only), so this pins the implementation against ITSELF over time, nothing more; the
independent checks are SciPy (tests/test_oracle_model200.py).

    python tests/golden/make_model200_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from tests import routed_ref  # noqa: E402
from tiger_hlm_gpu_b200 import routing, synthetic  # noqa: E402


def case():
    ns = 48
    sp = synthetic.apply_network(synthetic.make_spatial_params(ns), synthetic.make_network(ns, subbasin_links=12, seed=200))
    col, ncells = synthetic.make_cells(ns, links_per_cell=6)
    pr, t2m = synthetic.make_forcing_grid(ncells, 1)
    rng = np.random.default_rng(200)
    y0 = np.tile(synthetic.Y0_200, (ns, 1))
    y0[:, 0] = rng.uniform(0.05, 5.0, ns)
    y0[::3, 2] = 0.005                         # ponded water on a third of the links: the surface store's pow() runs
    return sp, col, pr, t2m, y0


if __name__ == "__main__":
    sp, col, pr, t2m, y0 = case()
    prm = O.Params.make(initialStep=1e-6)
    F = O.Forcing([pr, t2m], [1.0, 24.0], col=col)
    tq = 60.0 * np.arange(1, 13)
    un = O.run_rk45(200, prm, y0, 0.0, 720.0, tq, sp=sp, forcing=F, device_pow=True, max_attempts=1_000_000)
    p1 = routing.plan(sp["stream"], sp["next_stream"], 1, subbasin_links=12)
    fin, dense, tqr, na = routed_ref.run_single(sp, F, y0, prm, p1, 0.0, 360.0, 20.0, threads=1)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "model200_routed.npz"),
                        unrouted_final=un["final"], unrouted_dense=un["dense"], unrouted_n_accept=un["n_accept"],
                        unrouted_n_reject=un["n_reject"], unrouted_stiff=un["stiff"],
                        routed_final=fin, routed_dense=dense, routed_tq=tqr, routed_n_accept=na)
    print("written", un["n_accept"].sum(), na.sum())
