"""Bit-for-bit differential test against the UNCHANGED reference kernel, built for sm_100a
(oracle/_ref/libref_cuda.so: solver/rk45_kernel.cu + models/model_204_global.cu +
I_O/forcing_data.cu compiled from /root/reference with the reference Makefile's flags).

This is the pin for everything the CPU oracle cannot promise: libdevice pow() on the Model204
surface-storage branch and nvcc/ptxas' own FMA contraction of the reference sources.  The new
kernel must reproduce the reference kernel's final states, dense output, stiff flags and (via the
counting wrappers described in oracle/ref_cuda_harness.cu) attempt counts exactly.
"""
import numpy as np
import pytest

from tests import refs
from tiger_hlm_gpu_b200 import Parameters, synthetic

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not refs.have("libref_cuda.so"), reason="oracle/_ref/libref_cuda.so not built")]

PRM6 = [1e-6, 1e-6, 1e-9, 0.9, 0.2, 10.0]


def build_case(ns, days, wet_fraction, seed=204):
    sp = synthetic.make_spatial_params(ns, seed)
    col, ncells = synthetic.make_cells(ns, links_per_cell=61)
    pr, t2m = synthetic.make_forcing_grid(ncells, days)
    y0 = synthetic.make_y0(ns, wet_fraction)
    tf = days * 1440.0
    tq = synthetic.hourly_queries(0.0, tf)
    return sp, col, pr, t2m, y0, tf, tq


def run_new(solver, sp, col, pr, t2m, y0, tf, tq):
    solver.set_model_parameters(204, Parameters(*PRM6))
    solver.set_max_attempts(2_000_000)
    solver.upload_spatial_params(sp)
    solver.clear_forcings()
    solver.upload_forcing(0, 1.0, pr)
    solver.upload_forcing(1, 24.0, t2m)
    solver.set_forcing_columns(col)
    return solver.run_rk45(204, y0, 0.0, tf, tq)


def test_counting_wrappers_do_not_change_the_reference_kernel():
    sp, col, pr, t2m, y0, tf, tq = build_case(300, 1, 0.5)
    blocks = [synthetic.expand_forcing_per_link(pr, col), synthetic.expand_forcing_per_link(t2m, col)]
    a = refs.ref_cuda_run204(PRM6, y0, 0.0, tf, tq, sp, blocks, [1.0, 24.0], counted=False)
    b = refs.ref_cuda_run204(PRM6, y0, 0.0, tf, tq, sp, blocks, [1.0, 24.0], counted=True)
    assert np.array_equal(a["final"], b["final"]) and np.array_equal(a["dense"], b["dense"])
    assert np.array_equal(a["stiff"], b["stiff"])


@pytest.mark.parametrize("ns,days,wet", [(500, 2, 0.0), (500, 2, 1.0), (1500, 3, 0.4)])
def test_new_kernel_equals_reference_kernel_bit_for_bit(solver, ns, days, wet):
    sp, col, pr, t2m, y0, tf, tq = build_case(ns, days, wet)
    blocks = [synthetic.expand_forcing_per_link(pr, col), synthetic.expand_forcing_per_link(t2m, col)]
    ref = refs.ref_cuda_run204(PRM6, y0, 0.0, tf, tq, sp, blocks, [1.0, 24.0], counted=True)
    new = run_new(solver, sp, col, pr, t2m, y0, tf, tq)
    assert ref["n_accept"].min() > 100 * days
    for key in ("n_accept", "n_reject", "n_jump"):
        assert np.array_equal(new[key], ref[key]), key
    assert np.array_equal(new["stiff"], ref["stiff"])
    assert np.array_equal(new["final"], ref["final"])
    assert np.array_equal(new["dense"], ref["dense"])


def test_reference_kernel_on_its_own_golden(small_test_params, golden204):
    """The rebuilt reference kernel reproduces the reference's committed output (float forcing 0.001f
    instead of the golden's double literal: solver-tolerance agreement, see test_gpu_parity)."""
    sp = small_test_params
    ns = len(sp)
    y0 = np.tile(synthetic.Y0_204, (ns, 1))
    blocks = [np.full((48, ns), 0.001, np.float32), np.full((2, ns), 1.0, np.float32)]
    r = refs.ref_cuda_run204(PRM6, y0, 0.0, 2880.0, golden204["query_times"], sp, blocks, [1.0, 24.0])
    assert np.all(np.abs(r["final"] - golden204["final"]) <= 10 * (1e-9 + 1e-6 * np.abs(golden204["final"])))
    assert np.all(r["n_accept"] > 800)
