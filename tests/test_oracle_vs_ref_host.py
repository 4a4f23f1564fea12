"""The oracle against the reference's own rk45_step / rk45_dense / Model204::rhs compiled for the
host (oracle/_ref/libref_host.so).  g++ chooses its own FMA contraction for the reference sources,
so agreement is to rounding level (1e-13 relative), not bit for bit — the bit-for-bit pin of the
oracle is the reference's committed GPU output (test_oracle_golden.py)."""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle as O
from tests import refs

pytestmark = pytest.mark.skipif(not refs.have("libref_host.so"), reason="oracle/_ref/libref_host.so not built")


def test_sizeof_spatial_params_is_136():
    assert refs.ref_host().ref_host_sizeof_spatial_params() == 136 == O.SPATIAL_DTYPE.itemsize


def test_step_and_dense_agree_with_reference_functions(small_test_params):
    sp = small_test_params
    L = refs.ref_host()
    rng = np.random.default_rng(0)
    for _ in range(500):
        y = np.array([rng.uniform(0, 0.02), rng.uniform(0, 5), rng.choice([0.0, rng.uniform(0, 0.05)]),
                      rng.uniform(0, 6), rng.uniform(0, 1)])
        h = 10 ** rng.uniform(-6, 1.2)
        F = np.array([rng.choice([0.0, rng.exponential(2) * 0.001 / 60]), rng.normal(5, 8)], np.float32)
        s = int(rng.integers(0, 10))
        yo, err, k = np.zeros(5), C.c_double(), np.zeros((7, 5))
        L.ref_host_step204(sp.ctypes.data, s, y.ctypes.data, h, 1e-6, 1e-9, F.ctypes.data, 2, yo.ctypes.data,
                           C.addressof(err), k.ctypes.data)
        yo2, err2, k2 = O.step(204, sp, s, y, h, 1e-6, 1e-9, float(F[0]), float(F[1]))
        np.testing.assert_allclose(yo2, yo, rtol=1e-12, atol=1e-15 * h * np.abs(k).max())
        np.testing.assert_allclose(k2, k, rtol=1e-11, atol=1e-18)
        # err is the ratio of a cancelling 7-term sum to tol: allow rounding noise of that sum
        tol_i = 1e-9 + 1e-6 * np.maximum(np.abs(y), np.abs(yo))
        noise = np.max(50 * 2.2e-16 * h * np.abs(k).max(axis=0) / tol_i)
        assert abs(err2 - err.value) <= 1e-9 * err.value + noise
        th = rng.uniform(0, 1)
        d = np.zeros(5)
        L.ref_host_dense204(y.ctypes.data, k.ctypes.data, h, th, d.ctypes.data)
        np.testing.assert_allclose(O.dense_eval(y, k, h, th), d, rtol=1e-12, atol=1e-15 * h * np.abs(k).max())


def test_full_run_agrees_with_reference_functions(small_test_params):
    sp = small_test_params
    ns = len(sp)
    rng = np.random.default_rng(1)
    pr = (rng.exponential(2.0, (48, ns)) * (rng.random((48, ns)) > 0.5) * 0.001 / 60).astype(np.float32)
    t2m = rng.normal(8, 4, (2, ns)).astype(np.float32)
    y0 = np.tile([0.01, 3.0, 0.002, 5.0, 0.2], (ns, 1))
    tq = np.arange(0, 2881, 60.0)
    prm = [1e-6, 1e-6, 1e-9, 0.9, 0.2, 10.0]
    a = refs.ref_host_run204(prm, y0, 0.0, 2880.0, tq, sp, [pr, t2m], [1.0, 24.0])
    b = O.run_rk45(204, O.Params.make(initialStep=1e-6), y0, 0.0, 2880.0, tq, sp=sp,
                   forcing=O.Forcing([pr, t2m], [1.0, 24.0]))
    assert np.array_equal(a["stiff"], b["stiff"])
    tol = 10 * (1e-9 + 1e-6 * np.abs(a["final"]))
    assert np.all(np.abs(a["final"] - b["final"]) <= tol)
    assert np.all(np.abs(a["dense"] - b["dense"]) <= 10 * (1e-9 + 1e-6 * np.abs(a["dense"])))
    # different contraction => trajectories differ in the last bits; counts must still be close
    assert np.all(np.abs(a["n_accept"] - b["n_accept"]) <= 2)
