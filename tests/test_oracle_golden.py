"""The oracle against every golden vector the reference holds for this path."""
import numpy as np

from oracle import oracle as O

Y0_204 = [0.01, 3.0, 0.0, 5.0, 0.2]


def run_204_stub(sp):
    y0 = np.tile(Y0_204, (len(sp), 1))
    tq = np.arange(0, 2881, 1.0)
    return O.run_rk45(204, O.Params.make(initialStep=1e-6), y0, 0.0, 2880.0, tq, sp=sp,
                      forcing=O.Forcing(stub=[0.001, 1.0]))


def test_model204_final_and_dense_match_reference_netcdf_bit_for_bit(small_test_params, golden204):
    """src/final_example.nc / src/dense_example.nc were written by the reference's GPU build.
    The oracle reproduces all 10 x 2881 x 5 values EXACTLY, which pins its FMA-contraction model."""
    r = run_204_stub(small_test_params)
    assert np.array_equal(r["final"], golden204["final"])
    for s in range(10):
        assert np.array_equal(r["dense"][s], golden204["dense_sys0"])
    assert not r["stiff"].any()
    # row t=0 is never written (tq <= t0), SURVEY F10
    assert np.all(r["dense"][:, 0, :] == 0.0)
    # the survey's line-by-line Python transcription counted 901 accepted / 28 rejected
    assert np.all(r["n_accept"] == 901) and np.all(r["n_reject"] == 28) and np.all(r["n_jump"] == 0)


def test_model204_golden_printed_values(golden204):
    """The 9-digit values quoted in SURVEY §0.3 / BASELINE.md are what the fixture holds."""
    np.testing.assert_allclose(golden204["final"][0],
                               [1.00000033e-03, 3.45073806, 0.0, 1.83939721, 1.92857916e-01], rtol=5e-9)
    np.testing.assert_allclose(golden204["dense_sys0"][1],
                               [4.31091494e-03, 3.00600174, 0.0, 4.99826419, 1.99997475e-01], rtol=5e-9)


def test_dummy_matches_reference_csv(golden_dummy):
    """src/final.csv, src/dense.csv: DummyModel, y0 = ones, 6 significant digits."""
    tq = golden_dummy["query_times"]
    r = O.run_rk45(O.UID_DUMMY, O.Params.make(), np.ones((4, 5)), 0.0, 5.0, tq)
    assert np.all(r["final"] == r["final"][0])
    np.testing.assert_allclose(r["final"], golden_dummy["final_csv"], rtol=3e-6)
    # dense.csv times are printed with 6 digits too; compare at the exact query grid with that slack
    np.testing.assert_allclose(golden_dummy["time_csv"], tq, rtol=2e-5)
    np.testing.assert_allclose(r["dense"][0], golden_dummy["dense_csv_sys0"], rtol=2e-5, atol=1e-6)


def test_dummy_matches_scipy_like_the_notebook(golden_dummy):
    """model_dummy_python.ipynb compares the GPU result with scipy solve_ivp(RK45); tolerance-level."""
    tq = golden_dummy["query_times"]
    r = O.run_rk45(O.UID_DUMMY, O.Params.make(), np.ones((1, 5)), 0.0, 5.0, tq)
    tol = 10 * (1e-9 + 1e-6 * np.abs(golden_dummy["scipy_final"]))
    assert np.all(np.abs(r["final"][0] - golden_dummy["scipy_final"]) < tol)
    assert np.all(np.abs(r["dense"][0] - golden_dummy["scipy_dense"]) < 10 * (1e-9 + 1e-6 * np.abs(golden_dummy["scipy_dense"])))
    assert int(golden_dummy["scipy_len_t"]) == 12 and int(golden_dummy["scipy_nfev"]) == 68  # ipynb:351


def test_threads_give_identical_results(small_test_params):
    a = run_204_stub(small_test_params)
    y0 = np.tile(Y0_204, (10, 1))
    b = O.run_rk45(204, O.Params.make(initialStep=1e-6), y0, 0.0, 2880.0, np.arange(0, 2881, 1.0),
                   sp=small_test_params, forcing=O.Forcing(stub=[0.001, 1.0]), threads=4)
    for k in a:
        assert np.array_equal(a[k], b[k])
