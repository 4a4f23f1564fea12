"""Model 200's power laws, x^(1/5) and x^(2/3) (oracle/devroot.h, csrc/fp_exact.cuh root5 / cbrt2).  Model 200 is
project-defined (the reference ships none), so these two routines are a definition, not a restatement: what is
checked is (1) their accuracy against long-double pow over every binade of the positive normal doubles and over the
range the model feeds them, and (2) on the GPU, that the device code and the C twin give the same bits."""
import numpy as np
import pytest

from oracle import oracle as O


def _samples(n=400_000, seed=7):
    rng = np.random.default_rng(seed)
    wide = np.exp(rng.uniform(np.log(2.3e-308), np.log(1.7e308), n))
    model = np.exp(rng.uniform(np.log(1e-6), np.log(1e6), n))  # discharge in m3/s, surface storage in m
    edge = np.array([2.2250738585072014e-308, 1e-6, 1.0, 32.0, 1.0 + 2.0 ** -52, 2.0 - 2.0 ** -52, 1.7976931348623157e308])
    return np.concatenate([wide, model, edge])


@pytest.mark.parametrize("which,num,den,bound", [(5, 1, 5, 3.0), (6, 2, 3, 2.0)])
def test_roots_within_a_few_ulp_of_long_double_pow(which, num, den, bound):
    x = _samples()
    got = O.eval_root(which, x)
    ref = np.power(x.astype(np.longdouble), np.longdouble(num) / np.longdouble(den))
    ulp = np.spacing(ref.astype(np.float64)).astype(np.longdouble)
    err = np.abs((got.astype(np.longdouble) - ref) / ulp)
    assert float(err.max()) < bound, float(err.max())
    assert np.all(np.diff(O.eval_root(which, np.sort(x[:50_000]))) >= 0)  # monotone on the sample


def test_roots_of_perfect_powers_and_binade_edges():
    """x = n^5 and x = n^3 (exactly representable): the roots come back within 3 ulp of n and n^2; and the routines
    cross binade edges of the argument without a jump (neighbouring doubles give results at most 2 ulp apart)."""
    n = np.arange(1, 2000, dtype=np.float64)
    r5 = O.eval_root(5, n ** 5)
    r3 = O.eval_root(6, n ** 3)
    assert np.max(np.abs(r5 - n) / np.spacing(n)) <= 3
    assert np.max(np.abs(r3 - n * n) / np.spacing(n * n)) <= 3
    for e in range(-40, 41, 3):
        edge = 2.0 ** e
        x = np.array([np.nextafter(edge, 0.0), edge, np.nextafter(edge, np.inf)])
        for which in (5, 6):
            r = O.eval_root(which, x)
            assert np.all(np.abs(np.diff(r)) <= 2 * np.spacing(r[1])), (which, e)


@pytest.mark.gpu
@pytest.mark.parametrize("which", [5, 6])
def test_device_roots_equal_the_c_twin_bit_for_bit(solver, which):
    x = _samples(seed=11)
    assert np.array_equal(solver.debug_eval(which, x), O.eval_root(which, x))
