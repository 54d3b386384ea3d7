"""
Prior sampling and the guessers that start fits from a prior: the draws must
come from the caller's RandomState in the reference's order, so a pipeline
seeded the same way starts its fits from the same points.  Golden vectors:
oracle/gen_golden.py --only guessers (reference: ngmix/guessers.py,
ngmix/joint_prior.py, ngmix/priors/*.py).  Host logic only: no GPU needed.
"""
import numpy as np
import pytest

from conftest import load_golden
from guesser_cases import guesser_cases


@pytest.fixture(scope="module")
def cases():
    import ngmix_b200
    return guesser_cases(ngmix_b200, lambda seed: np.random.RandomState(seed))


@pytest.mark.parametrize("name", [
    "sample_simple_n5", "sample_simple_one", "sample_bdf_n5",
    "sample_bdf_one", "sample_bd_n5", "sample_bd_one", "erf_sample",
    "ba_sample2d", "ba_sample2d_again", "tflux_and_prior", "tflux_fix",
    "tflux_and_prior_fix", "bdf_guess", "bd_guess", "prior_guess",
    "pars_guess_fix"])
def test_draws_equal_reference(cases, name):
    g = load_golden("guessers")
    got, want = np.asarray(cases[name]), g[name]
    assert got.shape == want.shape
    # same random stream, same arithmetic: equal to rounding of cos/sin/erf
    assert np.allclose(got, want, rtol=1e-13, atol=1e-15), (got, want)


def test_fixed_guesses_satisfy_prior(cases):
    """the fix-up cases start outside the tight prior (flux 100 > 80,
    T*1.1 > 0.3) and must end inside it"""
    for name in ("tflux_fix", "tflux_and_prior_fix", "pars_guess_fix"):
        g = np.asarray(cases[name])
        assert np.all((g[:, 4] >= 0.0) & (g[:, 4] <= 0.3)), name
        assert np.all((g[:, 5] >= 0.0) & (g[:, 5] <= 80.0)), name


def test_ba_maxval_is_the_peak():
    from ngmix_b200.priors import GPriorBA
    p = GPriorBA(0.3, rng=np.random.RandomState(1))
    p._set_maxval1d()
    gg = np.linspace(1e-3, 0.999, 20001)
    assert p.maxval1d >= p.get_prob_array1d(gg).max()
    assert abs(p.maxval1d / p.get_prob_array1d(gg).max() - 1) < 1e-6
