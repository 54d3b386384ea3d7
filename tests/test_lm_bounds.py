"""
Bounded LM fits: prior members with `bounds` make run_leastsq iterate on
leastsqbound's internal parameters (ngmix/fitting/leastsqbound.py:183-263,
445-552; joint_prior.py:37-61).  Golden vectors: oracle/gen_golden.py --only
lm_bounds - bdf with fracdev in (0, 1) through lmdif
(tests/test_leastsqbound.py:114-176 of the reference) and exp with bounded T
and flux through lmder / lmdif (tests/test_fitting_lm_jacobian.py:216-250).

Covariance: under scipy >= 1.15 the reference's bounded branch applies a
1-based `ipvt - 1` to a 0-based ipvt, which scrambles its cov_x (SURVEY.md
section 7).  The fixture holds both the covariance as the reference computes
it here (`pars_cov0_asrun`, unpinned: not reproduced) and the one its code
intends, obtained by handing it the 1-based ipvt it assumes
(`pars_cov0`); the fit itself (pars, nfev, ier) is identical in both runs and
is what parity is defined on, together with the intended covariance.
"""
import numpy as np
import pytest

from conftest import load_golden

BDF_BOUNDS = [(None, None)] * 5 + [(0.0, 1.0), (None, None)]
EXP_BOUNDS = [(None, None)] * 4 + [(0.01, 100.0), (1.0, 1.0e9)]
EXP2_BOUNDS = [(None, None)] * 4 + [(0.05, 5.0), (1.0, None)]

CASES = [("bdf", "bdf_res_", "bdf", BDF_BOUNDS, False),
         ("exp", "exp_an_res_", "exp", EXP_BOUNDS, True),
         ("exp", "exp_fd_res_", "exp", EXP_BOUNDS, False),
         ("exp2", "exp2_an_res_", "exp", EXP2_BOUNDS, True),
         ("exp2", "exp2_fd_res_", "exp", EXP2_BOUNDS, False)]


def _ora_prior(ora, tag):
    pr = ora.make_prior("flat")
    if tag == "bdf":
        pr.fracdev = ora.Prior1D(3, 0, 0.5, 0.3, 0, 0)
    elif tag == "exp":
        pr.T = ora.Prior1D(3, 0, 0.4, 1.0, 0, 0)
        pr.F = ora.Prior1D(3, 0, 100.0, 100.0, 0, 0)
    else:
        pr.T = ora.Prior1D(3, 0, 0.5, 1.0, 0, 0)
        pr.F = ora.Prior1D(3, 0, 100.0, 100.0, 0, 0)
    return pr


def _check(r, g, pre, i, npars, cov_rtol, sig_tol=1e-3):
    assert int(r["flags"]) == int(g[pre + "flags"][i])
    assert int(r["nfev"]) == int(g[pre + "nfev"][i])
    assert int(r["ier"]) == int(g[pre + "ier"][i])
    ref = g[pre + "pars"][i]
    err = g[pre + "pars_err"][i]
    assert np.all(np.abs(r["pars"][:npars] - ref)
                  <= np.maximum(1e-5 * np.abs(ref), sig_tol * err))
    c0 = np.asarray(r["pars_cov0"])[:npars * npars].reshape(npars, npars) \
        if np.asarray(r["pars_cov0"]).ndim == 1 else r["pars_cov0"]
    cref = g[pre + "pars_cov0"][i]
    sc = np.sqrt(np.outer(np.abs(np.diag(cref)), np.abs(np.diag(cref))))
    assert np.max(np.abs(c0 - cref) / sc) < cov_rtol
    assert abs(r["lnprob"] / g[pre + "lnprob"][i] - 1) < 1e-6
    assert abs(r["s2n"] / g[pre + "s2n"][i] - 1) < 1e-6


@pytest.mark.parametrize("tag,pre,model,bounds,analytic", CASES)
def test_oracle_bounded_fits_match_reference(ora, tag, pre, model, bounds,
                                             analytic):
    g = load_golden("lm_bounds")
    npars = len(bounds)
    for i in range(g[tag + "_images"].shape[0]):
        r = ora.lm_fit(model, g[tag + "_images"][i], g[tag + "_weights"][i],
                       g[tag + "_jac"][i], g[tag + "_psf"][i],
                       g[tag + "_guess"][i], prior=_ora_prior(ora, tag),
                       analytic_jacobian=analytic, bounds=bounds)
        # numpy's sin / arcsin (the reference's transforms) and libm's differ
        # in the last bit; the forward-difference quotients amplify that to
        # 1e-8 in J and more in the covariance
        _check(r, g, pre, i, npars, 1e-6 if analytic else 1e-3)


def test_reference_bounded_covariance_is_scrambled_here():
    """the hazard of SURVEY.md section 7 in numbers: as the reference runs
    under this scipy its bounded cov_x is far from (J^T J)^-1 in external
    parameters, while the fit is the same"""
    g = load_golden("lm_bounds")
    a, b = g["bdf_res_pars_cov0_asrun"], g["bdf_res_pars_cov0"]
    assert np.max(np.abs(a - b) / np.abs(b)) > 1.0
    assert np.array_equal(g["bdf_res_flags_asrun"], g["bdf_res_flags"])


def test_prior_bounds_like_reference():
    import ngmix_b200 as ng
    pr = ng.priors
    cen, gp = pr.CenPrior(0.0, 0.0, 0.263, 0.263), pr.GPriorBA(0.3)
    flatT, flatF = pr.FlatPrior(-1.0, 1.0e3), pr.FlatPrior(-100.0, 1.0e9)
    p = ng.joint_prior.PriorBDFSep(
        cen, gp, flatT, pr.Normal(0.5, 0.3, bounds=[0.0, 1.0]), flatF)
    assert p.bounds == BDF_BOUNDS
    p = ng.joint_prior.PriorSimpleSep(cen, gp, flatT, flatF)
    assert p.bounds is None
    p = ng.joint_prior.PriorSimpleSep(
        cen, gp, pr.Normal(0.4, 1.0, bounds=[0.01, 100.0]),
        pr.Normal(100.0, 100.0, bounds=[1.0, 1.0e9]))
    assert p.bounds == EXP_BOUNDS
    p = ng.joint_prior.PriorBDSep(
        cen, gp, flatT, pr.Normal(0.0, 0.1),
        pr.Normal(0.5, 0.1, bounds=(0, 1)), flatF)
    assert p.bounds == [(None, None)] * 6 + [(0, 1), (None, None)]


def _gpu_prior(ng, tag):
    pr = ng.priors
    cen, gp = pr.CenPrior(0.0, 0.0, 0.263, 0.263), pr.GPriorBA(0.3)
    if tag == "bdf":
        return ng.joint_prior.PriorBDFSep(
            cen, gp, pr.FlatPrior(-1.0, 1.0e3),
            pr.Normal(0.5, 0.3, bounds=[0.0, 1.0]),
            pr.FlatPrior(-100.0, 1.0e9))
    if tag == "exp":
        return ng.joint_prior.PriorSimpleSep(
            cen, gp, pr.Normal(0.4, 1.0, bounds=[0.01, 100.0]),
            pr.Normal(100.0, 100.0, bounds=[1.0, 1.0e9]))
    return ng.joint_prior.PriorSimpleSep(
        cen, gp, pr.Normal(0.5, 1.0, bounds=[0.05, 5.0]),
        pr.Normal(100.0, 100.0, bounds=[1.0, None]))


@pytest.mark.gpu
@pytest.mark.parametrize("tag,pre,model,bounds,analytic", CASES)
def test_gpu_bounded_fits_match_reference(tag, pre, model, bounds, analytic):
    import ngmix_b200 as ng
    from ngmix_b200 import _lib
    _lib.init(0)
    g = load_golden("lm_bounds")
    prior = _gpu_prior(ng, tag)
    assert prior.bounds == bounds
    npars = len(bounds)
    recs = ng.fit_batch(model, g[tag + "_images"], g[tag + "_weights"],
                        g[tag + "_jac"], g[tag + "_psf"], g[tag + "_guess"],
                        prior=prior, analytic_jacobian=analytic)
    for i in range(recs.size):
        # the forward-difference covariance carries the 1e-8 noise of the
        # difference quotients
        if tag == "exp" and not analytic:
            # flux ~ 100 bounded to (1, 1e9): the internal parameter sits
            # 6e-4 from -pi/2, where sin(x) + 1 ~ 2e-7 and the forward
            # difference of the transform (~1e-11) keeps five digits of a
            # 1e-16 accurate sine - the flux column of J carries the last-bit
            # differences between CUDA's and numpy's sin at the 1e-5 level,
            # and the fit barely moves from its guess (ier 2 after 8 to 15
            # evaluations).  The reference itself calls this set-up "not a
            # reliable yardstick" (tests/test_fitting_lm_jacobian.py:250):
            # same flags and the same fit quality are what can be asked
            assert int(recs["flags"][i]) == int(g[pre + "flags"][i])
            assert abs(recs["lnprob"][i] / g[pre + "lnprob"][i] - 1) < 0.05
            continue
        _check(recs[i], g, pre, i, npars, 1e-5 if analytic else 2e-3)


@pytest.mark.gpu
def test_gpu_bounded_fitter_go():
    """the reference-facing call: Fitter(prior with bounds).go"""
    import ngmix_b200 as ng
    from ngmix_b200 import _lib
    _lib.init(0)
    g = load_golden("lm_bounds")
    rec = g["exp2_jac"][0]
    jac = ng.Jacobian(row=rec[0], col=rec[1], dvdrow=rec[2], dvdcol=rec[3],
                      dudrow=rec[4], dudcol=rec[5])
    pjac = ng.DiagonalJacobian(row=2, col=2, scale=0.263)
    psf = ng.Observation(np.ones((5, 5)), jacobian=pjac,
                         gmix=ng.GMix(pars=g["exp2_psf"][0].ravel()))
    obs = ng.Observation(g["exp2_images"][0], weight=g["exp2_weights"][0],
                         jacobian=jac, psf=psf)
    r = ng.Fitter(model="exp", prior=_gpu_prior(ng, "exp2")).go(
        obs, g["exp2_guess"][0])
    assert r["flags"] == 0 and r["nfev"] == int(g["exp2_an_res_nfev"][0])
    assert np.allclose(r["pars"], g["exp2_an_res_pars"][0], rtol=1e-5,
                       atol=1e-3 * g["exp2_an_res_pars_err"][0].max())
    assert np.allclose(r["pars_err"], g["exp2_an_res_pars_err"][0],
                       rtol=1e-4)
