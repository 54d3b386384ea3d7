"""
The reference's own recovery tests, re-hosted without galsim: the truth is
rendered by the engine (exact exp, ngmix-style sims as tests/_sims.py does)
instead of by galsim, everything else follows the reference's test logic:

  tests/test_ml_fitting_exp_obj_gauss_psf.py:13-135  sheared-WCS jacobians,
      guesses drawn from the joint prior, exp / bdf / bd fits with priors,
      <g> over 50 noise-free realisations within 1e-5 (2e-3 for bd), centre
      unbiased within 5 sigma
  tests/test_fitting_lm_jacobian.py:171-213  analytic vs forward-difference
      fits agree to 0.02 sigma, the analytic fit needs < 0.6x the nfev
  tests/test_runners.py:155-158  run-to-run bit determinism
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ng():
    import ngmix_b200
    from ngmix_b200 import _lib
    _lib.init(0)
    return ngmix_b200


def _get_prior(ng, fit_model, rng, scale):
    """tests/_priors.py:4-55"""
    pr = ng.priors
    g_prior = pr.GPriorBA(sigma=0.1, rng=rng)
    cen_prior = pr.CenPrior(cen1=0, cen2=0, sigma1=scale, sigma2=scale,
                            rng=rng)
    T_prior = pr.FlatPrior(minval=-1.0, maxval=1.0e3, rng=rng)
    F_prior = pr.FlatPrior(minval=-100.0, maxval=1.0e9, rng=rng)
    if fit_model == "bd":
        return ng.joint_prior.PriorBDSep(
            cen_prior=cen_prior, g_prior=g_prior, T_prior=T_prior,
            logTratio_prior=pr.Normal(mean=0.0, sigma=0.1, rng=rng),
            fracdev_prior=pr.Normal(mean=0.5, sigma=0.1, rng=rng),
            F_prior=F_prior)
    if fit_model == "bdf":
        return ng.joint_prior.PriorBDFSep(
            cen_prior=cen_prior, g_prior=g_prior, T_prior=T_prior,
            fracdev_prior=pr.Normal(mean=0.5, sigma=0.1, rng=rng),
            F_prior=F_prior)
    return ng.joint_prior.PriorSimpleSep(
        cen_prior=cen_prior, g_prior=g_prior, T_prior=T_prior,
        F_prior=F_prior)


def _shear_wcs(scale, g1, g2):
    """world = M image for a sheared pixel grid of the given scale
    (what galsim.ShearWCS(scale, Shear(g1, g2)).jacobian() holds)"""
    f = scale / np.sqrt(1.0 - g1 * g1 - g2 * g2)
    return dict(dudx=f * (1 - g1), dudy=-f * g2, dvdx=-f * g2,
                dvdy=f * (1 + g1))


@pytest.mark.parametrize("wcs_g", [(-0.5, -0.2), (0.0, 0.0)])
@pytest.mark.parametrize("g_true", [(-0.1, -0.2), (0.0, 0.0)])
@pytest.mark.parametrize("fit_model", ["exp", "bdf", "bd"])
def test_ml_fitting_exp_obj_gauss_psf(ng, wcs_g, g_true, fit_model):
    rng = np.random.RandomState(seed=10)
    image_size = 33
    cen = (image_size - 1) / 2
    wcs = _shear_wcs(0.25, *wcs_g)
    scale = np.sqrt(abs(wcs["dudx"] * wcs["dvdy"] - wcs["dudy"] * wcs["dvdx"]))
    g1_true, g2_true = g_true
    Tpsf = ng.moments.fwhm_to_T(0.5)
    T_true = 0.5
    psf_gmix = ng.gmix.make_gmix_model([0, 0, 0, 0, Tpsf, 1], "gauss")
    psf_obs = ng.Observation(image=np.ones((5, 5)), gmix=psf_gmix)

    def render(shift):
        # object centre at world position `shift` = jacobian centre shifted
        # by M^-1 shift in pixels
        M = np.array([[wcs["dudx"], wcs["dudy"]], [wcs["dvdx"], wcs["dvdy"]]])
        dxy = np.linalg.solve(M, shift)
        jac = ng.Jacobian(y=cen + dxy[1], x=cen + dxy[0], **wcs)
        gm = ng.GMixModel([0.0, 0.0, g1_true, g2_true, T_true, 400.0],
                          "exp").convolve(psf_gmix)
        return gm.make_image((image_size, image_size), jacobian=jac), jac

    im0, _ = render(np.zeros(2))
    noise = np.sqrt(np.sum(im0 ** 2)) / 1e16
    wgt = np.ones_like(im0) / noise ** 2

    prior = _get_prior(ng, fit_model, rng, scale)
    guess = prior.sample()
    guess[0:5] = [0, 0, g1_true, g2_true, Tpsf]
    if fit_model == "bd":
        guess[5:8] = [1.0, 0.5, 400]
    elif fit_model == "bdf":
        guess[6] = 400
    else:
        guess[5] = 400

    fitter = ng.Fitter(model=fit_model, prior=prior)
    g1arr, g2arr, xarr, yarr = [], [], [], []
    for _ in range(50):
        shift = rng.uniform(low=-scale / 2, high=scale / 2, size=2)
        im, jac = render(shift)
        _im = im + rng.normal(size=im.shape) * noise
        obs = ng.Observation(image=_im, weight=wgt, jacobian=jac, psf=psf_obs)
        res = fitter.go(obs=obs,
                        guess=guess + rng.normal(size=guess.size) * 0.01)
        if res["flags"] == 0:
            g1arr.append(res["g"][0])
            g2arr.append(res["g"][1])
            xarr.append(res["pars"][1])
            yarr.append(res["pars"][0])
    assert len(g1arr) >= 45
    gtol = 0.002 if fit_model == "bd" else 1.0e-5
    assert abs(np.mean(g1arr) - g1_true) < gtol
    assert abs(np.mean(g2arr) - g2_true) < gtol
    # the truth sits at the jacobian centre: fitted centres scatter around 0
    for arr in (xarr, yarr):
        err = np.std(arr) / np.sqrt(len(arr))
        assert abs(np.mean(arr)) < max(5 * err, 1e-6)


@pytest.mark.parametrize("model", ["gauss", "exp", "dev"])
def test_analytic_vs_forward_difference_fits(ng, model):
    """test_fitting_lm_jacobian.py:171-213"""
    from ngmix_b200 import sims
    b = sims.make_lm_batch(24, seed0=4242, noise=0.05, model=model,
                           dims=(40, 40))
    pr = ng.priors
    prior = pr.PriorSimpleSep(
        pr.CenPrior(0.0, 0.0, 0.263, 0.263), pr.GPriorBA(0.3),
        pr.FlatPrior(-1.0, 1.0e3), pr.FlatPrior(-100.0, 1.0e9))
    ra = ng.fit_batch(model, b["images"], b["weights"], b["jac"], b["psf"],
                      b["guess"], prior=prior, analytic_jacobian=True)
    rf = ng.fit_batch(model, b["images"], b["weights"], b["jac"], b["psf"],
                      b["guess"], prior=prior, analytic_jacobian=False)
    ok = (ra["flags"] == 0) & (rf["flags"] == 0)
    assert ok.sum() >= 22
    dev = np.abs(ra["pars"][ok, :6] - rf["pars"][ok, :6]) \
        / ra["pars_err"][ok, :6]
    assert dev.max() < 0.02
    assert ra["nfev"][ok].sum() < 0.6 * rf["nfev"][ok].sum()
    # same call twice: bit-identical results
    ra2 = ng.fit_batch(model, b["images"], b["weights"], b["jac"], b["psf"],
                       b["guess"], prior=prior, analytic_jacobian=True)
    assert np.array_equal(ra["pars"], ra2["pars"])
    assert np.array_equal(ra["pars_cov"], ra2["pars_cov"])
