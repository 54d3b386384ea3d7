"""
GPU tests of the reference-facing Python API (ngmix_b200 package): the same
names, arguments and error behaviour as ngmix, with known answers from the
reference (generated from /root/reference) and the
logic of the reference's own tests (tests/test_gmix.py, test_pixels.py,
test_em.py, test_admom.py, test_runners.py determinism).
"""
import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ng():
    import ngmix_b200
    from ngmix_b200 import _lib
    _lib.init(0)
    return ngmix_b200


def rel(a, b):
    return np.max(np.abs(np.asarray(a) - np.asarray(b))
                  / np.maximum(np.abs(np.asarray(b)), 1e-300))


def _appendix_b(ng):
    jac = ng.DiagonalJacobian(row=23.5, col=23.5, scale=0.263)
    gm0 = ng.GMixModel([0.1, -0.2, 0.1, 0.05, 0.4, 100.0], 'exp')
    psf = ng.GMixModel([0, 0, 0.01, -0.01, 0.27, 1.0], 'turb')
    return jac, gm0, psf, gm0.convolve(psf)


def test_gmix_records_appendix_b(ng):
    jac, gm0, psf, gm = _appendix_b(ng)
    assert len(gm) == 18
    d = gm.get_data()
    want0 = [0.036745752151942465, 0.1, -0.2, 0.07704576145631171,
             -0.0015152294140495146, 0.08036861911599331]
    want17 = [3.455951941645767e-06, 0.1, -0.2, 1.2757041332519736,
              0.02376425561057645, 1.484451841587274]
    names = ("p", "row", "col", "irr", "irc", "icc")
    assert rel([d[n][0] for n in names], want0) < 1e-14
    assert rel([d[n][17] for n in names], want17) < 1e-14
    # pure host helpers
    assert abs(gm0.get_flux() - 100.0) < 1e-12
    assert abs(gm0.get_T() - 0.4) < 1e-6 * 0.4 + 1e-3   # exp table sums ~1
    row, col = gm.get_cen()
    assert abs(row - 0.1) < 1e-12 and abs(col + 0.2) < 1e-12


def test_make_image_and_loglike_appendix_b(ng):
    jac, gm0, psf, gm = _appendix_b(ng)
    fast = gm.make_image((48, 48), jacobian=jac, fast_exp=True)
    assert abs(fast.sum() / 99.99852896955666 - 1) < 1e-13
    assert abs(fast.max() / 4.278568261192419 - 1) < 1e-13
    assert fast[0, 0] == 0.0
    exact = gm.make_image((48, 48), jacobian=jac)
    assert abs(exact.sum() / 100.00000000013918 - 1) < 1e-13
    assert abs(exact[0, 0] / 4.983193526679106e-20 - 1) < 1e-10
    obs = ng.Observation(np.zeros((48, 48)), weight=np.zeros((48, 48)) + 4.0,
                         jacobian=jac)
    ll = gm.get_loglike(obs, more=True)
    assert abs(ll["loglike"] / -378.46427897109817 - 1) < 1e-13
    assert ll["s2n_numer"] == 0.0 and ll["npix"] == 2304
    assert abs(gm.get_model_s2n_sum(obs) / 756.9285579421963 - 1) < 1e-13
    fdiff = np.zeros(2304 + 13)
    gm.fill_fdiff(obs, fdiff, start=13)
    assert abs(fdiff.sum() / 199.99705793911332 - 1) < 1e-12
    assert abs(fdiff[13 + 1174] / 7.907652379929074 - 1) < 1e-12
    assert np.all(fdiff[:13] == 0)


@pytest.mark.parametrize("start", [0, 13])
def test_fdiff_closed_form(ng, start):
    """tests/test_gmix.py::test_gmix_loglike_fdiff: fill_fdiff against the
    closed-form gaussian inside 2 sigma, rtol 4e-5 (fexp accuracy)"""
    rng = np.random.RandomState(7)
    sigma, flux, noise = 3.0, 50.0, 0.1
    dims = (41, 41)
    jac = ng.UnitJacobian(row=20.0, col=20.0)
    T = 2 * sigma ** 2
    gm = ng.GMixModel([0.0, 0.0, 0.0, 0.0, T, flux], "gauss")
    image = np.zeros(dims)
    obs = ng.Observation(image, weight=np.zeros(dims) + 1 / noise ** 2,
                         jacobian=jac)
    fdiff = np.zeros(image.size + start)
    gm.fill_fdiff(obs, fdiff, start=start)
    rows, cols = np.mgrid[0:41, 0:41]
    r2 = (rows - 20.0) ** 2 + (cols - 20.0) ** 2
    model = flux / (2 * np.pi * sigma ** 2) * np.exp(-0.5 * r2 / sigma ** 2)
    want = (model - image) / noise
    got = fdiff[start:].reshape(dims)
    w = r2 < (2 * sigma) ** 2
    assert np.allclose(got[w], want[w], rtol=4e-5, atol=0)


def test_pixels_exact(ng):
    """tests/test_pixels.py: coords equal jac(row, col) exactly; zero-weight
    pixels are dropped unless ignore_zero_weight=False"""
    rng = np.random.RandomState(3)
    dims = (13, 17)
    jac = ng.Jacobian(row=6.1, col=8.3, dvdrow=0.25, dvdcol=0.01,
                      dudrow=-0.02, dudcol=0.27)
    coords = ng.make_coords(dims, jac)
    rows, cols = np.mgrid[0:dims[0], 0:dims[1]]
    v, u = jac(rows.ravel(), cols.ravel())
    assert np.array_equal(coords["v"], v) and np.array_equal(coords["u"], u)
    assert np.all(coords["area"] == jac.scale ** 2)
    image = rng.normal(size=dims)
    weight = rng.uniform(1, 2, size=dims)
    weight[3, 4] = 0.0
    weight[5, 6] = -1.0
    pix = ng.make_pixels(image, weight, jac)
    assert pix.size == image.size - 2
    keep = weight.ravel() > 0
    assert np.array_equal(pix["val"], image.ravel()[keep])
    assert np.array_equal(pix["ierr"], np.sqrt(weight.ravel()[keep]))
    assert np.array_equal(pix["v"], v[keep])
    pix_all = ng.make_pixels(image, weight, jac, ignore_zero_weight=False)
    assert pix_all.size == image.size
    assert pix_all["ierr"][3 * 17 + 4] == 0.0
    assert pix_all["ierr"][5 * 17 + 6] == 0.0
    with pytest.raises(ng.GMixFatalError):
        ng.make_pixels(image, weight * 0, jac)
    obs = ng.Observation(image, weight=weight, jacobian=jac)
    assert obs.pixels.size == image.size - 2
    assert not obs.pixels.flags["WRITEABLE"]


def test_gmix_errors(ng):
    with pytest.raises(ng.GMixRangeError):
        ng.GMixModel([0, 0, 0.8, 0.7, 1.0, 1.0], "exp")
    gm = ng.GMixModel([0, 0, 0.1, 0.1, -1.0, 1.0], "gauss")
    with pytest.raises(ng.GMixRangeError):
        gm.set_norms()
    with pytest.raises(ValueError):
        ng.GMix(pars=[1, 2, 3])
    with pytest.raises(ValueError):
        ng.GMixModel([0, 0, 0, 0, 1.0], "exp")


def test_fitter_appendix_b(ng):
    """LM known answer: flags 0, ier 2, nfev 6 and the fitted parameters"""
    jac, gm0, psf, gm = _appendix_b(ng)
    image = gm.make_image((48, 48), jacobian=jac)
    psf_obs = ng.Observation(psf.make_image((25, 25)), gmix=psf)
    obs = ng.Observation(image, weight=np.zeros((48, 48)) + 1e4,
                         jacobian=jac, psf=psf_obs)
    res = ng.Fitter(model='exp').go(obs, [0, 0, 0, 0, 0.5, 90.0])
    assert res["flags"] == 0 and res["ier"] == 2 and res["nfev"] == 6
    want = [0.10000004842391022, -0.19999994194872406, 0.09999990736127166,
            0.049999853131482246, 0.40000170401267765, 100.00011113574566]
    assert rel(res["pars"], want) < 1e-5
    werr = [7.808496001747722e-08, 8.718474622916069e-08,
            2.4026740153524596e-07, 2.397642276634209e-07,
            2.175658449742147e-07, 1.859937509674429e-05]
    assert rel(res["pars_err"], werr) < 1e-3
    assert abs(res["s2n"] / 1375.6166879333296 - 1) < 1e-6
    assert res["npix"] == 2304 and res["dof"] == 2298
    assert res["model"] == "exp"
    for k in ("g", "g_cov", "g_err", "T", "T_err", "flux", "flux_err",
              "chi2per", "lnprob", "pars_cov0", "pars_cov", "errmsg"):
        assert k in res
    im = res.make_image()
    assert np.abs(im - image).max() < 1e-4
    # forward differences (lmdif) agree with the analytic fit
    res2 = ng.Fitter(model='exp', analytic_jacobian=False).go(
        obs, [0, 0, 0, 0, 0.5, 90.0])
    assert res2["flags"] == 0 and res2["nfev"] == 36
    assert rel(res2["pars"], res["pars"]) < 1e-6


def test_fitter_with_prior_and_determinism(ng):
    g = load_golden("lm_exp")
    pr = ng.priors
    prior = pr.PriorSimpleSep(
        pr.CenPrior(0.0, 0.0, 0.263, 0.263), pr.GPriorBA(0.3),
        pr.FlatPrior(-1.0, 1.0e3), pr.FlatPrior(-100.0, 1.0e9))
    fitter = ng.Fitter(model="exp", prior=prior)
    for i in range(3):
        rec = g["lo_jac"][i]
        jac = ng.Jacobian(row=rec[0], col=rec[1], dvdrow=rec[2],
                          dvdcol=rec[3], dudrow=rec[4], dudcol=rec[5])
        psf_obs = ng.Observation(np.ones((5, 5)),
                                 gmix=ng.GMix(pars=g["lo_psf"][i].ravel()))
        obs = ng.Observation(g["lo_images"][i], weight=g["lo_weights"][i],
                             jacobian=jac, psf=psf_obs)
        r1 = fitter.go(obs, g["lo_guess"][i])
        r2 = fitter.go(obs, g["lo_guess"][i])
        assert r1["flags"] == g["lo_flat_flags"][i]
        assert r1["nfev"] == g["lo_flat_nfev"][i]
        err = g["lo_flat_pars_err"][i]
        assert np.all(np.abs(r1["pars"] - g["lo_flat_pars"][i])
                      <= np.maximum(1e-5 * np.abs(g["lo_flat_pars"][i]),
                                    1e-3 * err))
        assert abs(r1["lnprob"] / g["lo_flat_lnprob"][i] - 1) < 1e-6
        # run-to-run bit determinism (tests/test_runners.py:155-158)
        assert np.array_equal(r1["pars"], r2["pars"])
        assert np.array_equal(r1["pars_cov"], r2["pars_cov"])


def test_fitter_bd_with_joint_prior(ng):
    """Fitter(model='bd', prior=PriorBDSep(...)) against the reference"""
    g = load_golden("lm_bd")
    pr = ng.priors
    prior = ng.joint_prior.PriorBDSep(
        pr.CenPrior(0.0, 0.0, 0.263, 0.263), pr.GPriorBA(0.3),
        pr.FlatPrior(-1.0, 1.0e3), pr.Normal(0.0, 0.5), pr.Normal(0.5, 0.1),
        pr.FlatPrior(-100.0, 1.0e9))
    fitter = ng.Fitter(model="bd", prior=prior)
    for i in range(2):
        rec = g["jac"][i]
        jac = ng.Jacobian(row=rec[0], col=rec[1], dvdrow=rec[2],
                          dvdcol=rec[3], dudrow=rec[4], dudcol=rec[5])
        psf_obs = ng.Observation(np.ones((5, 5)),
                                 gmix=ng.GMix(pars=g["psf"][i].ravel()))
        obs = ng.Observation(g["images"][i], weight=g["weights"][i],
                             jacobian=jac, psf=psf_obs)
        r = fitter.go(obs, g["guess"][i])
        assert r["flags"] == g["bd_flags"][i]
        assert r["nfev"] == g["bd_nfev"][i]
        err = g["bd_pars_err"][i]
        assert np.all(np.abs(r["pars"] - g["bd_pars"][i])
                      <= np.maximum(1e-5 * np.abs(g["bd_pars"][i]),
                                    1e-2 * err))
        assert abs(r["lnprob"] / g["bd_lnprob"][i] - 1) < 1e-6
        # the prior object's own rows equal what the kernel used
        fd = np.zeros(8)
        assert prior.fill_fdiff(r["pars"], fd) == 7


def test_fitmodel_calc_fdiff_and_jacobian(ng):
    """FitModel.calc_fdiff / calc_jacobian / calc_lnprob at the guess against
    the reference's vectors (lm_trace.npz: prior rows, pixel rows, J^T J,
    J^T f), the central-difference check of tests/test_fitting_lm_jacobian.py
    :44-68, and the zero Jacobian / -inf residuals on range errors (:108-125)"""
    g = load_golden("lm_exp")
    t = load_golden("lm_trace")
    pr = ng.priors
    priors = {
        "none": None,
        "flat": pr.PriorSimpleSep(
            pr.CenPrior(0.0, 0.0, 0.263, 0.263), pr.GPriorBA(0.3),
            pr.FlatPrior(-1.0, 1.0e3), pr.FlatPrior(-100.0, 1.0e9)),
        "erf": pr.PriorSimpleSep(
            pr.CenPrior(0.0, 0.0, 0.263, 0.263), pr.GPriorBA(0.3),
            pr.TwoSidedErf(-1.0, 0.1, 1.0e3, 1.0),
            pr.TwoSidedErf(-1.0e2, 1.0, 1.0e9, 1.0e8)),
    }
    for pk, prior in priors.items():
        for i in range(2):
            rec = g["lo_jac"][i]
            jac = ng.Jacobian(row=rec[0], col=rec[1], dvdrow=rec[2],
                              dvdcol=rec[3], dudrow=rec[4], dudcol=rec[5])
            psf_obs = ng.Observation(
                np.ones((5, 5)), gmix=ng.GMix(pars=g["lo_psf"][i].ravel()))
            obs = ng.Observation(g["lo_images"][i], weight=g["lo_weights"][i],
                                 jacobian=jac, psf=psf_obs)
            guess = g["lo_guess"][i]
            fm = ng.fitting.FitModel(obs, "exp", guess, prior=prior)
            want = t["%s_%d_fdiff" % (pk, i)]
            assert fm.fdiff_size == want.size
            fd = fm.calc_fdiff(guess)
            assert np.abs(fd - want).max() <= 1e-11 * np.abs(want).max()
            J = fm.calc_jacobian(guess)
            assert J.shape == (want.size, 6)
            jp = t["%s_%d_jac_prior" % (pk, i)]
            assert np.allclose(J[:jp.shape[0]], jp, rtol=1e-12, atol=1e-14)
            jtj = t["%s_%d_jtj" % (pk, i)]
            assert np.abs(J.T @ J - jtj).max() <= 1e-10 * np.abs(jtj).max()
            assert np.abs(J.T @ fd - t["%s_%d_jtf" % (pk, i)]).max() \
                <= 1e-9 * np.abs(t["%s_%d_jtf" % (pk, i)]).max()
            # analytic Jacobian vs central differences of calc_fdiff
            for k, h in enumerate([1e-4, 1e-4, 1e-5, 1e-5, 1e-4, 1e-2]):
                p1, p2 = guess.copy(), guess.copy()
                p1[k] += h
                p2[k] -= h
                num = (fm.calc_fdiff(p1) - fm.calc_fdiff(p2)) / (2 * h)
                scale = np.abs(J[:, k]).max()
                assert np.abs(num - J[:, k]).max() <= 1e-4 * scale, (pk, k)
            # ln(prob) = -chi2/2 of the pixel rows + ln(prior)
            lnp = fm.calc_lnprob(guess)
            npri = 0 if prior is None else prior.fill_fdiff(guess,
                                                            np.zeros(8))
            lnprior = 0.0 if prior is None else \
                prior.get_lnprob_scalar(guess)
            assert abs(lnp - (-0.5 * (fd[npri:] ** 2).sum() + lnprior)) \
                < 1e-9 * abs(lnp)
            # range errors: -inf residuals, zero Jacobian, LOWVAL ln(prob)
            bad = guess.copy()
            bad[2] = 1.5
            assert np.all(np.isneginf(fm.calc_fdiff(bad)))
            assert not fm.calc_jacobian(bad).any()
            assert fm.calc_lnprob(bad) == -np.inf
            zero = guess.copy()
            zero[5] = 0.0
            assert not fm.calc_jacobian(zero).any()


def test_fit_batch_lognormal_and_truncated_priors(ng):
    """LogNormal / TruncatedGaussian priors through the Python prior classes
    (to_ngx) against the reference's fits"""
    g = load_golden("lm_priors2")
    pr = ng.priors
    cen, gp = pr.CenPrior(0.0, 0.0, 0.263, 0.263), pr.GPriorBA(0.3)
    sets = {
        "lognormal": pr.PriorSimpleSep(
            cen, gp, pr.LogNormal(0.5, 0.3),
            pr.LogNormal(100.0, 50.0, shift=-5.0)),
        "trunc": pr.PriorSimpleSep(
            cen, gp, pr.TruncatedGaussian(0.5, 0.4, 0.0, 5.0),
            pr.LogNormal(100.0, 50.0)),
    }
    for pk, prior in sets.items():
        recs = ng.fit_batch("exp", g["exp_images"], g["exp_weights"],
                            g["exp_jac"], g["exp_psf"], g["exp_guess"],
                            prior=prior)
        pre = "exp_%s_" % pk
        assert np.array_equal(recs["flags"], g[pre + "flags"])
        assert np.array_equal(recs["nfev"], g[pre + "nfev"])
        assert np.array_equal(recs["ier"], g[pre + "ier"])
        err = g[pre + "pars_err"]
        assert np.all(np.abs(recs["pars"][:, :6] - g[pre + "pars"])
                      <= 1e-3 * err)
        assert np.allclose(recs["lnprob"], g[pre + "lnprob"], rtol=1e-6)
        # host-side rows equal the kernel's convention
        fd = np.zeros(6)
        assert prior.fill_fdiff(g[pre + "pars"][0], fd) == 5
    prior = ng.joint_prior.PriorBDFSep(
        cen, gp, pr.LogNormal(0.5, 0.3),
        pr.TruncatedGaussian(0.5, 0.1, 0.0, 1.0), pr.LogNormal(100.0, 50.0))
    recs = ng.fit_batch("bdf", g["bdf_images"], g["bdf_weights"],
                        g["bdf_jac"], g["bdf_psf"], g["bdf_guess"],
                        prior=prior)
    assert np.array_equal(recs["flags"], g["bdf_res_flags"])
    assert np.array_equal(recs["nfev"], g["bdf_res_nfev"])
    assert np.all(np.abs(recs["pars"][:, :7] - g["bdf_res_pars"])
                  <= 1e-2 * g["bdf_res_pars_err"])
    assert np.allclose(recs["lnprob"], g["bdf_res_lnprob"], rtol=1e-6)
    # draws of the new priors follow the reference's recipe
    rng = np.random.RandomState(3)
    t = pr.TruncatedGaussian(0.5, 0.4, 0.0, 5.0, rng=rng).sample(100)
    assert t.min() > 0.0 and t.max() < 5.0
    z = pr.LogNormal(2.0, 0.5, rng=np.random.RandomState(3)).sample(5)
    z0 = np.random.RandomState(3).normal(size=5)
    lv = np.log(1 + 0.25 / 4.0)
    assert np.allclose(z, np.exp(np.log(2.0) - 0.5 * lv + np.sqrt(lv) * z0))


def test_em_appendix_b(ng):
    jac = ng.DiagonalJacobian(row=12, col=12, scale=0.263)
    truth = ng.GMix(pars=[60, -0.3, 0.2, 0.12, 0.01, 0.15,
                          40, 0.5, -0.4, 0.08, -0.005, 0.07])
    image = truth.make_image((25, 25), jacobian=jac)
    obs = ng.Observation(image, jacobian=jac)
    guess = ng.GMix(pars=[55, -0.25, 0.25, 0.10, 0, 0.13,
                          45, 0.45, -0.35, 0.09, 0, 0.08])
    res = ng.run_em(obs, guess)
    assert res["flags"] == 0 and res["numiter"] == 59
    assert abs(res["fdiff"] / 9.858073971976432e-06 - 1) < 1e-6
    assert abs(res["sky"] / 0.005290326870325679 - 1) < 1e-12
    want = [59.99682803412306, -0.3000231749588638, 0.20001847710298928,
            0.11998667675691517, 0.010007868605085228, 0.14998861694832763]
    assert rel(res.get_gmix().get_full_pars()[:6], want) < 1e-9
    assert res["message"] == "OK"
    r = ng.run_em(obs, guess, fixcen=True)
    assert r["numiter"] == 40
    r = ng.run_em(obs, guess, fluxonly=True)
    assert r["numiter"] == 20 and r["fdiff"] == 0.0
    # maxiter flag (tests/test_em.py:328-332)
    r = ng.run_em(obs, guess, maxiter=45, tol=1e-12)
    assert r["flags"] == ng.flags.EM_MAXITER and r["numiter"] == 45
    # a sky of zero on an image with zero pixels -> EM_RANGE_ERROR
    im0 = image.copy()
    im0[:, :] = 0.0
    r = ng.run_em(ng.Observation(im0, jacobian=jac), guess, sky=0.0)
    assert r["flags"] == ng.flags.EM_RANGE_ERROR
    assert not r.has_gmix()
    # the 3-gaussian turb fit of Appendix B runs to maxiter
    psf = ng.GMixModel([0, 0, 0.01, -0.01, 0.27, 1.0], 'turb')
    pim = psf.make_image((25, 25), jacobian=jac)
    g3 = ng.GMix(pars=[0.5, 0, 0, 0.06, 0, 0.06, 0.4, 0, 0, 0.15, 0, 0.15,
                       0.1, 0, 0, 0.5, 0, 0.5])
    r = ng.em.EMFitter(tol=1e-5).go(ng.Observation(pim, jacobian=jac), g3)
    assert r["flags"] == 32 and r["numiter"] == 500
    assert abs(r["sky"] / 0.00010428745730873525 - 1) < 1e-12


def test_admom_appendix_b(ng):
    jac, gm0, psf, gm = _appendix_b(ng)
    image = gm.make_image((48, 48), jacobian=jac)
    obs = ng.Observation(image, weight=np.zeros((48, 48)) + 1e4, jacobian=jac)
    guess = ng.GMixModel([0, 0, 0, 0, 0.5, 1], 'gauss')
    res = ng.run_admom(obs, guess)
    assert res["flags"] == 0 and res["numiter"] == 9 and res["npix"] == 2304
    want = [0.09982728486221708, -0.19960107899791002, 0.057239031493414166,
            0.021157418384838412, 0.513338931509584, 2.0176633395849266]
    assert rel(res["pars"], want) < 1e-9
    assert rel(res["e"], [0.11150339080086998, 0.04121530062530121]) < 1e-9
    assert abs(res["rho4"] / 2.1246812211272452 - 1) < 1e-9
    assert abs(res["wsum"] / 0.9999860084748363 - 1) < 1e-11
    assert np.isfinite(res["flux"]) and np.isfinite(res["T_err"])
    gmfit = res.get_gmix()
    assert abs(gmfit.get_T() / res["T"] - 1) < 1e-12
    # T guess + rng, and the centre finder
    r2 = ng.run_admom(obs, 0.5, rng=np.random.RandomState(1))
    assert r2["flags"] == 0
    # converged from a different start: equal within the etol/Ttol scale
    assert rel(r2["pars"][:5], want[:5]) < 5e-3
    r3 = ng.admom.find_cen_admom(obs, fwhm=1.2)
    assert r3["flags"] == 0
    assert abs(r3["cen"][0] - 0.1) < 0.01 and abs(r3["cen"][1] + 0.2) < 0.01


def test_admom_golden_results(ng):
    """get_result post-processing against the reference's dict values"""
    g = load_golden("em_admom")
    am = ng.AdmomFitter()
    for i in range(8):
        rec = g["jac"][i]
        jac = ng.Jacobian(row=rec[0], col=rec[1], dvdrow=rec[2],
                          dvdcol=rec[3], dudrow=rec[4], dudcol=rec[5])
        obs = ng.Observation(g["images"][i], weight=g["weights"][i],
                             jacobian=jac)
        r = am.go(obs, ng.GMix(pars=g["am_guess"][i]))
        assert r["flags"] == g["am_flags_final"][i]
        assert r["numiter"] == g["am_numiter"][i]
        for k in ("flux", "flux_err", "T", "T_err", "s2n"):
            assert abs(r[k] / g["am_" + k][i] - 1) < 1e-8, k
        assert rel(r["e"], g["am_e"][i]) < 1e-7
        assert rel(r["e_err"], g["am_e_err"][i]) < 1e-7


def test_batch_api_matches_single(ng):
    g = load_golden("lm_exp")
    recs = ng.fit_batch("exp", g["hi_images"], g["hi_weights"], g["hi_jac"],
                        g["hi_psf"], g["hi_guess"])
    assert np.array_equal(recs["flags"], g["hi_none_flags"])
    assert np.array_equal(recs["nfev"], g["hi_none_nfev"])
    res, pars = ng.em_batch(load_golden("em_admom")["images"][:4],
                            load_golden("em_admom")["weights"][:4],
                            load_golden("em_admom")["jac"][:4],
                            load_golden("em_admom")["em_guess"][:4])
    assert np.array_equal(res["numiter"],
                          load_golden("em_admom")["em_numiter"][:4])


def test_gaussap_fluxes(ng):
    """gaussian-aperture fluxes against the reference (ngmix/gaussap.py,
    GMix.get_gaussap_flux): values, range-error and not-attempted flags"""
    g = load_golden("gaussap")
    for model in ("exp", "dev", "bdf", "gauss"):
        flux, flags = ng.gaussap.get_gaussap_flux(
            g[model + "_pars"], model, 2.5, mask=g[model + "_mask"])
        assert np.array_equal(flags, g[model + "_flags"])
        assert np.allclose(flux, g[model + "_flux"], rtol=1e-12,
                           equal_nan=True)
    flux, flags = ng.gaussap.get_gaussap_flux(
        g["cm_pars"], "cm", 2.0, fracdev=g["cm_fracdev"],
        TdByTe=g["cm_TdByTe"])
    assert np.array_equal(flags, g["cm_flags"])
    assert np.allclose(flux, g["cm_flux"], rtol=1e-12)
