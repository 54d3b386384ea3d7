"""
Coelliptical-gaussian LM fits (CoellipFitter, PriorCoellipSame,
CoellipPSFGuesser: ngmix/fitting/fitters.py:120-141, results.py:628-674,
joint_prior.py:874-1031, guessers.py:1106-1216).  Golden vectors:
oracle/gen_golden.py --only lm_coellip (reference fits of PSF stamps with 1-5
gaussians, with and without the prior, from the reference guesser's guesses).

The oracle follows the reference on all 60 fits (flags up to the 1024 bit,
nfev, ier, parameters).  The 3- to 5-gaussian fits of a 3-gaussian psf are
degenerate and chaotic in the last bit of the model values (see
test_degenerate_coellip_fits_are_chaotic), which sets what "parity" can mean
for any implementation that is not bit-identical to the reference.
"""
import numpy as np
import pytest

from conftest import load_golden

CASES = [(ng, pk) for ng in (1, 2, 3, 4, 5) for pk in ("none", "prior")]


def _check(r, g, pre, i, np_, par_tol):
    mask = ~np.int32(1024)
    assert int(r["flags"]) & mask == int(g[pre + "flags"][i]) & mask
    assert int(r["nfev"]) == int(g[pre + "nfev"][i])
    assert int(r["ier"]) == int(g[pre + "ier"][i])
    err = g[pre + "pars_err"][i]
    ref = g[pre + "pars"][i]
    if int(g[pre + "flags"][i]) == 0 and int(r["flags"]) == 0:
        tol = np.maximum(1e-5 * np.abs(ref), par_tol * err)
        assert np.all(np.abs(r["pars"][:np_] - ref) <= tol)
        assert abs(r["lnprob"] - g[pre + "lnprob"][i]) \
            <= 1e-6 * abs(g[pre + "lnprob"][i])
        assert int(r["dof"]) == int(g[pre + "dof"][i])
        assert abs(r["s2n"] / g[pre + "s2n"][i] - 1) < 1e-6


@pytest.mark.parametrize("ngauss,pk", CASES)
def test_oracle_coellip_matches_reference(ora, ngauss, pk):
    g = load_golden("lm_coellip")
    pre = "n%d_%s_" % (ngauss, pk)
    for i in range(g["images"].shape[0]):
        prior = ora.make_prior("erf" if pk == "prior" else "none")
        r = ora.lm_fit("coellip", g["images"][i], g["weights"][i],
                       g["jac"][i], None, g[pre + "guess"][i], prior=prior)
        _check(r, g, pre, i, 4 + 2 * ngauss, 1e-6)


def test_coellip_guesser_and_prior_draws():
    """CoellipPSFGuesser reproduces the guesses the golden fits started from
    (same RandomState stream as the reference)"""
    import ngmix_b200 as ng
    g = load_golden("lm_coellip")
    for ngauss, pk in CASES:
        rng = np.random.RandomState(100 * ngauss + (pk == "prior"))
        guesser = ng.guessers.CoellipPSFGuesser(rng=rng, ngauss=ngauss)
        pre = "n%d_%s_" % (ngauss, pk)
        for i in range(g["images"].shape[0]):
            rec = g["jac"][i]
            jac = ng.Jacobian(row=rec[0], col=rec[1], dvdrow=rec[2],
                              dvdcol=rec[3], dudrow=rec[4], dudcol=rec[5])
            obs = ng.Observation(g["images"][i], weight=g["weights"][i],
                                 jacobian=jac, store_pixels=False)
            assert np.allclose(guesser(obs=obs), g[pre + "guess"][i],
                               rtol=1e-14, atol=0)
    pr = ng.priors
    p = ng.joint_prior.PriorCoellipSame(
        2, pr.CenPrior(0.0, 0.0, 0.263, 0.263), pr.GPriorBA(0.3),
        pr.TwoSidedErf(-1.0, 0.1, 1.0e3, 1.0),
        pr.TwoSidedErf(-1.0e2, 1.0, 1.0e9, 1.0e8))
    fd = np.zeros(8)
    assert p.fill_fdiff(g["n2_prior_pars"][0], fd) == 7
    assert p.sample(3).shape == (3, 8)
    with pytest.raises(ValueError):
        p.get_lnprob_scalar(np.zeros(6))


@pytest.fixture(scope="module")
def ng():
    import ngmix_b200
    from ngmix_b200 import _lib
    _lib.init(0)
    return ngmix_b200


def _prior(ng, ngauss, pk):
    if pk == "none":
        return None
    pr = ng.priors
    return ng.joint_prior.PriorCoellipSame(
        ngauss, pr.CenPrior(0.0, 0.0, 0.263, 0.263), pr.GPriorBA(0.3),
        pr.TwoSidedErf(-1.0, 0.1, 1.0e3, 1.0),
        pr.TwoSidedErf(-1.0e2, 1.0, 1.0e9, 1.0e8))


def test_degenerate_coellip_fits_are_chaotic(ora, tmp_path):
    """
    Evidence for the parity definition of the degenerate fits below: the
    oracle itself - the same MINPACK restatement, the same Householder QR in
    the same order - stops reproducing the reference's evaluation counts on
    the 4- and 5-gaussian fits as soon as the compiler may contract a*b+c
    into fused multiply-adds (a 1e-16 change of the model values, amplified
    to 1e-8 by the forward-difference Jacobian's step of 1.5e-8 |x| and then
    by the flat valley).  No implementation whose arithmetic is not
    bit-identical to the reference's (an FMA-based GPU kernel cannot be) can
    follow these iterates; what it can match is where the fit ends up.
    The 1- and 2-gaussian fits are unaffected.
    """
    import ctypes as C
    import os
    import subprocess
    src = [os.path.join(os.path.dirname(ora.LIBPATH), f)
           for f in ("ora_gmix.c", "ora_minpack.c", "ora_lm.c",
                     "ora_em_admom.c")]
    so = str(tmp_path / "liboracle_fma.so")
    subprocess.check_call(["gcc", "-O2", "-fPIC", "-std=gnu11",
                           "-ffp-contract=fast", "-mfma", "-pthread",
                           "-shared", "-o", so] + src + ["-lm", "-lpthread"])
    fma = C.CDLL(so)
    g = load_golden("lm_coellip")

    def nfev_equal(ngauss, lib):
        pre = "n%d_none_" % ngauss
        same = 0
        for i in range(g["images"].shape[0]):
            prior = ora.make_prior("none")
            prior.ncoellip = ngauss
            res = np.zeros(1, dtype=ora.LM_RESULT_DTYPE)
            im, wt = g["images"][i], g["weights"][i]
            jac, guess = g["jac"][i], g[pre + "guess"][i]
            lib.ora_lm_fit(
                C.c_int(ora.MODELS["coellip"]), im.ctypes.data_as(C.c_void_p),
                wt.ctypes.data_as(C.c_void_p), C.c_int(im.shape[0]),
                C.c_int(im.shape[1]), jac.ctypes.data_as(C.c_void_p), None,
                C.c_int(0), guess.ctypes.data_as(C.c_void_p), C.byref(prior),
                C.c_int(4000), C.c_double(1e-5), C.c_double(1e-5), C.c_int(1),
                C.c_int(1), res.ctypes.data_as(C.c_void_p))
            same += int(res["nfev"][0] == g[pre + "nfev"][i])
        return same

    n = g["images"].shape[0]
    assert nfev_equal(2, ora.lib()) == n and nfev_equal(5, ora.lib()) == n
    assert nfev_equal(2, fma) == n
    assert nfev_equal(5, fma) < n // 2


@pytest.mark.gpu
@pytest.mark.parametrize("ngauss,pk", CASES)
def test_gpu_coellip_matches_reference(ng, ngauss, pk):
    """
    1 and 2 gaussians: flags, nfev, ier and parameters follow the reference,
    up to the forward-difference noise floor: the oracle itself, compiled with
    FMA contraction, differs from its unfused build in the evaluation count of
    0.2-0.4 % of such fits (one lmdif iteration more or less at the ftol /
    xtol boundary, the same optimum), and the CUDA kernels differ from the
    oracle at the same rate (profiles/coellip_parity_r2c.json, 2 x 2000 fits;
    test_gpu_coellip_evaluation_counts_at_the_noise_floor).  A golden fit may
    therefore end one iteration apart if it ends at the same point.
    3 to 5 gaussians on a 3-gaussian psf are degenerate (hundreds of
    evaluations along a flat valley) and chaotic in the last bit of the model
    values (test_degenerate_coellip_fits_are_chaotic): there parity means the
    same flags up to the covariance bits (its eigenvalues are at rounding
    level) and the same fit quality, ln(prob) within 5e-3 relative.
    """
    g = load_golden("lm_coellip")
    pre = "n%d_%s_" % (ngauss, pk)
    recs = ng.fit_batch_coellip(ngauss, g["images"], g["weights"], g["jac"],
                                g[pre + "guess"], prior=_prior(ng, ngauss, pk))
    apart = 0
    for i in range(recs.size):
        if ngauss <= 2:
            np_ = 4 + 2 * ngauss
            dn = abs(int(recs["nfev"][i]) - int(g[pre + "nfev"][i]))
            if dn == 0:
                _check(recs[i], g, pre, i, np_, 0.05)
                continue
            # one lmdif iteration = 1 + npars evaluations
            apart += 1
            assert dn <= np_ + 1
            assert int(recs["flags"][i]) == 0 == int(g[pre + "flags"][i])
            ref, err = g[pre + "pars"][i], g[pre + "pars_err"][i]
            assert np.all(np.abs(recs["pars"][i][:np_] - ref)
                          <= np.maximum(1e-5 * np.abs(ref), 0.05 * err))
            assert abs(recs["lnprob"][i] / g[pre + "lnprob"][i] - 1) < 1e-5
            continue
        mask = ~np.int32(512 | 1024 | 2048)
        assert int(recs["flags"][i]) & mask == \
            int(g[pre + "flags"][i]) & mask
        if int(g[pre + "flags"][i]) == 0 and int(recs["flags"][i]) == 0:
            assert abs(recs["lnprob"][i] / g[pre + "lnprob"][i] - 1) < 5e-3
            assert int(recs["dof"][i]) == int(g[pre + "dof"][i])
    assert apart <= 1


@pytest.mark.gpu
def test_gpu_coellip_evaluation_counts_at_the_noise_floor(ng, ora):
    """400 one- and two-gaussian fits of psf stamps against the oracle: the
    same flags everywhere, the same evaluation counts on >= 98.5 % (the
    oracle's own FMA noise floor is 99.6 - 99.8 %), >= 99 % of the fits within
    0.05 sigma, all within 0.2 sigma and 1e-5 in ln(prob)"""
    from ngmix_b200 import sims, _lib
    n = 400
    b = sims.make_psf_batch(n, seed0=515151)
    T = ng.moments.fwhm_to_T(0.263 * 3.5)
    flux = b["images"].sum(axis=(1, 2))
    for ngauss in (1, 2):
        np_ = 4 + 2 * ngauss
        rng = np.random.RandomState(11)
        guess = np.zeros((n, np_))
        guess[:, 0:2] = rng.uniform(-0.01, 0.01, (n, 2))
        guess[:, 2:4] = rng.uniform(-0.05, 0.05, (n, 2))
        fr = np.array(ng.guessers._COELLIP_F[ngauss])
        pr = np.array(ng.guessers._COELLIP_P[ngauss])
        guess[:, 4:4 + ngauss] = T * fr * rng.uniform(0.99, 1.01, (n, ngauss))
        guess[:, 4 + ngauss:] = flux[:, None] * pr \
            * rng.uniform(0.99, 1.01, (n, ngauss))
        r = _lib.lm_fit_coellip(ngauss, b["images"], b["weights"], b["jac"],
                                None, guess)
        same = close = 0
        for i in range(n):
            o = ora.lm_fit("coellip", b["images"][i], b["weights"][i],
                           b["jac"][i], None, guess[i],
                           prior=ora.make_prior("none"))
            assert int(r["flags"][i]) == int(o["flags"]) == 0
            same += int(r["nfev"][i] == o["nfev"])
            dev = np.abs(r["pars"][i][:np_] - o["pars"][:np_])
            floor = 1e-5 * np.abs(o["pars"][:np_])
            close += int(np.all(dev <= np.maximum(
                floor, 0.05 * o["pars_err"][:np_])))
            assert np.all(dev <= np.maximum(floor, 0.2 * o["pars_err"][:np_]))
            assert abs(r["lnprob"][i] / o["lnprob"] - 1) < 1e-5
        assert same >= int(0.985 * n), same
        assert close >= int(0.99 * n), close


@pytest.mark.gpu
def test_gpu_coellip_fitter_and_psf_runner(ng):
    """the reference-facing calls: CoellipFitter.go, and PSFRunner with a
    CoellipPSFGuesser setting obs.psf.gmix"""
    g = load_golden("lm_coellip")
    rec = g["jac"][0]
    jac = ng.Jacobian(row=rec[0], col=rec[1], dvdrow=rec[2], dvdcol=rec[3],
                      dudrow=rec[4], dudcol=rec[5])
    psf_obs = ng.Observation(g["images"][0], weight=g["weights"][0],
                             jacobian=jac)
    fitter = ng.CoellipFitter(ngauss=2)
    r = fitter.go(psf_obs, g["n2_none_guess"][0])
    assert r["flags"] == 0 and r["nfev"] == int(g["n2_none_nfev"][0])
    assert np.allclose(r["pars"], g["n2_none_pars"][0], rtol=1e-5,
                       atol=5e-2 * g["n2_none_pars_err"][0].max())
    assert r["pars_cov"].shape == (8, 8) and "flux" not in r
    assert r.get_gmix().get_data().size == 2
    obs = ng.Observation(g["images"][1], weight=g["weights"][1], jacobian=jac,
                         psf=psf_obs)
    runner = ng.runners.PSFRunner(
        fitter=fitter,
        guesser=ng.guessers.CoellipPSFGuesser(
            rng=np.random.RandomState(5), ngauss=2), ntry=2)
    runner.go(obs)
    assert obs.psf.meta["result"]["flags"] == 0
    assert obs.psf.has_gmix() and len(obs.psf.gmix) == 2
    with pytest.raises(NotImplementedError):
        ng.CoellipFitter(ngauss=6)
