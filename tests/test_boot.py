"""
The production caller: template psf fluxes,
guessers that draw from the caller's RandomState in the reference's order,
PSFRunner / Runner / Bootstrapper.  Golden vectors come from the reference
run with the same seeds (oracle/gen_golden.py: psfflux, boot).
"""
import os
import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def test_oracle_template_flux(ora):
    g = _load("psfflux")
    for normalize in (True, False):
        tag = "n%d_" % int(normalize)
        for ig in range(int(g["ngroups"])):
            grp = g["g%d" % ig]
            psf = g["psf"][grp].copy()
            if not normalize:
                psf[:, :, 0] *= 1.7
            r = ora.template_flux(g["images"][grp], g["weights"][grp],
                                  g["jac"][grp], psf, normalize=normalize)
            assert r["flags"] == g[tag + "flags"][ig]
            for k in ("flux", "flux_err", "chi2per", "dof"):
                assert abs(r[k] / g[tag + k][ig] - 1) < 1e-13, k


def test_runner_control_flow():
    """retry logic with stand-in fitters (no device needed)"""
    from ngmix_b200 import runners

    class Fake(object):
        def __init__(self, fail_first):
            self.calls = 0
            self.fail_first = fail_first

        def go(self, obs, guess=None):
            self.calls += 1
            return {"flags": 1 if self.calls <= self.fail_first else 0,
                    "guess": guess}

    f = Fake(2)
    res = runners.run_fitter(obs=None, fitter=f,
                             guesser=lambda obs: np.ones(3), ntry=5)
    assert res["flags"] == 0 and f.calls == 3
    f = Fake(9)
    res = runners.Runner(f, ntry=2).go(obs=None)
    assert res["flags"] == 1 and f.calls == 2 and res["guess"] is None


def test_guessers_without_device():
    """the guessers that need no measurement, against the reference's draws"""
    from ngmix_b200 import guessers
    g = _load("boot")
    got = guessers.TFluxGuesser(np.random.RandomState(8), 0.4, 100.0)()
    assert np.array_equal(got, g["gtflux1"])
    got = guessers.TFluxGuesser(np.random.RandomState(8), 0.4,
                                [100.0, 50.0, 20.0])(nrand=3)
    assert np.array_equal(got, g["gtflux3"])
    got = guessers.ParsGuesser(np.random.RandomState(9),
                               [0.1, -0.1, 0.2, 0.7, 0.5, 80.0])()
    assert np.allclose(got, g["gpars"], rtol=1e-14, atol=0)
    got = guessers.ParsGuesser(np.random.RandomState(9),
                               [0.1, -0.1, 0.2, 0.85, 0.5, 80.0],
                               widths=[0.01, 0.01, 0.05, 0.05, 0.2, 0.2])(
        nrand=4)
    assert np.allclose(got, g["gpars4"], rtol=1e-14, atol=0)
    # guessers with a prior: tests/test_guessers.py


# ------------------------------------------------------------------ GPU
@pytest.fixture(scope="module")
def ngx():
    from ngmix_b200 import _lib
    _lib.init(0)
    return _lib


def _obs(ngmix, g, i):
    def jac(j):
        return ngmix.Jacobian(row=j[0], col=j[1], dvdrow=j[2], dvdcol=j[3],
                              dudrow=j[4], dudcol=j[5])
    psf_obs = ngmix.Observation(g["psf_images"][i],
                                weight=g["psf_weights"][i],
                                jacobian=jac(g["psf_jac"][i]))
    return ngmix.Observation(g["images"][i], weight=g["weights"][i],
                             jacobian=jac(g["jac"][i]), psf=psf_obs)


@pytest.mark.gpu
def test_gpu_psf_flux_fitter(ngx):
    import ngmix_b200 as ngmix
    g = _load("psfflux")
    for normalize in (True, False):
        tag = "n%d_" % int(normalize)
        for ig in range(int(g["ngroups"])):
            ol = ngmix.ObsList()
            for i in g["g%d" % ig]:
                ppars = g["psf"][i].copy()
                if not normalize:
                    ppars[:, 0] *= 1.7
                j = g["jac"][i]
                psf_obs = ngmix.Observation(
                    np.zeros((5, 5)) + 1.0, gmix=ngmix.GMix(pars=ppars.ravel()))
                ol.append(ngmix.Observation(
                    g["images"][i], weight=g["weights"][i],
                    jacobian=ngmix.Jacobian(row=j[0], col=j[1], dvdrow=j[2],
                                            dvdcol=j[3], dudrow=j[4],
                                            dudcol=j[5]), psf=psf_obs))
            obs_in = ol if len(ol) > 1 else ol[0]
            r = ngmix.PSFFluxFitter(normalize_psf=normalize).go(obs=obs_in)
            assert r["flags"] == g[tag + "flags"][ig]
            for k in ("flux", "flux_err", "chi2per", "dof"):
                assert abs(r[k] / g[tag + k][ig] - 1) < 1e-11, k


@pytest.mark.gpu
def test_gpu_psf_guessers_match_reference_draws(ngx):
    import ngmix_b200 as ngmix
    g = _load("boot")
    obs0 = _obs(ngmix, g, 0)
    for ng in (1, 2, 3, 4, 5):
        for moms in (0, 1):
            guesser = ngmix.guessers.GMixPSFGuesser(
                np.random.RandomState(100 + ng), ng,
                guess_from_moms=bool(moms))
            got = guesser(obs0.psf).get_full_pars()
            # from_moms goes through the device's weighted moments
            assert np.allclose(got, g["gpsf_%d_%d" % (ng, moms)],
                               rtol=1e-14 if moms == 0 else 1e-9, atol=1e-15)
    got = ngmix.guessers.SimplePSFGuesser(np.random.RandomState(7))(obs0.psf)
    assert np.allclose(got, g["gsimple"], rtol=1e-14, atol=0)
    got = ngmix.guessers.SimplePSFGuesser(
        np.random.RandomState(7), guess_from_moms=True)(obs0.psf)
    assert np.allclose(got, g["gsimple_moms"], rtol=1e-9, atol=0)


@pytest.mark.gpu
def test_gpu_bootstrap_chain_matches_reference(ngx):
    """PSF EM (2 gaussians) -> psf-flux guess -> exp LM fit, seeded like the
    reference run: same psf iteration counts, same fit"""
    import ngmix_b200 as ngmix
    from ngmix_b200.runners import Runner, PSFRunner
    from ngmix_b200.bootstrap import Bootstrapper
    g = _load("boot")
    for i in range(int(g["n"])):
        rng = np.random.RandomState(9800 + i)
        obs = _obs(ngmix, g, i)
        psf_runner = PSFRunner(
            fitter=ngmix.em.EMFitter(tol=1.0e-5),
            guesser=ngmix.guessers.GMixPSFGuesser(rng=rng, ngauss=2), ntry=2)
        runner = Runner(
            fitter=ngmix.Fitter(model="exp"),
            guesser=ngmix.guessers.TPSFFluxGuesser(rng=rng, T=0.25), ntry=2)
        res = Bootstrapper(runner=runner, psf_runner=psf_runner).go(obs)
        pres = obs.psf.meta["result"]
        assert pres["flags"] == g["boot_psf_flags"][i]
        assert pres["numiter"] == g["boot_psf_numiter"][i]
        assert np.allclose(obs.psf.gmix.get_full_pars(), g["boot_psf_pars"][i],
                           rtol=1e-7, atol=1e-9)
        assert res["flags"] == g["boot_flags"][i]
        assert res["nfev"] == g["boot_nfev"][i]
        tol = np.maximum(1e-5 * np.abs(g["boot_pars"][i]),
                         1e-3 * g["boot_pars_err"][i])
        assert np.all(np.abs(res["pars"] - g["boot_pars"][i]) <= tol)
        assert abs(res["s2n"] / g["boot_s2n"][i] - 1) < 1e-5
        flux = ngmix.PSFFluxFitter().go(obs)["flux"]
        assert abs(flux / g["boot_psf_flux"][i] - 1) < 1e-7


@pytest.mark.gpu
def test_gpu_bootstrap_failed_psf(ngx):
    import ngmix_b200 as ngmix
    from ngmix_b200.runners import Runner, PSFRunner
    from ngmix_b200.bootstrap import bootstrap
    g = _load("boot")
    obs = _obs(ngmix, g, 0)
    rng = np.random.RandomState(1)
    # one iteration cannot converge: the psf fit reports EM_MAXITER
    psf_runner = PSFRunner(
        fitter=ngmix.em.EMFitter(maxiter=1, miniter=1),
        guesser=ngmix.guessers.GMixPSFGuesser(rng=rng, ngauss=2), ntry=1)
    runner = Runner(fitter=ngmix.Fitter(model="gauss"),
                    guesser=ngmix.guessers.TFluxGuesser(rng, 0.3, 100.0))
    with pytest.raises(ngmix.gexceptions.BootPSFFailure):
        bootstrap(obs, runner, psf_runner=psf_runner)


@pytest.mark.gpu
def test_gpu_psf_flux_fitter_pixel_templates(ngx):
    """PSFFluxFitter with pixel templates (obs.psf.template / obs.template,
    fitting/results.py:834-864) against the reference's fluxes"""
    import ngmix_b200 as ng
    g = _load("psfflux_template")
    for do_psf in (True, False):
        for normalize in (True, False):
            tag = "p%d_n%d_" % (int(do_psf), int(normalize))
            for ig in range(int(g["ngroups"])):
                ol = ng.ObsList()
                for i in g["g%d" % ig]:
                    rec = g["jac"][i]
                    jac = ng.Jacobian(row=rec[0], col=rec[1], dvdrow=rec[2],
                                      dvdcol=rec[3], dudrow=rec[4],
                                      dudcol=rec[5])
                    obs = ng.Observation(g["images"][i],
                                         weight=g["weights"][i], jacobian=jac)
                    if do_psf:
                        psf_obs = ng.Observation(np.zeros((5, 5)) + 1.0)
                        psf_obs.template = g["templates"][i]
                        obs.set_psf(psf_obs)
                    else:
                        obs.template = g["templates"][i]
                    ol.append(obs)
                obs_in = ol if len(ol) > 1 else ol[0]
                r = ng.PSFFluxFitter(do_psf=do_psf,
                                     normalize_psf=normalize).go(obs=obs_in)
                assert r["flags"] == int(g[tag + "flags"][ig])
                for k in ("flux", "flux_err", "chi2per", "dof"):
                    assert abs(r[k] - g[tag + k][ig]) \
                        <= 1e-11 * abs(g[tag + k][ig]), (tag, ig, k)
