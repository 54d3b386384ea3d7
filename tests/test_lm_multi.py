"""
Multi-epoch / multi-band LM fits (Fitter.go on ObsList / MultiBandObsList,
SURVEY.md 8f rank 4): the oracle against golden vectors from the reference
(CPU), the CUDA path against the same vectors and the oracle (GPU).
"""
import os
import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
PAR_TOL = 1e-5


def _golden(name="lm_multi"):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def _case(g, k):
    p = "c%d_" % k
    d = {key[len(p):]: g[key] for key in g.files if key.startswith(p)}
    d["model"] = str(d["model"])
    d["nband"] = int(d["nband"])
    d["prior"] = int(d["prior"])
    return d


def _ora_priors(ora, c):
    if not c["prior"]:
        return None, None
    pr = ora.make_prior("erf")
    fps = [ora.Prior1D(2, 0, -1.0e2, 1.0, 1.0e9, 1.0e8)
           for _ in range(c["nband"])]
    return pr, fps


def _check(r, c, np_, exact_counts=True):
    assert int(r["flags"]) == int(c["flags"])
    if exact_counts:
        assert int(r["nfev"]) == int(c["nfev"])
        assert int(r["ier"]) == int(c["ier"])
    err = c["pars_err"]
    tol = np.maximum(PAR_TOL * np.abs(c["pars"]), 1e-3 * err)
    assert np.all(np.abs(r["pars"][:np_] - c["pars"]) <= tol)
    cov = np.asarray(r["pars_cov"][:np_ * np_]).reshape(np_, np_)
    scale = np.sqrt(np.outer(np.diag(c["pars_cov"]), np.diag(c["pars_cov"])))
    assert np.abs(cov - c["pars_cov"]).max() <= 1e-3 * scale.max()
    if int(c["flags"]) == 0:
        assert int(r["npix"]) == int(c["npix"])
        assert int(r["dof"]) == int(c["dof"])
        for key in ("lnprob", "s2n", "chi2per", "T", "T_err"):
            assert abs(float(r[key]) - float(c[key])) \
                <= 1e-5 * max(abs(float(c[key])), 1.0), key


@pytest.mark.parametrize("name", ["lm_multi", "lm_multi_wide",
                                  "lm_multi_xwide"])
def test_oracle_multi_matches_reference(ora, name):
    g = _golden(name)
    for k in range(int(g["ncases"])):
        c = _case(g, k)
        pr, fps = _ora_priors(ora, c)
        r = ora.lm_fit_multi(c["model"], c["images"], c["weights"], c["jac"],
                             c["psf"], c["bands"], c["nband"], c["guess"],
                             prior=pr, Fpriors=fps)
        _check(r, c, 5 + c["nband"])


def _ora_priors_fd(ora, c):
    """the priors of gen_golden.gen_lm_multi_fd"""
    if not c["prior"]:
        return None, None
    pr = ora.make_prior("flat")
    pr.fracdev = ora.Prior1D(3, 0, 0.5, 0.1, 0, 0)
    pr.logTratio = ora.Prior1D(3, 0, 0.0, 0.5, 0, 0)
    fps = [ora.Prior1D(1, 0, -100.0, 1.0e9, 0, 0) for _ in range(c["nband"])]
    return pr, fps


NSHARED = {"exp": 5, "bdf": 6, "bd": 7}


def test_oracle_multi_fd_matches_reference(ora):
    """bdf / bd objects with several stamps and bands, and exp objects with
    analytic_jacobian=False: the lmdif path of Fitter.go on MultiBandObsList"""
    g = _golden("lm_multi_fd")
    for k in range(int(g["ncases"])):
        c = _case(g, k)
        pr, fps = _ora_priors_fd(ora, c)
        r = ora.lm_fit_multi(c["model"], c["images"], c["weights"], c["jac"],
                             c["psf"], c["bands"], c["nband"], c["guess"],
                             prior=pr, Fpriors=fps, analytic_jacobian=False)
        _check(r, c, NSHARED[c["model"]] + c["nband"])


def _bcase(g, k):
    p = "c%d_" % k
    d = {key[len(p):]: g[key] for key in g.files if key.startswith(p)}
    d["model"] = str(d["model"])
    d["nband"] = int(d["nband"])
    d["analytic"] = bool(int(d["analytic"]))
    d["rep"] = k % 2
    d["bounds_list"] = [tuple(None if np.isnan(v) else float(v) for v in bb)
                        for bb in d["bounds"]]
    return d


def test_oracle_multi_bounded_matches_reference(ora):
    """prior.bounds on ObsList / MultiBandObsList fits (leastsqbound's
    transforms around the multi-stamp residuals): exp through lmder and
    lmdif, bdf through lmdif; golden from the reference with the 1-based
    ipvt its covariance code assumes (gen_golden.gen_lm_multi_bounds)"""
    g = _golden("lm_multi_bounds")
    for k in range(int(g["ncases"])):
        c = _bcase(g, k)
        pr = ora.make_prior("flat")
        if c["model"] == "exp":
            pr.T = ora.Prior1D(3, 0, [0.4, 0.5][c["rep"]], 1.0, 0, 0)
            fps = [ora.Prior1D(3, 0, 100.0, 100.0, 0, 0)
                   for _ in range(c["nband"])]
        else:
            pr.fracdev = ora.Prior1D(3, 0, 0.5, 0.3, 0, 0)
            fps = [ora.Prior1D(1, 0, -100.0, 1.0e9, 0, 0)
                   for _ in range(c["nband"])]
        r = ora.lm_fit_multi(c["model"], c["images"], c["weights"], c["jac"],
                             c["psf"], c["bands"], c["nband"], c["guess"],
                             prior=pr, Fpriors=fps,
                             analytic_jacobian=c["analytic"],
                             bounds=c["bounds_list"])
        _check(r, c, NSHARED[c["model"]] + c["nband"])


# ------------------------------------------------------------------ GPU
@pytest.fixture(scope="module")
def ngx():
    from ngmix_b200 import _lib
    _lib.init(0)
    return _lib


def _ngx_prior(c):
    import ngmix_b200 as ngmix
    if not c["prior"]:
        return None
    return ngmix.joint_prior.PriorSimpleSep(
        ngmix.priors.CenPrior(0.0, 0.0, 0.263, 0.263),
        ngmix.priors.GPriorBA(0.3),
        ngmix.priors.TwoSidedErf(-1.0, 0.1, 1.0e3, 1.0),
        [ngmix.priors.TwoSidedErf(-1.0e2, 1.0, 1.0e9, 1.0e8)
         for _ in range(c["nband"])])


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["lm_multi", "lm_multi_wide",
                                  "lm_multi_xwide"])
def test_gpu_multi_matches_reference(ngx, name):
    """1-3 bands (8-parameter records) and 4-7 bands (12-parameter records)"""
    import ngmix_b200 as ngmix
    g = _golden(name)
    for k in range(int(g["ncases"])):
        c = _case(g, k)
        ns = c["images"].shape[0]
        rec = ngmix.fit_batch_multi(
            c["model"], [0, ns], c["bands"], c["nband"], c["images"],
            c["weights"], c["jac"], c["psf"], c["guess"][None],
            prior=_ngx_prior(c))[0]
        _check(rec, c, 5 + c["nband"])


@pytest.mark.gpu
def test_gpu_fitter_go_obslist_and_mbobs(ngx):
    """the reference-facing call: Fitter.go(ObsList / MultiBandObsList)"""
    import ngmix_b200 as ngmix
    g = _golden()
    gw = _golden("lm_multi_wide")
    for c in [_case(g, k) for k in (1, 7, 13, 16)] \
            + [_case(gw, k) for k in (0, 4, 12)]:
        mbobs = ngmix.MultiBandObsList()
        i = 0
        for b in range(c["nband"]):
            ol = ngmix.ObsList()
            while i < c["bands"].size and c["bands"][i] == b:
                j = c["jac"][i]
                jac = ngmix.Jacobian(row=j[0], col=j[1], dvdrow=j[2],
                                     dvdcol=j[3], dudrow=j[4], dudcol=j[5])
                psf_obs = ngmix.Observation(
                    np.zeros((5, 5)) + 1.0,
                    gmix=ngmix.GMix(pars=c["psf"][i].ravel()))
                ol.append(ngmix.Observation(c["images"][i],
                                            weight=c["weights"][i],
                                            jacobian=jac, psf=psf_obs))
                i += 1
            mbobs.append(ol)
        fitter = ngmix.Fitter(model=c["model"], prior=_ngx_prior(c))
        obs_in = mbobs if c["nband"] > 1 else mbobs[0]
        res = fitter.go(obs=obs_in, guess=c["guess"])
        np_ = 5 + c["nband"]
        assert res["flags"] == int(c["flags"])
        assert res["nfev"] == int(c["nfev"]) and res["ier"] == int(c["ier"])
        assert np.allclose(res["pars"], c["pars"], rtol=1e-5,
                           atol=1e-3 * c["pars_err"].max())
        assert res["pars_cov"].shape == (np_, np_)
        assert np.allclose(np.atleast_1d(res["flux"]), c["flux"], rtol=1e-5)
        assert np.allclose(np.atleast_1d(res["flux_err"]), c["flux_err"],
                           rtol=1e-3)
        assert abs(res["s2n"] - float(c["s2n"])) < 1e-5 * float(c["s2n"])
        gm = res.get_convolved_gmix(band=c["nband"] - 1, obsnum=0)
        assert gm.get_data().size == c["psf"].shape[1] * \
            {"gauss": 1, "exp": 6, "dev": 10}[c["model"]]


@pytest.mark.gpu
def test_gpu_multi_batch_vs_oracle(ngx, ora):
    """a batch of objects with ragged epoch counts, two bands, masked pixels"""
    import ngmix_b200 as ngmix
    from ngmix_b200 import sims
    rng = np.random.RandomState(77)
    nobj, nband = 12, 2
    b = sims.make_lm_batch(60, seed0=77000, dims=(24, 24), noise=0.05)
    starts, bands, guesses = [0], [], []
    k = 0
    for i in range(nobj):
        ne = [int(rng.randint(1, 4)), int(rng.randint(1, 3))]
        for band in range(nband):
            bands += [band] * ne[band]
        k += sum(ne)
        starts.append(k)
        gs = b["guess"][starts[i]].copy()
        guesses.append(np.append(gs, gs[5] * 1.02))
    ns = starts[-1]
    images, weights = b["images"][:ns].copy(), b["weights"][:ns].copy()
    weights[3][5:9, :] = 0.0
    jac, psf = b["jac"][:ns], b["psf"][:ns]
    guesses = np.array(guesses)
    recs = ngmix.fit_batch_multi("exp", starts, bands, nband, images, weights,
                                 jac, psf, guesses)
    assert recs.shape == (nobj,)
    for i in range(nobj):
        sl = slice(starts[i], starts[i + 1])
        r = ora.lm_fit_multi("exp", images[sl], weights[sl], jac[sl], psf[sl],
                             np.array(bands[sl]), nband, guesses[i])
        assert recs["flags"][i] == r["flags"]
        assert recs["npix"][i] == r["npix"]
        if r["flags"] == 0:
            tol = np.maximum(PAR_TOL * np.abs(r["pars"][:7]),
                             1e-3 * r["pars_err"][:7])
            assert np.all(np.abs(recs["pars"][i][:7] - r["pars"][:7]) <= tol)
            assert abs(recs["lnprob"][i] - r["lnprob"]) \
                <= 1e-6 * abs(r["lnprob"])
    # an object without usable pixels is reported, not fitted
    weights[starts[2]:starts[3]] = 0.0
    recs = ngmix.fit_batch_multi("exp", starts, bands, nband, images, weights,
                                 jac, psf, guesses)
    assert recs["flags"][2] == 1 and recs["nfev"][2] == 0


@pytest.mark.gpu
def test_gpu_multi_pageable_equals_pinned(ngx):
    """pageable planes are staged through pinned slots by the copy pool
    (csrc/ngx_copy_pool.h): the same records as from pinned planes"""
    import ngmix_b200 as ngmix
    from ngmix_b200 import sims
    nobj, nep = 700, 3
    b = sims.make_lm_batch(nobj * nep, seed0=78000, noise=0.05)
    starts = np.arange(nobj + 1) * nep
    bands = np.zeros(nobj * nep, dtype="i4")
    guess = b["guess"][::nep].copy()
    args = (b["jac"], b["psf"], guess)
    pageable = ngmix.fit_batch_multi("exp", starts, bands, 1, b["images"],
                                     b["weights"], *args)
    pim = ngmix.pinned_copy(b["images"])
    pwt = ngmix.pinned_copy(b["weights"])
    pinned = ngmix.fit_batch_multi("exp", starts, bands, 1, pim, pwt, *args)
    assert pageable.tobytes() == pinned.tobytes()
    assert (pinned["flags"] == 0).sum() > 650


def _ngx_prior_fd(c):
    """the priors of gen_golden.gen_lm_multi_fd"""
    import ngmix_b200 as ngmix
    if not c["prior"]:
        return None
    pr = ngmix.priors
    cen, gp = pr.CenPrior(0.0, 0.0, 0.263, 0.263), pr.GPriorBA(0.3)
    Tp = pr.FlatPrior(-1.0, 1.0e3)
    Fp = [pr.FlatPrior(-100.0, 1.0e9) for _ in range(c["nband"])]
    if c["model"] == "bdf":
        return ngmix.joint_prior.PriorBDFSep(cen, gp, Tp,
                                             pr.Normal(0.5, 0.1), Fp)
    if c["model"] == "bd":
        return ngmix.joint_prior.PriorBDSep(cen, gp, Tp, pr.Normal(0.0, 0.5),
                                            pr.Normal(0.5, 0.1), Fp)
    return ngmix.joint_prior.PriorSimpleSep(cen, gp, Tp, Fp)


@pytest.mark.gpu
def test_gpu_multi_fd_matches_reference(ngx):
    """bdf / bd objects with several stamps and bands (up to 11 parameters)
    and exp objects with analytic_jacobian=False: the lmdif path of Fitter.go
    on MultiBandObsList.  The forward-difference quotients carry the 1e-8
    noise of the model's last bits, so the covariance is compared at 2e-3."""
    import ngmix_b200 as ngmix
    g = _golden("lm_multi_fd")
    nsame = 0
    ncases = int(g["ncases"])
    for k in range(ncases):
        c = _case(g, k)
        ns = c["images"].shape[0]
        np_ = NSHARED[c["model"]] + c["nband"]
        rec = ngmix.fit_batch_multi(
            c["model"], [0, ns], c["bands"], c["nband"], c["images"],
            c["weights"], c["jac"], c["psf"], c["guess"][None],
            prior=_ngx_prior_fd(c), analytic_jacobian=False)[0]
        assert int(rec["flags"]) == int(c["flags"])
        nsame += int(int(rec["nfev"]) == int(c["nfev"])
                     and int(rec["ier"]) == int(c["ier"]))
        tol = np.maximum(PAR_TOL * np.abs(c["pars"]), 2e-2 * c["pars_err"])
        assert np.all(np.abs(rec["pars"][:np_] - c["pars"]) <= tol)
        assert int(rec["npix"]) == int(c["npix"])
        assert abs(float(rec["lnprob"]) - float(c["lnprob"])) \
            <= 1e-5 * abs(float(c["lnprob"]))
        assert abs(float(rec["flux"]) - float(c["flux"][0])) \
            <= 1e-4 * abs(float(c["flux"][0]))
        cov = np.asarray(rec["pars_cov"][:np_ * np_]).reshape(np_, np_)
        sc = np.sqrt(np.outer(np.diag(c["pars_cov"]), np.diag(c["pars_cov"])))
        assert np.abs(cov - c["pars_cov"]).max() <= 2e-3 * sc.max()
    # evaluation counts follow the reference up to forward-difference noise
    assert nsame >= ncases - 2


@pytest.mark.gpu
def test_gpu_multi_bounded_matches_reference(ngx):
    """bounded fits of objects with several stamps (prior.bounds through
    fit_batch_multi): analytic cases follow the reference in flags, nfev, ier
    and parameters; forward-difference cases up to the fd noise floor"""
    import ngmix_b200 as ngmix
    pr = ngmix.priors
    g = _golden("lm_multi_bounds")
    nsame = nfd = 0
    for k in range(int(g["ncases"])):
        c = _bcase(g, k)
        ns = c["images"].shape[0]
        np_ = NSHARED[c["model"]] + c["nband"]
        cen, gp = pr.CenPrior(0.0, 0.0, 0.263, 0.263), pr.GPriorBA(0.3)
        if c["model"] == "exp" and c["rep"] == 0:
            prior = ngmix.joint_prior.PriorSimpleSep(
                cen, gp, pr.Normal(0.4, 1.0, bounds=[0.01, 100.0]),
                [pr.Normal(100.0, 100.0, bounds=[1.0, 1.0e9])
                 for _ in range(c["nband"])])
        elif c["model"] == "exp":
            prior = ngmix.joint_prior.PriorSimpleSep(
                cen, gp, pr.Normal(0.5, 1.0, bounds=[0.05, 5.0]),
                [pr.Normal(100.0, 100.0, bounds=[1.0, None])
                 for _ in range(c["nband"])])
        else:
            prior = ngmix.joint_prior.PriorBDFSep(
                cen, gp, pr.FlatPrior(-1.0, 1.0e3),
                pr.Normal(0.5, 0.3, bounds=[0.0, 1.0]),
                [pr.FlatPrior(-100.0, 1.0e9) for _ in range(c["nband"])])
        assert [tuple(b) for b in prior.bounds] == c["bounds_list"]
        rec = ngmix.fit_batch_multi(
            c["model"], [0, ns], c["bands"], c["nband"], c["images"],
            c["weights"], c["jac"], c["psf"], c["guess"][None], prior=prior,
            analytic_jacobian=c["analytic"])[0]
        if c["analytic"]:
            _check(rec, c, np_)
            continue
        nfd += 1
        assert int(rec["flags"]) == int(c["flags"])
        nsame += int(int(rec["nfev"]) == int(c["nfev"])
                     and int(rec["ier"]) == int(c["ier"]))
        tol = np.maximum(PAR_TOL * np.abs(c["pars"]), 2e-2 * c["pars_err"])
        assert np.all(np.abs(rec["pars"][:np_] - c["pars"]) <= tol)
        assert abs(float(rec["lnprob"]) - float(c["lnprob"])) \
            <= 1e-5 * abs(float(c["lnprob"]))
        cov = np.asarray(rec["pars_cov"][:np_ * np_]).reshape(np_, np_)
        sc = np.sqrt(np.outer(np.diag(c["pars_cov"]), np.diag(c["pars_cov"])))
        assert np.abs(cov - c["pars_cov"]).max() <= 2e-3 * sc.max()
    assert nsame >= nfd - 1


@pytest.mark.gpu
def test_gpu_fitter_go_mbobs_bdf(ngx):
    """Fitter(model='bdf', prior=PriorBDFSep with a flux prior per band).go on
    a MultiBandObsList"""
    import ngmix_b200 as ngmix
    g = _golden("lm_multi_fd")
    c = _case(g, 4)         # bdf, 3 bands, 4 stamps
    mbobs = ngmix.MultiBandObsList()
    i = 0
    for b in range(c["nband"]):
        ol = ngmix.ObsList()
        while i < c["bands"].size and c["bands"][i] == b:
            j = c["jac"][i]
            jac = ngmix.Jacobian(row=j[0], col=j[1], dvdrow=j[2], dvdcol=j[3],
                                 dudrow=j[4], dudcol=j[5])
            psf_obs = ngmix.Observation(
                np.zeros((5, 5)) + 1.0,
                gmix=ngmix.GMix(pars=c["psf"][i].ravel()))
            ol.append(ngmix.Observation(c["images"][i], weight=c["weights"][i],
                                        jacobian=jac, psf=psf_obs))
            i += 1
        mbobs.append(ol)
    res = ngmix.Fitter(model="bdf", prior=_ngx_prior_fd(c)).go(
        obs=mbobs, guess=c["guess"])
    assert res["flags"] == 0 and res["pars"].size == 9
    assert np.allclose(res["pars"], c["pars"], rtol=1e-5,
                       atol=2e-2 * c["pars_err"].max())
    assert np.allclose(res["flux"], c["flux"], rtol=1e-4)
    assert res["flux_cov"].shape == (3, 3)
    gm = res.get_convolved_gmix(band=2, obsnum=0)
    assert gm.get_data().size == 48
