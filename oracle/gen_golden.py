"""
TEST INFRASTRUCTURE - golden fixture generator.

Runs the UNMODIFIED reference (/root/reference/ngmix, numba + scipy) on seeded
synthetic inputs and writes input/output vectors to tests/golden/*.npz.  The
reference is a Python package and cannot travel to the GPU box, so the
vectors are committed together with this script (task statement, section 3).

Usage (in the build container only; needs /root/reference):

    python oracle/gen_golden.py [--only NAME ...]

galsim is not installed; ngmix/__init__.py imports ngmix.metacal, which does
a module-level `import galsim`, so an EMPTY stub package is put on sys.path.
No galsim function is ever called by anything used here.
"""
import argparse
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)


def _import_reference():
    stub = tempfile.mkdtemp(prefix="galsim_stub_")
    os.makedirs(os.path.join(stub, "galsim"))
    open(os.path.join(stub, "galsim", "__init__.py"), "w").close()
    os.environ.setdefault("NUMBA_CACHE_DIR",
                          os.path.join(tempfile.gettempdir(), "numba_cache"))
    sys.dont_write_bytecode = True
    sys.path.insert(0, stub)
    sys.path.insert(0, "/root/reference")
    import ngmix
    assert ngmix.__file__.startswith("/root/reference"), ngmix.__file__
    return ngmix


ngmix = _import_reference()
from ngmix_b200 import sims  # noqa: E402


def jac_from_record(rec):
    return ngmix.Jacobian(row=rec[0], col=rec[1], dvdrow=rec[2],
                          dvdcol=rec[3], dudrow=rec[4], dudcol=rec[5])


def gm_records(gm):
    """(ngauss, 12) float view of a gmix record array, norm_set dropped"""
    d = gm.get_data()
    names = ["p", "row", "col", "irr", "irc", "icc", "det",
             "drr", "drc", "dcc", "norm", "pnorm"]
    return np.array([[d[n][i] for n in names] for i in range(d.size)])


def save(name, **arrs):
    path = os.path.join(GOLDEN, name + ".npz")
    np.savez_compressed(path, **arrs)
    print("wrote %s (%.1f KB)" % (path, os.path.getsize(path) / 1024))


# ---------------------------------------------------------------- fastexp
def gen_fastexp():
    from ngmix.fastexp_nb import (
        fexp_arr, apod_window, apod_window_deriv, _EXP5_SMOOTH_COEFFS,
        _EXP_LOOKUP,
    )
    from numba import vectorize

    rng = np.random.RandomState(1)
    x = np.concatenate([
        np.array([-0.25, -0.5, -0.75, -3.3, -7.8864744, -12.4999, 0.0,
                  -1e-300, -1e-17, -12.5]),
        -0.5 * np.arange(0, 26),            # integers and half integers
        np.nextafter(-0.5 * np.arange(0, 26) - 0.0, 1.0),
        np.nextafter(-0.5 * np.arange(0, 26) - 0.0, -20.0),
        rng.uniform(-12.5, 0.0, size=4000),
    ])
    x = x[(x >= -12.5) & (x <= 0)]

    @vectorize
    def aw(c):
        return apod_window(c)

    @vectorize
    def awd(c):
        return apod_window_deriv(c)

    chi2 = np.concatenate([np.array([20.0, 21.0, 22.5, 24.0, 25.0]),
                           rng.uniform(20.0, 25.0, size=500)])
    save("fastexp", x=x, fexp=fexp_arr(x), chi2=chi2, apod=aw(chi2),
         apod_deriv=awd(chi2), coeffs=_EXP5_SMOOTH_COEFFS,
         lookup=_EXP_LOOKUP)


# ---------------------------------------------------------------- gmix
def gen_gmix():
    from ngmix.gmix.gmix_nb import get_cm_Tfactor
    rng = np.random.RandomState(2)
    out = {}
    models = ["gauss", "turb", "exp", "dev", "bdf", "bd"]
    for model in models:
        npars = ngmix.gmix.get_model_npars(model)
        pars = np.zeros((8, npars))
        recs = []
        for i in range(8):
            p = np.zeros(npars)
            p[0:2] = rng.uniform(-0.5, 0.5, size=2)
            p[2:4] = rng.uniform(-0.4, 0.4, size=2)
            if i == 0:
                p[2:4] = 0.0            # g == 0 branch
            p[4] = rng.uniform(0.1, 4.0)
            if model == "bdf":
                p[5] = rng.uniform(0.0, 1.0)
            elif model == "bd":
                p[5] = rng.uniform(-0.5, 0.5)
                p[6] = rng.uniform(0.0, 1.0)
            p[-1] = rng.uniform(1.0, 500.0)
            pars[i] = p
            gm = ngmix.GMixModel(p, model)
            gm.set_norms()
            recs.append(gm_records(gm))
        out["pars_" + model] = pars
        out["recs_" + model] = np.array(recs)

    # coellip (3 gaussians) and full
    cpars = np.zeros((4, 10))
    crecs = []
    for i in range(4):
        p = np.concatenate([rng.uniform(-0.3, 0.3, size=2),
                            rng.uniform(-0.3, 0.3, size=2),
                            rng.uniform(0.1, 2.0, size=3),
                            rng.uniform(0.1, 10.0, size=3)])
        cpars[i] = p
        gm = ngmix.GMixCoellip(p)
        gm.set_norms()
        crecs.append(gm_records(gm))
    out["pars_coellip"] = cpars
    out["recs_coellip"] = np.array(crecs)

    # cm: fracdev, TdByTe, pars
    cm_in = np.zeros((4, 8))
    cm_recs = []
    cm_tf = []
    for i in range(4):
        fracdev = rng.uniform(0, 1)
        TdByTe = rng.uniform(0.5, 3.0)
        p = np.array([0.1, -0.1, rng.uniform(-0.3, 0.3),
                      rng.uniform(-0.3, 0.3), rng.uniform(0.2, 2.0), 100.0])
        cm_in[i] = np.concatenate([[fracdev, TdByTe], p])
        gm = ngmix.gmix.GMixCM(fracdev, TdByTe, p)
        gm.set_norms()
        cm_recs.append(gm_records(gm))
        cm_tf.append(get_cm_Tfactor(fracdev, TdByTe))
    out["cm_in"] = cm_in
    out["recs_cm"] = np.array(cm_recs)
    out["cm_tfactor"] = np.array(cm_tf)

    # convolution: exp (6) x 3-gaussian psf with offset centres
    conv_obj = np.zeros((4, 6))
    conv_psf = np.zeros((4, 18))
    conv_out = []
    for i in range(4):
        p = np.array([0.2, -0.1, rng.uniform(-0.3, 0.3),
                      rng.uniform(-0.3, 0.3), rng.uniform(0.2, 2.0), 77.0])
        conv_obj[i] = p
        ppars = []
        for k in range(3):
            ppars += [rng.uniform(0.1, 1.0), rng.uniform(-0.05, 0.05),
                      rng.uniform(-0.05, 0.05), rng.uniform(0.05, 0.3),
                      rng.uniform(-0.01, 0.01), rng.uniform(0.05, 0.3)]
        conv_psf[i] = ppars
        gm = ngmix.GMixModel(p, "exp").convolve(ngmix.GMix(pars=ppars))
        gm.set_norms()
        conv_out.append(gm_records(gm))
    out["conv_obj"] = conv_obj
    out["conv_psf"] = conv_psf
    out["conv_recs"] = np.array(conv_out)
    save("gmix", **out)


# ---------------------------------------------------------------- pixels
def gen_pixels():
    rng = np.random.RandomState(3)
    H, W = 11, 13
    image = rng.normal(size=(H, W))
    weight = rng.uniform(0.5, 2.0, size=(H, W))
    weight[2, 3] = 0.0
    weight[0, 0] = -1.0
    weight[10, 12] = 0.0
    jac = ngmix.Jacobian(row=4.7, col=6.1, dvdrow=0.26, dvdcol=0.01,
                         dudrow=-0.02, dudcol=0.27)
    rec = np.array([jac._data[n][0] for n in jac._data.dtype.names])
    out = dict(image=image, weight=weight, jac=rec)
    for tag, ignore in (("ign", True), ("all", False)):
        pix = ngmix.pixels.make_pixels(image, weight, jac,
                                       ignore_zero_weight=ignore)
        out["pixels_" + tag] = np.array(
            [pix[n] for n in ("u", "v", "area", "val", "ierr")]).T
    coords = ngmix.pixels.make_coords((H, W), jac)
    out["coords"] = np.array([coords[n] for n in ("u", "v", "area")]).T
    save("pixels", **out)


# ---------------------------------------------------------------- render
def gen_render():
    from ngmix.fitting.derivs_nb import deriv_images
    from ngmix.fitting.results import get_model_deriv_data
    rng = np.random.RandomState(4)
    out = {}

    # Appendix B mixture of SURVEY.md + a few random ones, incl. an offset
    # non-diagonal jacobian and a psf with off-centre components
    cases = []
    jac0 = sims.diag_jacobian(23.5, 23.5, 0.263)
    cases.append(([0.1, -0.2, 0.1, 0.05, 0.4, 100.0], "exp",
                  sims.simple_full_pars([0, 0, 0.01, -0.01, 0.27, 1.0],
                                        "turb"), jac0, (48, 48)))
    for i in range(3):
        jr = ngmix.Jacobian(row=15.3 + i, col=16.9 - i, dvdrow=0.25,
                            dvdcol=0.02 * i, dudrow=-0.015, dudcol=0.27)
        jrec = np.array([jr._data[n][0] for n in jr._data.dtype.names])
        psf = np.array([
            [0.6, 0.02, -0.01, 0.09, 0.004, 0.11],
            [0.3, -0.03, 0.02, 0.2, -0.01, 0.18],
            [0.1, 0.0, 0.0, 0.6, 0.02, 0.5]])
        model = ["exp", "dev", "gauss"][i]
        pars = [rng.uniform(-0.3, 0.3), rng.uniform(-0.3, 0.3),
                rng.uniform(-0.3, 0.3), rng.uniform(-0.3, 0.3),
                rng.uniform(0.2, 1.5), rng.uniform(10, 100)]
        cases.append((pars, model, psf, jrec, (32, 34)))

    for k, (pars, model, psf, jrec, dims) in enumerate(cases):
        gm0 = ngmix.GMixModel(pars, model)
        gpsf = ngmix.GMix(pars=psf.ravel())
        gm = gm0.convolve(gpsf)
        jac = jac_from_record(jrec)
        fast = gm.make_image(dims, jacobian=jac, fast_exp=True)
        exact = gm.make_image(dims, jacobian=jac)
        rs = np.random.RandomState(100 + k)
        image = exact + rs.normal(scale=0.05, size=dims)
        weight = np.zeros(dims) + rs.uniform(2.0, 6.0)
        weight[3, 4] = 0.0
        obs = ngmix.Observation(image, weight=weight, jacobian=jac)
        ll = gm.get_loglike(obs, more=True)
        s2n_sum = gm.get_model_s2n_sum(obs)
        fdiff = np.zeros(image.size + 13)
        gm.fill_fdiff(obs, fdiff, start=13)

        # derivative images on ALL pixels
        coords = ngmix.pixels.make_coords(dims, jac)
        gpars, dcov = get_model_deriv_data(
            gm0=gm0, gmc=gm, g1=pars[2], g2=pars[3], T=pars[4])
        dout = np.zeros((6, coords.size))
        deriv_images(gpars, dcov, coords["v"].copy(), coords["u"].copy(),
                     coords["area"].copy(), dout)

        p = "c%d_" % k
        out[p + "pars"] = np.array(pars)
        out[p + "model"] = np.array(model)
        out[p + "psf"] = psf
        out[p + "jac"] = jrec
        out[p + "dims"] = np.array(dims)
        out[p + "fast"] = fast
        out[p + "exact"] = exact
        out[p + "image"] = image
        out[p + "weight"] = weight
        out[p + "loglike"] = np.array(
            [ll["loglike"], ll["s2n_numer"], ll["s2n_denom"], ll["npix"]])
        out[p + "s2n_sum"] = np.array(s2n_sum)
        out[p + "fdiff"] = fdiff
        out[p + "gpars"] = gpars
        out[p + "dcov"] = dcov
        out[p + "deriv"] = dout
    out["ncases"] = np.array(len(cases))
    save("render", **out)


# ---------------------------------------------------------------- LM
LM_KEYS_F = ["lnprob", "s2n_numer", "s2n_denom", "chi2per", "s2n", "T",
             "T_err", "flux", "flux_err"]


def _get_priors(kind, rng):
    from ngmix import priors, joint_prior
    if kind == "none":
        return None
    cen = priors.CenPrior(0.0, 0.0, 0.263, 0.263, rng=rng)
    g = priors.GPriorBA(0.3, rng=rng)
    if kind == "flat":
        T = priors.FlatPrior(-1.0, 1.0e3, rng=rng)
        F = priors.FlatPrior(-100.0, 1.0e9, rng=rng)
    elif kind == "erf":
        T = priors.TwoSidedErf(-1.0, 0.1, 1.0e3, 1.0, rng=rng)
        F = priors.TwoSidedErf(-1.0e2, 1.0, 1.0e9, 1.0e8, rng=rng)
    elif kind == "lognormal":
        T = priors.LogNormal(0.5, 0.3, rng=rng)
        F = priors.LogNormal(100.0, 50.0, rng=rng, shift=-5.0)
    elif kind == "trunc":
        T = priors.TruncatedGaussian(0.5, 0.4, 0.0, 5.0, rng=rng)
        F = priors.LogNormal(100.0, 50.0, rng=rng)
    else:
        raise ValueError(kind)
    return joint_prior.PriorSimpleSep(cen, g, T, F)


def _run_lm_set(batch, prior_kind, analytic=True, model=None):
    from ngmix.fitting import Fitter
    model = model or batch["model"]
    n = batch["images"].shape[0]
    npars = batch["guess"].shape[1]
    res = dict(
        flags=np.zeros(n, "i4"), nfev=np.zeros(n, "i4"),
        ier=np.zeros(n, "i4"), npix=np.zeros(n, "i4"),
        dof=np.zeros(n, "i4"),
        pars=np.zeros((n, npars)), pars_err=np.zeros((n, npars)),
        pars_cov0=np.zeros((n, npars, npars)),
        pars_cov=np.zeros((n, npars, npars)),
    )
    for k in LM_KEYS_F:
        res[k] = np.full(n, np.nan)
    rng = np.random.RandomState(99)
    prior = _get_priors(prior_kind, rng)
    fitter = Fitter(model=model, prior=prior, analytic_jacobian=analytic)
    for i in range(n):
        jac = jac_from_record(batch["jac"][i])
        psf_gm = ngmix.GMix(pars=batch["psf"][i].ravel())
        psf_obs = ngmix.Observation(np.zeros((5, 5)) + 1.0,
                                    jacobian=ngmix.DiagonalJacobian(
                                        row=2, col=2, scale=0.263),
                                    gmix=psf_gm)
        obs = ngmix.Observation(batch["images"][i],
                                weight=batch["weights"][i],
                                jacobian=jac, psf=psf_obs)
        r = fitter.go(obs, batch["guess"][i].copy())
        res["flags"][i] = r["flags"]
        res["nfev"][i] = r["nfev"]
        res["ier"][i] = r.get("ier", -1)
        res["pars"][i] = r["pars"]
        res["pars_err"][i] = r.get("pars_err", np.nan)
        c0 = r["pars_cov0"]
        res["pars_cov0"][i] = np.nan if c0 is None else c0
        res["pars_cov"][i] = r["pars_cov"]
        if r["flags"] == 0:
            res["npix"][i] = r["npix"]
            res["dof"][i] = r["dof"]
            for k in LM_KEYS_F:
                res[k][i] = r[k]
    return res


def gen_lm_exp():
    out = {}
    # config 1: one stamp, 1-gaussian psf, RandomState(42)
    b1 = sims.make_lm_batch(1, seed0=42, psf_model="gauss")
    # config 2 subsets: high S/N and low S/N, 3-gaussian psf
    bhi = sims.make_lm_batch(24, seed0=1000, noise=0.01)
    blo = sims.make_lm_batch(24, seed0=2000, noise=1.0, guess_frac=0.3)
    for tag, b, priors_ in (("c1", b1, ["none"]),
                            ("hi", bhi, ["none", "flat"]),
                            ("lo", blo, ["none", "flat", "erf"])):
        for k in ("images", "weights", "jac", "psf", "guess", "truth"):
            out["%s_%s" % (tag, k)] = b[k]
        for pk in priors_:
            r = _run_lm_set(b, pk)
            for k, v in r.items():
                out["%s_%s_%s" % (tag, pk, k)] = v
            print(tag, pk, "flags", np.bincount(r["flags"] != 0),
                  "nfev med", np.median(r["nfev"]), "ier", r["ier"][:8])
    save("lm_exp", **out)

    # intermediate trace: calc_fdiff / calc_jacobian at the guess, 2 objects
    from ngmix.fitting.results import FitModel
    tr = {}
    rng = np.random.RandomState(99)
    for pk in ("none", "flat", "erf"):
        prior = _get_priors(pk, rng)
        for i in range(2):
            b = blo
            jac = jac_from_record(b["jac"][i])
            psf_obs = ngmix.Observation(
                np.ones((5, 5)),
                jacobian=ngmix.DiagonalJacobian(row=2, col=2, scale=0.263),
                gmix=ngmix.GMix(pars=b["psf"][i].ravel()))
            obs = ngmix.Observation(b["images"][i], weight=b["weights"][i],
                                    jacobian=jac, psf=psf_obs)
            fm = FitModel(obs=obs, model="exp", guess=b["guess"][i],
                          prior=prior)
            fd = fm.calc_fdiff(b["guess"][i])
            jc = fm.calc_jacobian(b["guess"][i])
            tr["%s_%d_fdiff" % (pk, i)] = fd
            tr["%s_%d_jtj" % (pk, i)] = jc.T @ jc
            tr["%s_%d_jtf" % (pk, i)] = jc.T @ fd
            tr["%s_%d_jac_prior" % (pk, i)] = jc[:fm.n_prior_pars + 2]
    save("lm_trace", **tr)


def gen_lm_other():
    """gauss and dev models (analytic jacobian) + lmdif path for exp"""
    out = {}
    b = sims.make_lm_batch(8, seed0=3000, noise=0.1, model="gauss",
                           dims=(32, 32))
    for k in ("images", "weights", "jac", "psf", "guess"):
        out["gauss_" + k] = b[k]
    for k, v in _run_lm_set(b, "none").items():
        out["gauss_none_" + k] = v
    b = sims.make_lm_batch(8, seed0=4000, noise=0.1, model="dev",
                           dims=(32, 32))
    for k in ("images", "weights", "jac", "psf", "guess"):
        out["dev_" + k] = b[k]
    for k, v in _run_lm_set(b, "flat").items():
        out["dev_flat_" + k] = v
    # exp via forward differences (MINPACK lmdif)
    b = sims.make_lm_batch(8, seed0=5000, noise=0.1, dims=(32, 32))
    for k in ("images", "weights", "jac", "psf", "guess"):
        out["expfd_" + k] = b[k]
    for k, v in _run_lm_set(b, "none", analytic=False).items():
        out["expfd_none_" + k] = v
    # poor guesses: trial steps leave the model's domain (|g| >= 1, T < 0),
    # calc_fdiff returns -inf and MINPACK must reject the step
    b = sims.make_lm_batch(24, seed0=12000, noise=1.0, guess_frac=0.0,
                           dims=(32, 32))
    rng = np.random.RandomState(5)
    for i in range(24):
        guess = b["truth"][i].copy()
        guess[0:2] = rng.uniform(-0.3, 0.3, 2)
        guess[2:4] = rng.uniform(-0.65, 0.65, 2)
        guess[4] = rng.uniform(0.02, 3.0)
        guess[5] *= rng.uniform(0.1, 5)
        b["guess"][i] = guess
    for k in ("images", "weights", "jac", "psf", "guess"):
        out["bad_" + k] = b[k]
    for pk in ("none", "flat"):
        r = _run_lm_set(b, pk)
        print("bad", pk, "nfev", r["nfev"], "flags", r["flags"])
        for k, v in r.items():
            out["bad_%s_%s" % (pk, k)] = v
    save("lm_other", **out)


# ---------------------------------------------------------------- EM / admom
def _em_obs(image, weight, jrec):
    return ngmix.Observation(image, weight=weight,
                             jacobian=jac_from_record(jrec))


def gen_em_admom():
    from ngmix.em import EMFitter, run_em
    from ngmix.admom import AdmomFitter
    b = sims.make_psf_batch(64, seed0=7000)
    n = b["images"].shape[0]
    out = {k: b[k] for k in ("images", "weights", "jac", "em_guess",
                             "am_guess")}
    em = dict(flags=np.zeros(n, "i4"), numiter=np.zeros(n, "i4"),
              fdiff=np.zeros(n), sky=np.zeros(n), pars=np.zeros((n, 18)))
    fitter = EMFitter(tol=1e-5)
    for i in range(n):
        obs = _em_obs(b["images"][i], b["weights"][i], b["jac"][i])
        r = fitter.go(obs, ngmix.GMix(pars=b["em_guess"][i].ravel()))
        em["flags"][i] = r["flags"]
        if r["flags"] & ngmix.flags.EM_RANGE_ERROR == 0:
            em["numiter"][i] = r["numiter"]
            em["fdiff"][i] = r["fdiff"]
            em["sky"][i] = r["sky"]
            em["pars"][i] = r.get_gmix().get_full_pars()
    print("em numiter", em["numiter"][:16], "flags", em["flags"][:16])
    for k, v in em.items():
        out["em_" + k] = v

    # EM with a psf, the variants, vary_sky, and a zero-weight stamp
    var = {}
    psf = sims.simple_full_pars([0, 0, 0.0, 0.0, 0.05, 1.0], "gauss")
    for tag, kw in (("psf", {}), ("fixcen", dict(fixcen=True)),
                    ("fixcov", dict(fixcov=True)),
                    ("fluxonly", dict(fluxonly=True)),
                    ("varysky", dict(vary_sky=True)),
                    ("zerowt", {}), ("sky", dict(sky=0.0))):
        nv = 8
        numiter = np.zeros(nv, "i4")
        flags = np.zeros(nv, "i4")
        fdiff = np.zeros(nv)
        sky = np.zeros(nv)
        pars = np.zeros((nv, 12))
        for i in range(nv):
            image = b["images"][i]
            weight = b["weights"][i].copy()
            ignore = True
            if tag == "zerowt":
                weight[5:8, 6:9] = 0.0
                ignore = False
            psf_obs = ngmix.Observation(
                np.ones((5, 5)),
                jacobian=ngmix.DiagonalJacobian(row=2, col=2, scale=0.263),
                gmix=ngmix.GMix(pars=psf.ravel()))
            use_psf = tag in ("psf", "fixcen", "fixcov", "fluxonly")
            obs = ngmix.Observation(
                image if tag != "sky" else np.abs(image) + 1e-3,
                weight=weight, jacobian=jac_from_record(b["jac"][i]),
                psf=psf_obs if use_psf else None,
                ignore_zero_weight=ignore)
            guess = ngmix.GMix(pars=b["em_guess"][i][:2].ravel())
            r = run_em(obs, guess, **kw)
            flags[i] = r["flags"]
            if flags[i] & ngmix.flags.EM_RANGE_ERROR == 0:
                numiter[i] = r["numiter"]
                fdiff[i] = r["fdiff"]
                sky[i] = r["sky"]
                pars[i] = r.get_gmix().get_full_pars()
        var[tag + "_numiter"] = numiter
        var[tag + "_flags"] = flags
        var[tag + "_fdiff"] = fdiff
        var[tag + "_sky"] = sky
        var[tag + "_pars"] = pars
        print("em", tag, numiter, flags)
    var["psf"] = psf
    for k, v in var.items():
        out["emv_" + k] = v

    # admom
    am = dict(flags=np.zeros(n, "i4"), numiter=np.zeros(n, "i4"),
              npix=np.zeros(n, "i4"), wsum=np.zeros(n),
              sums=np.zeros((n, 7)), sums_cov=np.zeros((n, 7, 7)),
              pars=np.zeros((n, 6)), rho4=np.zeros(n),
              flux=np.zeros(n), flux_err=np.zeros(n), T=np.zeros(n),
              T_err=np.zeros(n), e=np.zeros((n, 2)), e_err=np.zeros((n, 2)),
              s2n=np.zeros(n), flags_final=np.zeros(n, "i4"))
    afit = AdmomFitter()
    for i in range(n):
        obs = _em_obs(b["images"][i], b["weights"][i], b["jac"][i])
        guess = ngmix.GMix(pars=b["am_guess"][i])
        r = afit.go(obs, guess)
        am["flags_final"][i] = r["flags"]
        for k in ("numiter", "npix", "wsum", "sums", "sums_cov", "pars",
                  "rho4", "flux", "flux_err", "T", "T_err", "e", "e_err",
                  "s2n"):
            am[k][i] = r[k]
    print("admom numiter", am["numiter"][:16], "flags", am["flags_final"][:16])
    for k, v in am.items():
        out["am_" + k] = v

    # admom failure modes: noise-only stamp, negative stamp, far-off guess
    rs = np.random.RandomState(5)
    fail_images = np.array([
        rs.normal(scale=1.0, size=(25, 25)),
        -np.abs(b["images"][0]),
        b["images"][1],
        b["images"][2],
    ])
    fail_guess = np.array([
        b["am_guess"][0], b["am_guess"][0],
        [1.0, 3.0, -3.0, 0.05, 0.0, 0.05],        # centre far off
        [1.0, 0.0, 0.0, 1e-3, 0.0, 1e-3],         # tiny weight
    ])
    ff = np.zeros(4, "i4")
    fn = np.zeros(4, "i4")
    for i in range(4):
        obs = _em_obs(fail_images[i], b["weights"][0], b["jac"][i])
        r = afit.go(obs, ngmix.GMix(pars=fail_guess[i]))
        ff[i] = r["flags"]
        fn[i] = r["numiter"]
    print("admom fail flags", ff, fn)
    out["amf_images"] = fail_images
    out["amf_guess"] = fail_guess
    out["amf_flags"] = ff
    out["amf_numiter"] = fn
    save("em_admom", **out)


def gen_lm_bdf():
    from ngmix import priors, joint_prior
    from ngmix.fitting import Fitter
    b = sims.make_bdf_batch(6, seed0=9000, dims=(48, 48), noise=0.02)
    rng = np.random.RandomState(7)
    prior = joint_prior.PriorBDFSep(
        cen_prior=priors.CenPrior(0.0, 0.0, 0.263, 0.263, rng=rng),
        g_prior=priors.GPriorBA(0.3, rng=rng),
        T_prior=priors.FlatPrior(-1.0, 1.0e3, rng=rng),
        fracdev_prior=priors.Normal(0.5, 0.1, rng=rng),
        F_prior=priors.FlatPrior(-100.0, 1.0e9, rng=rng))
    out = {k: b[k] for k in ("images", "weights", "jac", "psf", "guess")}
    n = b["images"].shape[0]
    res = dict(flags=np.zeros(n, "i4"), nfev=np.zeros(n, "i4"),
               ier=np.zeros(n, "i4"), pars=np.zeros((n, 7)),
               pars_err=np.zeros((n, 7)), pars_cov=np.zeros((n, 7, 7)),
               lnprob=np.zeros(n), s2n=np.zeros(n), chi2per=np.zeros(n))
    fitter = Fitter(model="bdf", prior=prior)
    for i in range(n):
        psf_obs = ngmix.Observation(
            np.ones((5, 5)),
            jacobian=ngmix.DiagonalJacobian(row=2, col=2, scale=0.263),
            gmix=ngmix.GMix(pars=b["psf"][i].ravel()))
        obs = ngmix.Observation(b["images"][i], weight=b["weights"][i],
                                jacobian=jac_from_record(b["jac"][i]),
                                psf=psf_obs)
        r = fitter.go(obs, b["guess"][i].copy())
        res["flags"][i] = r["flags"]
        res["nfev"][i] = r["nfev"]
        res["ier"][i] = r["ier"]
        res["pars"][i] = r["pars"]
        res["pars_err"][i] = r["pars_err"]
        res["pars_cov"][i] = r["pars_cov"]
        if r["flags"] == 0:
            res["lnprob"][i] = r["lnprob"]
            res["s2n"][i] = r["s2n"]
            res["chi2per"][i] = r["chi2per"]
    print("bdf flags", res["flags"], "nfev", res["nfev"], "ier", res["ier"])
    for k, v in res.items():
        out["res_" + k] = v
    save("lm_bdf", **out)


def gen_lm_bd():
    """bd model (8 parameters, lmdif) with and without PriorBDSep"""
    from ngmix import priors, joint_prior
    from ngmix.fitting import Fitter
    b = sims.make_bdf_batch(6, seed0=9100, dims=(40, 40), noise=0.02)
    n = b["images"].shape[0]
    rng = np.random.RandomState(8)
    guess = np.zeros((n, 8))
    guess[:, :5] = b["guess"][:, :5]
    guess[:, 5] = rng.uniform(-0.2, 0.2, n)          # log10(Td/Te)
    guess[:, 6:] = b["guess"][:, 5:]
    prior = joint_prior.PriorBDSep(
        cen_prior=priors.CenPrior(0.0, 0.0, 0.263, 0.263, rng=rng),
        g_prior=priors.GPriorBA(0.3, rng=rng),
        T_prior=priors.FlatPrior(-1.0, 1.0e3, rng=rng),
        logTratio_prior=priors.Normal(0.0, 0.5, rng=rng),
        fracdev_prior=priors.Normal(0.5, 0.1, rng=rng),
        F_prior=priors.FlatPrior(-100.0, 1.0e9, rng=rng))
    out = {k: b[k] for k in ("images", "weights", "jac", "psf")}
    out["guess"] = guess
    for pk, pr in (("none", None), ("bd", prior)):
        res = dict(flags=np.zeros(n, "i4"), nfev=np.zeros(n, "i4"),
                   ier=np.zeros(n, "i4"), pars=np.zeros((n, 8)),
                   pars_err=np.zeros((n, 8)), pars_cov=np.zeros((n, 8, 8)),
                   lnprob=np.zeros(n), s2n=np.zeros(n), chi2per=np.zeros(n),
                   dof=np.zeros(n, "i4"))
        fitter = Fitter(model="bd", prior=pr)
        for i in range(n):
            psf_obs = ngmix.Observation(
                np.ones((5, 5)),
                jacobian=ngmix.DiagonalJacobian(row=2, col=2, scale=0.263),
                gmix=ngmix.GMix(pars=b["psf"][i].ravel()))
            obs = ngmix.Observation(b["images"][i], weight=b["weights"][i],
                                    jacobian=jac_from_record(b["jac"][i]),
                                    psf=psf_obs)
            r = fitter.go(obs, guess[i].copy())
            res["flags"][i] = r["flags"]
            res["nfev"][i] = r["nfev"]
            res["ier"][i] = r["ier"]
            res["pars"][i] = r["pars"]
            if "pars_err" in r:
                res["pars_err"][i] = r["pars_err"]
                res["pars_cov"][i] = r["pars_cov"]
            if r["flags"] == 0:
                res["lnprob"][i] = r["lnprob"]
                res["s2n"][i] = r["s2n"]
                res["chi2per"][i] = r["chi2per"]
                res["dof"][i] = r["dof"]
        print("bd", pk, "flags", res["flags"], "nfev", res["nfev"], "ier",
              res["ier"])
        for k, v in res.items():
            out["%s_%s" % (pk, k)] = v
    save("lm_bd", **out)


def gen_noise_cov():
    """Fitter(use_noise_image=True): sandwich covariance under correlated
    noise (fitting/noise_cov.py); exp with the analytic derivative images,
    bdf with the central-difference ones"""
    from ngmix import priors, joint_prior
    from ngmix.fitting import Fitter
    out = {}
    for tag, model in (("exp", "exp"), ("bdf", "bdf")):
        if model == "exp":
            b = sims.make_lm_batch(5, seed0=15000, noise=0.05, dims=(32, 30))
            prior = None
        else:
            b = sims.make_bdf_batch(3, seed0=15100, dims=(36, 36), noise=0.02)
            rng = np.random.RandomState(7)
            prior = joint_prior.PriorBDFSep(
                cen_prior=priors.CenPrior(0.0, 0.0, 0.263, 0.263, rng=rng),
                g_prior=priors.GPriorBA(0.3, rng=rng),
                T_prior=priors.FlatPrior(-1.0, 1.0e3, rng=rng),
                fracdev_prior=priors.Normal(0.5, 0.1, rng=rng),
                F_prior=priors.FlatPrior(-100.0, 1.0e9, rng=rng))
        n = b["images"].shape[0]
        npars = b["guess"].shape[1]
        sig = b["noise"]
        rng = np.random.RandomState(77)
        # correlated noise: white noise smoothed by a small kernel, scaled
        # back to the nominal per-pixel variance; the same recipe for the
        # noise added to the image (on top of the sims' white noise it
        # replaces) and for the independent noise image
        def corr_noise(shape):
            w = rng.normal(size=shape)
            sm = (w + 0.6 * np.roll(w, 1, 0) + 0.6 * np.roll(w, 1, 1)
                  + 0.3 * np.roll(w, (1, 1), (0, 1)))
            return sm * (sig / sm.std())
        images = b["images"].copy()
        noise = np.zeros_like(images)
        for i in range(n):
            images[i] += corr_noise(images[i].shape)
            noise[i] = corr_noise(images[i].shape)
        for k in ("weights", "jac", "psf", "guess"):
            out["%s_%s" % (tag, k)] = b[k]
        out[tag + "_images"] = images
        out[tag + "_noise"] = noise
        res = dict(flags=np.zeros(n, "i4"), nfev=np.zeros(n, "i4"),
                   pars=np.zeros((n, npars)), pars_err=np.zeros((n, npars)),
                   pars_cov=np.zeros((n, npars, npars)),
                   pars_cov0=np.zeros((n, npars, npars)),
                   pars_err_plain=np.zeros((n, npars)))
        for i in range(n):
            psf_obs = ngmix.Observation(
                np.ones((5, 5)),
                jacobian=ngmix.DiagonalJacobian(row=2, col=2, scale=0.263),
                gmix=ngmix.GMix(pars=b["psf"][i].ravel()))
            obs = ngmix.Observation(images[i], weight=b["weights"][i],
                                    noise=noise[i],
                                    jacobian=jac_from_record(b["jac"][i]),
                                    psf=psf_obs)
            r = Fitter(model=model, prior=prior, use_noise_image=True).go(
                obs, b["guess"][i].copy())
            r0 = Fitter(model=model, prior=prior).go(obs, b["guess"][i].copy())
            res["flags"][i] = r["flags"]
            res["nfev"][i] = r["nfev"]
            res["pars"][i] = r["pars"]
            res["pars_err"][i] = r["pars_err"]
            res["pars_cov"][i] = r["pars_cov"]
            res["pars_cov0"][i] = r["pars_cov0"]
            res["pars_err_plain"][i] = r0["pars_err"]
        print(tag, "flags", res["flags"], "err ratio",
              res["pars_err"][:, 2] / res["pars_err_plain"][:, 2])
        for k, v in res.items():
            out["%s_res_%s" % (tag, k)] = v
    save("noise_cov", **out)


def gen_shape_moments():
    """known answers of the host-side shape / moment algebra
    (ngmix/shape.py, ngmix/moments.py)"""
    from ngmix import shape, moments
    rng = np.random.RandomState(21)
    g1 = rng.uniform(-0.6, 0.6, 12)
    g2 = rng.uniform(-0.6, 0.6, 12)
    s1 = rng.uniform(-0.1, 0.1, 12)
    s2 = rng.uniform(-0.1, 0.1, 12)
    T = rng.uniform(0.1, 4.0, 12)
    th = rng.uniform(-3, 3, 12)
    out = dict(g1=g1, g2=g2, s1=s1, s2=s2, T=T, theta=th)
    out["eta_from_g"] = np.array(shape.g1g2_to_eta1eta2(g1, g2))
    e1, e2 = np.array([shape.g1g2_to_e1e2(a, b) for a, b in zip(g1, g2)]).T
    out["eta_from_e"] = np.array(shape.e1e2_to_eta1eta2(e1, e2))
    out["g_from_eta"] = np.array(shape.eta1eta2_to_g1g2(
        out["eta_from_g"][0], out["eta_from_g"][1]))
    out["jacob"] = shape.dgs_by_dgo_jacob(g1, g2, s1, s2)
    out["rot"] = np.array(shape.rotate_shape(g1, g2, th))
    irr, irc, icc = moments.g2mom(g1[0], g2[0], T[0])
    out["g2mom"] = np.array([moments.g2mom(a, b, c)
                             for a, b, c in zip(g1, g2, T)])
    out["mom2g"] = np.array([moments.mom2g(*m) for m in out["g2mom"]])
    out["mom2e"] = np.array([moments.mom2e(*m) for m in out["g2mom"]])
    out["e2mom"] = np.array([moments.e2mom(a, b, c)
                             for a, b, c in zip(e1, e2, T)])
    out["sheared_g1g2T"] = np.array([
        moments.get_sheared_g1g2T(*v) for v in zip(g1, g2, T, s1, s2)])
    out["sheared_M1M2T"] = np.array([
        moments.get_sheared_M1M2T(c * a, c * b, c, d, e)
        for a, b, c, d, e in zip(e1, e2, T, s1, s2)])
    out["sheared_moments"] = np.array([
        moments.get_sheared_moments(m[0], m[1], m[2], d, e)
        for m, d, e in zip(out["g2mom"], s1, s2)])
    out["Tround"] = moments.get_Tround(T, g1, g2)
    out["T_from_round"] = moments.get_T(out["Tround"], g1, g2)
    out["r50"] = np.array([moments.T_to_r50(T), moments.r50_to_T(T),
                           moments.r50_to_sigma(T), moments.sigma_to_r50(T)])
    save("shape_moments", **out)


def gen_lm_priors2():
    """LogNormal and TruncatedGaussian priors: low-S/N exp fits where the
    priors shape the solution, and bdf fits with a truncated fracdev prior"""
    from ngmix import priors, joint_prior
    from ngmix.fitting import Fitter
    out = {}
    b = sims.make_lm_batch(10, seed0=17000, noise=1.0, dims=(32, 32))
    for k in ("images", "weights", "jac", "psf", "guess"):
        out["exp_" + k] = b[k]
    for pk in ("lognormal", "trunc"):
        r = _run_lm_set(b, pk)
        print(pk, "flags", r["flags"], "nfev", r["nfev"])
        for k, v in r.items():
            out["exp_%s_%s" % (pk, k)] = v
    b = sims.make_bdf_batch(4, seed0=17100, dims=(40, 40), noise=0.05)
    rng = np.random.RandomState(7)
    prior = joint_prior.PriorBDFSep(
        cen_prior=priors.CenPrior(0.0, 0.0, 0.263, 0.263, rng=rng),
        g_prior=priors.GPriorBA(0.3, rng=rng),
        T_prior=priors.LogNormal(0.5, 0.3, rng=rng),
        fracdev_prior=priors.TruncatedGaussian(0.5, 0.1, 0.0, 1.0, rng=rng),
        F_prior=priors.LogNormal(100.0, 50.0, rng=rng))
    for k in ("images", "weights", "jac", "psf", "guess"):
        out["bdf_" + k] = b[k]
    n = b["images"].shape[0]
    res = dict(flags=np.zeros(n, "i4"), nfev=np.zeros(n, "i4"),
               ier=np.zeros(n, "i4"), pars=np.zeros((n, 7)),
               pars_err=np.zeros((n, 7)), lnprob=np.zeros(n))
    fitter = Fitter(model="bdf", prior=prior)
    for i in range(n):
        psf_obs = ngmix.Observation(
            np.ones((5, 5)),
            jacobian=ngmix.DiagonalJacobian(row=2, col=2, scale=0.263),
            gmix=ngmix.GMix(pars=b["psf"][i].ravel()))
        obs = ngmix.Observation(b["images"][i], weight=b["weights"][i],
                                jacobian=jac_from_record(b["jac"][i]),
                                psf=psf_obs)
        r = fitter.go(obs, b["guess"][i].copy())
        res["flags"][i] = r["flags"]
        res["nfev"][i] = r["nfev"]
        res["ier"][i] = r["ier"]
        res["pars"][i] = r["pars"]
        if r["flags"] == 0:
            res["pars_err"][i] = r["pars_err"]
            res["lnprob"][i] = r["lnprob"]
    print("bdf flags", res["flags"], "nfev", res["nfev"])
    for k, v in res.items():
        out["bdf_res_" + k] = v
    save("lm_priors2", **out)


def gen_lm_coellip():
    """CoellipFitter.go (forward-difference LM of n coelliptical gaussians)
    on PSF stamps, ngauss 1-4, with and without PriorCoellipSame; guesses from
    the reference's CoellipPSFGuesser"""
    from ngmix import priors, joint_prior
    from ngmix.fitting import CoellipFitter
    from ngmix.guessers import CoellipPSFGuesser
    b = sims.make_psf_batch(6, seed0=19000)
    n = b["images"].shape[0]
    out = {k: b[k] for k in ("images", "weights", "jac")}
    for ngauss in (1, 2, 3, 4, 5):
        for pk in ("none", "prior"):
            rng = np.random.RandomState(100 * ngauss + (pk == "prior"))
            prior = None
            if pk == "prior":
                prior = joint_prior.PriorCoellipSame(
                    ngauss,
                    priors.CenPrior(0.0, 0.0, 0.263, 0.263, rng=rng),
                    priors.GPriorBA(0.3, rng=rng),
                    priors.TwoSidedErf(-1.0, 0.1, 1.0e3, 1.0, rng=rng),
                    priors.TwoSidedErf(-1.0e2, 1.0, 1.0e9, 1.0e8, rng=rng))
            np_ = 4 + 2 * ngauss
            res = dict(flags=np.zeros(n, "i4"), nfev=np.zeros(n, "i4"),
                       ier=np.zeros(n, "i4"), pars=np.zeros((n, np_)),
                       pars_err=np.zeros((n, np_)), guess=np.zeros((n, np_)),
                       pars_cov=np.zeros((n, np_, np_)),
                       lnprob=np.zeros(n), s2n=np.zeros(n),
                       chi2per=np.zeros(n), dof=np.zeros(n, "i4"),
                       npix=np.zeros(n, "i4"))
            guesser = CoellipPSFGuesser(rng=rng, ngauss=ngauss)
            fitter = CoellipFitter(ngauss=ngauss, prior=prior)
            for i in range(n):
                obs = ngmix.Observation(
                    b["images"][i], weight=b["weights"][i],
                    jacobian=jac_from_record(b["jac"][i]))
                guess = guesser(obs=obs)
                r = fitter.go(obs, guess.copy())
                res["guess"][i] = guess
                res["flags"][i] = r["flags"]
                res["nfev"][i] = r["nfev"]
                res["ier"][i] = r["ier"]
                res["pars"][i] = r["pars"]
                if "pars_err" in r:
                    res["pars_err"][i] = r["pars_err"]
                    res["pars_cov"][i] = r["pars_cov"]
                if r["flags"] == 0:
                    for k in ("lnprob", "s2n", "chi2per", "dof", "npix"):
                        res[k][i] = r[k]
            print("coellip", ngauss, pk, "flags", res["flags"], "nfev",
                  res["nfev"], "ier", res["ier"])
            for k, v in res.items():
                out["n%d_%s_%s" % (ngauss, pk, k)] = v
    save("lm_coellip", **out)


def gen_gaussap():
    """gaussian-aperture fluxes (GMix.get_gaussap_flux, ngmix.gaussap)"""
    from ngmix import gaussap
    rng = np.random.RandomState(41)
    out = {}
    for model, npars in (("exp", 7), ("dev", 6), ("bdf", 8), ("gauss", 6)):
        n = 6
        pars = np.zeros((n, npars))
        pars[:, 0:2] = rng.uniform(-0.1, 0.1, (n, 2))
        pars[:, 2:4] = rng.uniform(-0.4, 0.4, (n, 2))
        pars[:, 4] = rng.uniform(-0.1, 2.0, n)
        k = 5
        if model == "bdf":
            pars[:, 5] = rng.uniform(0.0, 1.0, n)
            k = 6
        pars[:, k:] = rng.uniform(1.0, 100.0, (n, npars - k))
        pars[1, 2] = 0.99
        pars[1, 3] = 0.5             # |g| > 1: range error
        mask = np.ones(n, dtype=bool)
        mask[2] = False
        flux, flags = gaussap.get_gaussap_flux(pars, model, 2.5, mask=mask,
                                               verbose=False)
        out[model + "_pars"], out[model + "_mask"] = pars, mask
        out[model + "_flux"], out[model + "_flags"] = flux, flags
    n = 4
    pars = np.zeros((n, 6))
    pars[:, 2:4] = rng.uniform(-0.3, 0.3, (n, 2))
    pars[:, 4] = rng.uniform(0.1, 2.0, n)
    pars[:, 5] = rng.uniform(1.0, 100.0, n)
    fd, td = rng.uniform(0, 1, n), rng.uniform(0.5, 2.0, n)
    flux, flags = gaussap.get_gaussap_flux(pars, "cm", 2.0, fracdev=fd,
                                           TdByTe=td, verbose=False)
    out["cm_pars"], out["cm_fracdev"], out["cm_TdByTe"] = pars, fd, td
    out["cm_flux"], out["cm_flags"] = flux, flags
    save("gaussap", **out)


def gen_moments():
    """GaussMom (6 and 17 moments) on galaxy stamps, and get_weighted_sums
    with a two-gaussian off-centre weight and a maxrad"""
    from ngmix.gaussmom import GaussMom
    b = sims.make_lm_batch(12, seed0=9100, noise=0.05)
    n = b["images"].shape[0]
    # a masked stamp and a low-S/N stamp
    rng = np.random.RandomState(91)
    b["weights"][3][rng.uniform(size=b["weights"][3].shape) < 0.15] = 0.0
    b["images"][5] = rng.normal(0.0, 0.05, size=b["images"][5].shape)
    out = {k: b[k] for k in ("images", "weights", "jac")}
    out["fwhm"] = np.array(1.2)
    stat_keys = ("flags", "flux_flags", "T_flags", "flux", "flux_err", "T",
                 "T_err", "s2n", "e1", "e2", "sums_norm", "wsum")
    for nmom, tag in ((6, "m6"), (17, "m17")):
        fitter = GaussMom(fwhm=1.2, with_higher_order=(nmom == 17))
        res = {k: np.zeros(n) for k in stat_keys}
        res["npix"] = np.zeros(n, "i4")
        res["sums"] = np.zeros((n, nmom))
        res["sums_cov"] = np.zeros((n, nmom, nmom))
        res["sums_err"] = np.zeros((n, 6))
        res["e_err"] = np.zeros((n, 2))
        res["pars"] = np.zeros((n, 6)) + np.nan
        for i in range(n):
            obs = ngmix.Observation(b["images"][i], weight=b["weights"][i],
                                    jacobian=jac_from_record(b["jac"][i]))
            r = fitter.go(obs)
            for k in stat_keys:
                res[k][i] = r[k]
            res["npix"][i] = r["npix"]
            res["sums"][i] = r["sums"]
            res["sums_cov"][i] = r["sums_cov"]
            res["sums_err"][i] = r["sums_err"][:6]
            res["e_err"][i] = r["e_err"]
            if "pars" in r and r["flags"] == 0:
                res["pars"][i] = r["pars"][:6]
        print(tag, "flags", res["flags"], "T", res["T"][:4])
        for k, v in res.items():
            out[tag + "_" + k] = v
    # raw sums with a general weight mixture and a radius cut
    wt_pars = np.array([0.6, 0.05, -0.03, 0.20, 0.02, 0.25,
                        0.4, -0.04, 0.06, 0.45, -0.05, 0.38])
    wt = ngmix.GMix(pars=wt_pars)
    out["wt_pars"] = wt_pars
    out["maxrad"] = np.array(2.5)
    for nmom, tag in ((6, "w6"), (17, "w17")):
        sums = np.zeros((n, nmom))
        cov = np.zeros((n, nmom, nmom))
        wsum = np.zeros(n)
        npix = np.zeros(n, "i4")
        for i in range(n):
            obs = ngmix.Observation(b["images"][i], weight=b["weights"][i],
                                    jacobian=jac_from_record(b["jac"][i]))
            r = wt.get_weighted_sums(obs, maxrad=2.5,
                                     with_higher_order=(nmom == 17))
            sums[i], cov[i] = r["sums"], r["sums_cov"]
            wsum[i], npix[i] = r["wsum"], r["npix"]
        out[tag + "_sums"], out[tag + "_sums_cov"] = sums, cov
        out[tag + "_wsum"], out[tag + "_npix"] = wsum, npix
    save("moments", **out)


def _multi_case(rng, nband, nepoch, model, noise, with_prior):
    """one object observed nepoch[b] times in band b: stamps, truth, guess"""
    from ngmix_b200.sims import (diag_jacobian, simple_full_pars,
                                 convolve_full_pars, render_full_pars)
    H = W = 32
    g1, g2 = rng.normal(0.0, 0.12, size=2)
    T = rng.uniform(0.2, 0.8)
    fluxes = rng.uniform(60.0, 180.0, size=nband)
    images, weights, jacs, psfs, bands = [], [], [], [], []
    for b in range(nband):
        for _ in range(nepoch[b]):
            dy, dx = rng.uniform(-0.5, 0.5, size=2)
            jac = diag_jacobian((H - 1) / 2 + dy, (W - 1) / 2 + dx, 0.263)
            ppars = [0.0, 0.0, rng.normal(0, 0.01), rng.normal(0, 0.01),
                     rng.uniform(0.22, 0.32), 1.0]
            psf = simple_full_pars(ppars, "turb")
            pars = [0.0, 0.0, g1, g2, T, fluxes[b]]
            full = convolve_full_pars(simple_full_pars(pars, model), psf)
            im = render_full_pars(full, (H, W), jac)
            images.append(im + rng.normal(0.0, noise, size=im.shape))
            wt = np.zeros((H, W)) + 1.0 / noise ** 2
            if rng.uniform() < 0.3:
                wt[rng.uniform(size=wt.shape) < 0.05] = 0.0
            weights.append(wt)
            jacs.append(jac)
            psfs.append(psf)
            bands.append(b)
    truth = np.array([0.0, 0.0, g1, g2, T] + list(fluxes))
    guess = truth.copy()
    guess[0:2] = rng.uniform(-0.01, 0.01, size=2)
    guess[2:] *= 1.0 + rng.uniform(-0.05, 0.05, size=guess.size - 2)
    return (np.array(images), np.array(weights), np.array(jacs),
            np.array(psfs), np.array(bands, dtype="i4"), truth, guess)


def gen_lm_multi():
    """Fitter.go on ObsList and MultiBandObsList inputs (multi-epoch,
    1-3 bands), with and without PriorSimpleSep (a flux prior per band)"""
    configs = [
        # (nband, epochs per band, model, with_prior)
        (1, [3], "exp", False), (1, [3], "exp", True),
        (2, [2, 2], "exp", True), (2, [1, 3], "gauss", False),
        (3, [1, 1, 1], "exp", True), (3, [2, 1, 2], "dev", True),
        (3, [1, 2, 1], "gauss", False), (1, [4], "dev", False),
    ]
    _gen_multi(configs, 31415, "lm_multi")


def gen_lm_multi_wide():
    """the same with 4 to 7 bands (9 to 12 parameters)"""
    configs = [
        (4, [1, 1, 1, 1], "exp", True), (5, [1, 2, 1, 1, 1], "exp", False),
        (5, [1, 1, 1, 1, 1], "gauss", True), (6, [1, 1, 1, 2, 1, 1], "dev", True),
        (7, [1, 1, 1, 1, 1, 1, 1], "exp", True),
    ]
    _gen_multi(configs, 27182, "lm_multi_wide")


def gen_lm_multi_xwide():
    """the same with 8 to 12 bands (13 to 17 parameters): the reference has
    no limit on the number of bands (results.py:212-260)"""
    configs = [
        (8, [1] * 8, "exp", True), (9, [1, 2, 1, 1, 1, 1, 1, 1, 1], "gauss", False),
        (10, [1] * 10, "exp", True), (12, [1] * 12, "dev", True),
    ]
    _gen_multi(configs, 16180, "lm_multi_xwide")


def _gen_multi(configs, seed, name):
    from ngmix.fitting import Fitter
    from ngmix import priors, joint_prior
    rng = np.random.RandomState(seed)
    out = {"ncases": np.array(0)}
    icase = 0
    for nband, nep, model, with_prior in configs:
        for rep in range(3):
            noise = [0.02, 0.05, 0.3][rep]
            (images, weights, jacs, psfs, bands, truth,
             guess) = _multi_case(rng, nband, nep, model, noise, with_prior)
            mbobs = ngmix.MultiBandObsList()
            k = 0
            for b in range(nband):
                ol = ngmix.ObsList()
                for _ in range(nep[b]):
                    psf_obs = ngmix.Observation(
                        np.zeros((5, 5)) + 1.0,
                        gmix=ngmix.GMix(pars=psfs[k].ravel()))
                    ol.append(ngmix.Observation(
                        images[k], weight=weights[k],
                        jacobian=jac_from_record(jacs[k]), psf=psf_obs))
                    k += 1
                mbobs.append(ol)
            prior = None
            if with_prior:
                prior = joint_prior.PriorSimpleSep(
                    priors.CenPrior(0.0, 0.0, 0.263, 0.263, rng=rng),
                    priors.GPriorBA(0.3, rng=rng),
                    priors.TwoSidedErf(-1.0, 0.1, 1.0e3, 1.0, rng=rng),
                    [priors.TwoSidedErf(-1.0e2, 1.0, 1.0e9, 1.0e8, rng=rng)
                     for _ in range(nband)])
            fitter = Fitter(model=model, prior=prior)
            obs_in = mbobs if nband > 1 else mbobs[0]
            r = fitter.go(obs=obs_in, guess=guess)
            np_ = 5 + nband
            p = "c%d_" % icase
            out[p + "images"], out[p + "weights"] = images, weights
            out[p + "jac"], out[p + "psf"], out[p + "bands"] = jacs, psfs, bands
            out[p + "guess"] = guess
            out[p + "model"] = np.array(model)
            out[p + "nband"] = np.array(nband)
            out[p + "prior"] = np.array(int(with_prior))
            out[p + "flags"] = np.array(r["flags"])
            out[p + "nfev"] = np.array(r["nfev"])
            out[p + "ier"] = np.array(r["ier"])
            out[p + "pars"] = np.array(r["pars"])
            out[p + "pars_cov"] = np.array(r["pars_cov"])
            out[p + "pars_err"] = np.array(r["pars_err"])
            if r["flags"] == 0:
                for key in ("lnprob", "s2n", "chi2per", "dof", "npix",
                            "s2n_numer", "s2n_denom", "T", "T_err"):
                    out[p + key] = np.array(r[key])
                out[p + "flux"] = np.atleast_1d(r["flux"])
                out[p + "flux_err"] = np.atleast_1d(r["flux_err"])
            print(icase, model, nband, nep, "flags", r["flags"], "nfev",
                  r["nfev"], "ier", r["ier"])
            assert len(r["pars"]) == np_
            icase += 1
    out["ncases"] = np.array(icase)
    save(name, **out)


def gen_psfflux():
    """PSFFluxFitter.go (template flux by cross-correlation) on single
    observations and ObsLists"""
    from ngmix.fitting import PSFFluxFitter
    b = sims.make_lm_batch(12, seed0=9400, noise=0.05)
    rng = np.random.RandomState(94)
    b["weights"][2][rng.uniform(size=b["weights"][2].shape) < 0.2] = 0.0
    out = {k: b[k] for k in ("images", "weights", "jac", "psf")}
    groups = [[0], [1], [2], [3, 4, 5], [6, 7], [8, 9, 10, 11]]
    out["ngroups"] = np.array(len(groups))
    for normalize in (True, False):
        tag = "n%d_" % int(normalize)
        res = {k: np.zeros(len(groups)) for k in
               ("flags", "flux", "flux_err", "chi2per", "dof")}
        for ig, grp in enumerate(groups):
            ol = ngmix.ObsList()
            for i in grp:
                ppars = b["psf"][i].copy()
                if not normalize:
                    ppars[:, 0] *= 1.7      # unnormalised psf mixture
                psf_obs = ngmix.Observation(
                    np.zeros((5, 5)) + 1.0, gmix=ngmix.GMix(pars=ppars.ravel()))
                ol.append(ngmix.Observation(
                    b["images"][i], weight=b["weights"][i],
                    jacobian=jac_from_record(b["jac"][i]), psf=psf_obs))
            obs_in = ol if len(grp) > 1 else ol[0]
            r = PSFFluxFitter(normalize_psf=normalize).go(obs=obs_in)
            for k in res:
                res[k][ig] = r[k]
        for k, v in res.items():
            out[tag + k] = v
        print(tag, res["flux"], res["flags"])
    for ig, grp in enumerate(groups):
        out["g%d" % ig] = np.array(grp)
    save("psfflux", **out)


def _boot_obs(i, b, p):
    psf_obs = ngmix.Observation(
        p["images"][i], weight=p["weights"][i],
        jacobian=jac_from_record(p["jac"][i]))
    return ngmix.Observation(
        b["images"][i], weight=b["weights"][i],
        jacobian=jac_from_record(b["jac"][i]), psf=psf_obs)


def gen_boot():
    """guessers with fixed seeds, and the PSF-EM -> psf-flux guess -> LM
    chain through PSFRunner / Runner / Bootstrapper"""
    from ngmix import guessers
    from ngmix.em import EMFitter
    from ngmix.fitting import Fitter
    from ngmix.runners import Runner, PSFRunner
    from ngmix.bootstrap import Bootstrapper
    n = 8
    b = sims.make_lm_batch(n, seed0=9600, noise=0.03)
    # psf images rendered from the same psf mixtures as the galaxy stamps
    p = dict(images=np.zeros((n, 25, 25)), weights=np.zeros((n, 25, 25)),
             jac=np.zeros((n, 8)))
    for i in range(n):
        rng = np.random.RandomState(9700 + i)
        p["jac"][i] = sims.diag_jacobian(12.0 + rng.uniform(-0.3, 0.3),
                                         12.0 + rng.uniform(-0.3, 0.3), 0.263)
        im = sims.render_full_pars(b["psf"][i], (25, 25), p["jac"][i])
        p["images"][i] = im + rng.normal(0.0, 1.0e-4, size=im.shape)
        p["weights"][i] = 1.0 / 1.0e-4 ** 2
    out = {"images": b["images"], "weights": b["weights"], "jac": b["jac"],
           "psf_images": p["images"], "psf_weights": p["weights"],
           "psf_jac": p["jac"], "n": np.array(n)}
    # ---- stand-alone guessers, fixed seeds
    obs0 = _boot_obs(0, b, p)
    for ng in (1, 2, 3, 4, 5):
        for moms in (0, 1):
            g = guessers.GMixPSFGuesser(np.random.RandomState(100 + ng), ng,
                                        guess_from_moms=bool(moms))
            out["gpsf_%d_%d" % (ng, moms)] = g(obs0.psf).get_full_pars()
    out["gsimple"] = guessers.SimplePSFGuesser(
        np.random.RandomState(7))(obs0.psf)
    out["gsimple_moms"] = guessers.SimplePSFGuesser(
        np.random.RandomState(7), guess_from_moms=True)(obs0.psf)
    out["gtflux1"] = guessers.TFluxGuesser(
        np.random.RandomState(8), 0.4, 100.0)()
    out["gtflux3"] = guessers.TFluxGuesser(
        np.random.RandomState(8), 0.4, [100.0, 50.0, 20.0])(nrand=3)
    out["gpars"] = guessers.ParsGuesser(
        np.random.RandomState(9), [0.1, -0.1, 0.2, 0.7, 0.5, 80.0])()
    out["gpars4"] = guessers.ParsGuesser(
        np.random.RandomState(9), [0.1, -0.1, 0.2, 0.85, 0.5, 80.0],
        widths=[0.01, 0.01, 0.05, 0.05, 0.2, 0.2])(nrand=4)
    # ---- the chain
    res = dict(flags=np.zeros(n, "i4"), nfev=np.zeros(n, "i4"),
               pars=np.zeros((n, 6)), pars_err=np.zeros((n, 6)),
               psf_flags=np.zeros(n, "i4"), psf_numiter=np.zeros(n, "i4"),
               psf_pars=np.zeros((n, 12)), psf_flux=np.zeros(n),
               s2n=np.zeros(n))
    for i in range(n):
        rng = np.random.RandomState(9800 + i)
        obs = _boot_obs(i, b, p)
        psf_runner = PSFRunner(
            fitter=EMFitter(tol=1.0e-5),
            guesser=guessers.GMixPSFGuesser(rng=rng, ngauss=2), ntry=2)
        runner = Runner(
            fitter=Fitter(model="exp"),
            guesser=guessers.TPSFFluxGuesser(rng=rng, T=0.25), ntry=2)
        r = Bootstrapper(runner=runner, psf_runner=psf_runner).go(obs)
        pres = obs.psf.meta["result"]
        res["flags"][i], res["nfev"][i] = r["flags"], r["nfev"]
        res["pars"][i], res["pars_err"][i] = r["pars"], r["pars_err"]
        res["s2n"][i] = r["s2n"]
        res["psf_flags"][i] = pres["flags"]
        res["psf_numiter"][i] = pres["numiter"]
        res["psf_pars"][i] = obs.psf.gmix.get_full_pars()
        res["psf_flux"][i] = ngmix.fitting.PSFFluxFitter().go(obs)["flux"]
    print("boot flags", res["flags"], "nfev", res["nfev"], "psf numiter",
          res["psf_numiter"])
    for k, v in res.items():
        out["boot_" + k] = v
    save("boot", **out)


def gen_guessers():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from guesser_cases import guesser_cases
    out = guesser_cases(ngmix, lambda seed: np.random.RandomState(seed))
    for k, v in out.items():
        print(k, np.asarray(v).shape)
    save("guessers", **{k: np.asarray(v) for k, v in out.items()})


class _OneBasedMinpack(object):
    """scipy >= 1.15 returns a 0-based ipvt from _lmdif/_lmder; the bounded
    branch of the reference (leastsqbound.py:533-540) still subtracts 1
    (Fortran convention), which scrambles its covariance under this scipy
    (SURVEY.md section 7).  This shim hands the reference the 1-based ipvt
    its code assumes, WITHOUT touching the reference: the fit itself (x, nfev,
    ier) is unaffected, only cov_x changes - to the intended
    (J_ext^T J_ext)^-1."""

    def __init__(self, real):
        self._real = real

    def _fix(self, r):
        r[1]["ipvt"] = r[1]["ipvt"] + 1
        return r

    def _lmdif(self, *a):
        return self._fix(self._real._lmdif(*a))

    def _lmder(self, *a):
        return self._fix(self._real._lmder(*a))

    def __getattr__(self, k):
        return getattr(self._real, k)


def _bounded_fits(fitter, b, npars):
    """run the fitter over the batch twice: as the reference runs here, and
    with the 1-based-ipvt shim (intended covariance)"""
    import ngmix.fitting.leastsqbound as lsb
    n = b["images"].shape[0]
    res = dict(flags=np.zeros(n, "i4"), nfev=np.zeros(n, "i4"),
               ier=np.zeros(n, "i4"), pars=np.zeros((n, npars)),
               pars_err=np.zeros((n, npars)),
               pars_cov0=np.zeros((n, npars, npars)),
               pars_cov=np.zeros((n, npars, npars)),
               pars_cov0_asrun=np.zeros((n, npars, npars)),
               flags_asrun=np.zeros(n, "i4"),
               lnprob=np.full(n, np.nan), s2n=np.full(n, np.nan),
               chi2per=np.full(n, np.nan))
    real = lsb._minpack
    for shim in (False, True):
        lsb._minpack = _OneBasedMinpack(real) if shim else real
        try:
            for i in range(n):
                psf_obs = ngmix.Observation(
                    np.ones((5, 5)),
                    jacobian=ngmix.DiagonalJacobian(row=2, col=2, scale=0.263),
                    gmix=ngmix.GMix(pars=b["psf"][i].ravel()))
                obs = ngmix.Observation(
                    b["images"][i], weight=b["weights"][i],
                    jacobian=jac_from_record(b["jac"][i]), psf=psf_obs)
                r = fitter.go(obs, b["guess"][i].copy())
                c0 = r["pars_cov0"]
                c0 = np.nan if c0 is None else c0
                if not shim:
                    res["pars_cov0_asrun"][i] = c0
                    res["flags_asrun"][i] = r["flags"]
                    res["nfev"][i] = r["nfev"]
                    res["ier"][i] = r["ier"]
                    res["pars"][i] = r["pars"]
                else:
                    assert r["nfev"] == res["nfev"][i]
                    assert np.array_equal(r["pars"], res["pars"][i])
                    res["flags"][i] = r["flags"]
                    res["pars_err"][i] = r["pars_err"]
                    res["pars_cov0"][i] = c0
                    res["pars_cov"][i] = r["pars_cov"]
                    if r["flags"] == 0:
                        res["lnprob"][i] = r["lnprob"]
                        res["s2n"][i] = r["s2n"]
                        res["chi2per"][i] = r["chi2per"]
        finally:
            lsb._minpack = real
    return res


def gen_lm_bounds():
    """bounded fits (prior members with `bounds`, leastsqbound.py:183-263,
    445-552): bdf with fracdev in (0, 1) through lmdif, as
    tests/test_leastsqbound.py:114-176 / tests/_priors.py:24-35, and exp with
    bounded T and flux through lmder with the analytic jacobian, as
    tests/test_fitting_lm_jacobian.py:216-250"""
    from ngmix import priors, joint_prior
    from ngmix.fitting import Fitter
    out = {}
    # ---- bdf, fracdev bounded, forward differences
    b = sims.make_bdf_batch(8, seed0=9100, dims=(48, 48), noise=0.02)
    rng = np.random.RandomState(7)
    prior = joint_prior.PriorBDFSep(
        cen_prior=priors.CenPrior(0.0, 0.0, 0.263, 0.263, rng=rng),
        g_prior=priors.GPriorBA(0.3, rng=rng),
        T_prior=priors.FlatPrior(-1.0, 1.0e3, rng=rng),
        fracdev_prior=priors.Normal(0.5, 0.3, rng=rng, bounds=[0.0, 1.0]),
        F_prior=priors.FlatPrior(-100.0, 1.0e9, rng=rng))
    assert prior.bounds is not None
    res = _bounded_fits(Fitter(model="bdf", prior=prior), b, 7)
    print("bounded bdf flags", res["flags"], "asrun", res["flags_asrun"],
          "nfev", res["nfev"], "ier", res["ier"])
    for k in ("images", "weights", "jac", "psf", "guess"):
        out["bdf_" + k] = b[k]
    for k, v in res.items():
        out["bdf_res_" + k] = v
    # ---- exp, T and flux bounded, analytic jacobian
    b = sims.make_lm_batch(8, seed0=9200, dims=(40, 40), noise=0.05)
    rng = np.random.RandomState(8)
    prior = joint_prior.PriorSimpleSep(
        cen_prior=priors.CenPrior(0.0, 0.0, 0.263, 0.263, rng=rng),
        g_prior=priors.GPriorBA(0.3, rng=rng),
        T_prior=priors.Normal(0.4, 1.0, bounds=[0.01, 100.0], rng=rng),
        F_prior=priors.Normal(100.0, 100.0, bounds=[1.0, 1.0e9], rng=rng))
    assert prior.bounds is not None
    for tag, analytic in (("an", True), ("fd", False)):
        res = _bounded_fits(
            Fitter(model="exp", prior=prior, analytic_jacobian=analytic),
            b, 6)
        print("bounded exp", tag, "flags", res["flags"], "asrun",
              res["flags_asrun"], "nfev", res["nfev"], "ier", res["ier"])
        for k, v in res.items():
            out["exp_%s_res_%s" % (tag, k)] = v
    for k in ("images", "weights", "jac", "psf", "guess"):
        out["exp_" + k] = b[k]
    # ---- exp with tight two-sided bounds and a one-sided lower bound, worse
    #      guesses: the transforms take part in many iterations
    b = sims.make_lm_batch(8, seed0=9300, dims=(40, 40), noise=0.5,
                           guess_frac=0.3)
    rng = np.random.RandomState(9)
    prior = joint_prior.PriorSimpleSep(
        cen_prior=priors.CenPrior(0.0, 0.0, 0.263, 0.263, rng=rng),
        g_prior=priors.GPriorBA(0.3, rng=rng),
        T_prior=priors.Normal(0.5, 1.0, bounds=[0.05, 5.0], rng=rng),
        F_prior=priors.Normal(100.0, 100.0, bounds=[1.0, None], rng=rng))
    assert prior.bounds is not None
    for tag, analytic in (("an", True), ("fd", False)):
        res = _bounded_fits(
            Fitter(model="exp", prior=prior, analytic_jacobian=analytic),
            b, 6)
        print("bounded exp2", tag, "flags", res["flags"], "asrun",
              res["flags_asrun"], "nfev", res["nfev"], "ier", res["ier"])
        for k, v in res.items():
            out["exp2_%s_res_%s" % (tag, k)] = v
    for k in ("images", "weights", "jac", "psf", "guess"):
        out["exp2_" + k] = b[k]
    save("lm_bounds", **out)


def gen_lm_multi_fd():
    """Fitter.go on ObsList / MultiBandObsList inputs through lmdif: bdf and
    bd objects (PriorBDFSep / PriorBDSep with a flux prior per band,
    joint_prior.py:267-676) and exp objects with analytic_jacobian=False"""
    from ngmix.fitting import Fitter
    from ngmix import priors, joint_prior
    from ngmix_b200.sims import (diag_jacobian, simple_full_pars,
                                 bdf_full_pars, convolve_full_pars,
                                 render_full_pars)
    rng = np.random.RandomState(16180)
    configs = [
        # (nband, epochs per band, model, with_prior)
        (1, [2], "bdf", True), (2, [1, 1], "bdf", True),
        (3, [1, 2, 1], "bdf", True), (2, [2, 1], "bdf", False),
        (5, [1, 1, 1, 1, 1], "bdf", True),
        (2, [1, 1], "bd", True), (1, [3], "bd", True),
        (2, [1, 2], "exp", True), (3, [1, 1, 1], "exp", False),
        # 8 and 9 bands: 14 and 16 parameters (ngx_lm_fit_multi_xwide)
        (8, [1] * 8, "bdf", True), (9, [1] * 9, "bd", True),
    ]
    out = {}
    icase = 0
    H = W = 32
    for nband, nep, model, with_prior in configs:
        for rep in range(2):
            noise = [0.02, 0.1][rep]
            g1, g2 = rng.normal(0.0, 0.12, size=2)
            T = rng.uniform(0.2, 0.8)
            fracdev = rng.uniform(0.2, 0.8)
            fluxes = rng.uniform(60.0, 180.0, size=nband)
            images, weights, jacs, psfs, bands = [], [], [], [], []
            for b in range(nband):
                for _ in range(nep[b]):
                    dy, dx = rng.uniform(-0.5, 0.5, size=2)
                    jac = diag_jacobian((H - 1) / 2 + dy, (W - 1) / 2 + dx,
                                        0.263)
                    ppars = [0.0, 0.0, rng.normal(0, 0.01),
                             rng.normal(0, 0.01), rng.uniform(0.22, 0.32), 1.0]
                    psf = simple_full_pars(ppars, "turb")
                    if model == "exp":
                        obj = simple_full_pars(
                            [0.0, 0.0, g1, g2, T, fluxes[b]], "exp")
                    else:
                        # bd truth with logTratio = 0 is the bdf mixture
                        obj = bdf_full_pars(
                            [0.0, 0.0, g1, g2, T, fracdev, fluxes[b]])
                    im = render_full_pars(convolve_full_pars(obj, psf),
                                          (H, W), jac)
                    images.append(im + rng.normal(0.0, noise, size=im.shape))
                    wt = np.zeros((H, W)) + 1.0 / noise ** 2
                    if rng.uniform() < 0.3:
                        wt[rng.uniform(size=wt.shape) < 0.05] = 0.0
                    weights.append(wt)
                    jacs.append(jac)
                    psfs.append(psf)
                    bands.append(b)
            shared = {"exp": [0.0, 0.0, g1, g2, T],
                      "bdf": [0.0, 0.0, g1, g2, T, fracdev],
                      "bd": [0.0, 0.0, g1, g2, T, 0.0, fracdev]}[model]
            truth = np.array(shared + list(fluxes))
            guess = truth.copy()
            guess[0:2] = rng.uniform(-0.01, 0.01, size=2)
            guess[2:] *= 1.0 + rng.uniform(-0.05, 0.05, size=guess.size - 2)
            if model == "bd":
                guess[5] = rng.uniform(-0.05, 0.05)
            mbobs = ngmix.MultiBandObsList()
            k = 0
            for b in range(nband):
                ol = ngmix.ObsList()
                for _ in range(nep[b]):
                    psf_obs = ngmix.Observation(
                        np.zeros((5, 5)) + 1.0,
                        gmix=ngmix.GMix(pars=psfs[k].ravel()))
                    ol.append(ngmix.Observation(
                        images[k], weight=weights[k],
                        jacobian=jac_from_record(jacs[k]), psf=psf_obs))
                    k += 1
                mbobs.append(ol)
            prior = None
            if with_prior:
                cen = priors.CenPrior(0.0, 0.0, 0.263, 0.263, rng=rng)
                gp = priors.GPriorBA(0.3, rng=rng)
                Tp = priors.FlatPrior(-1.0, 1.0e3, rng=rng)
                Fp = [priors.FlatPrior(-100.0, 1.0e9, rng=rng)
                      for _ in range(nband)]
                if model == "bdf":
                    prior = joint_prior.PriorBDFSep(
                        cen, gp, Tp, priors.Normal(0.5, 0.1, rng=rng), Fp)
                elif model == "bd":
                    prior = joint_prior.PriorBDSep(
                        cen, gp, Tp, priors.Normal(0.0, 0.5, rng=rng),
                        priors.Normal(0.5, 0.1, rng=rng), Fp)
                else:
                    prior = joint_prior.PriorSimpleSep(cen, gp, Tp, Fp)
            fitter = Fitter(model=model, prior=prior,
                            analytic_jacobian=False)
            obs_in = mbobs if nband > 1 else mbobs[0]
            r = fitter.go(obs=obs_in, guess=guess)
            p = "c%d_" % icase
            out[p + "images"] = np.array(images)
            out[p + "weights"] = np.array(weights)
            out[p + "jac"], out[p + "psf"] = np.array(jacs), np.array(psfs)
            out[p + "bands"] = np.array(bands, dtype="i4")
            out[p + "guess"] = guess
            out[p + "model"] = np.array(model)
            out[p + "nband"] = np.array(nband)
            out[p + "prior"] = np.array(int(with_prior))
            for key in ("flags", "nfev", "ier", "pars", "pars_cov",
                        "pars_err"):
                out[p + key] = np.array(r[key])
            if r["flags"] == 0:
                for key in ("lnprob", "s2n", "chi2per", "dof", "npix", "T",
                            "T_err"):
                    out[p + key] = np.array(r[key])
                out[p + "flux"] = np.atleast_1d(r["flux"])
                out[p + "flux_err"] = np.atleast_1d(r["flux_err"])
            print(icase, model, nband, nep, "flags", r["flags"], "nfev",
                  r["nfev"], "ier", r["ier"])
            icase += 1
    out["ncases"] = np.array(icase)
    save("lm_multi_fd", **out)


def gen_lm_multi_bounds():
    """Fitter.go on ObsList / MultiBandObsList inputs with prior.bounds
    (leastsqbound's transforms, fitters.py:99-108 -> run_leastsq(bounds=)):
    exp objects with bounded T and per-band fluxes through lmder and lmdif,
    bdf objects with fracdev in (0, 1) through lmdif.  Run with the
    1-based-ipvt shim (see _OneBasedMinpack); the fit itself is asserted
    identical without it."""
    import ngmix.fitting.leastsqbound as lsb
    from ngmix.fitting import Fitter
    from ngmix import priors, joint_prior
    from ngmix_b200.sims import (diag_jacobian, simple_full_pars,
                                 bdf_full_pars, convolve_full_pars,
                                 render_full_pars)
    rng = np.random.RandomState(27182)
    configs = [
        # (nband, epochs per band, model, analytic jacobian)
        (1, [3], "exp", True), (2, [1, 2], "exp", True),
        (3, [1, 1, 1], "exp", True), (2, [1, 1], "exp", False),
        (2, [1, 1], "bdf", False), (1, [2], "bdf", False),
    ]
    out = {}
    icase = 0
    H = W = 32
    real = lsb._minpack
    for nband, nep, model, analytic in configs:
        for rep in range(2):
            noise = [0.02, 0.2][rep]
            g1, g2 = rng.normal(0.0, 0.12, size=2)
            T = rng.uniform(0.2, 0.8)
            fracdev = rng.uniform(0.2, 0.8)
            fluxes = rng.uniform(60.0, 180.0, size=nband)
            images, weights, jacs, psfs, bands = [], [], [], [], []
            for b in range(nband):
                for _ in range(nep[b]):
                    dy, dx = rng.uniform(-0.5, 0.5, size=2)
                    jac = diag_jacobian((H - 1) / 2 + dy, (W - 1) / 2 + dx,
                                        0.263)
                    ppars = [0.0, 0.0, rng.normal(0, 0.01),
                             rng.normal(0, 0.01), rng.uniform(0.22, 0.32), 1.0]
                    psf = simple_full_pars(ppars, "turb")
                    if model == "exp":
                        obj = simple_full_pars(
                            [0.0, 0.0, g1, g2, T, fluxes[b]], "exp")
                    else:
                        obj = bdf_full_pars(
                            [0.0, 0.0, g1, g2, T, fracdev, fluxes[b]])
                    im = render_full_pars(convolve_full_pars(obj, psf),
                                          (H, W), jac)
                    images.append(im + rng.normal(0.0, noise, size=im.shape))
                    weights.append(np.zeros((H, W)) + 1.0 / noise ** 2)
                    jacs.append(jac)
                    psfs.append(psf)
                    bands.append(b)
            shared = {"exp": [0.0, 0.0, g1, g2, T],
                      "bdf": [0.0, 0.0, g1, g2, T, fracdev]}[model]
            truth = np.array(shared + list(fluxes))
            guess = truth.copy()
            guess[0:2] = rng.uniform(-0.01, 0.01, size=2)
            guess[2:] *= 1.0 + rng.uniform(-0.2, 0.2, size=guess.size - 2)
            mbobs = ngmix.MultiBandObsList()
            k = 0
            for b in range(nband):
                ol = ngmix.ObsList()
                for _ in range(nep[b]):
                    psf_obs = ngmix.Observation(
                        np.zeros((5, 5)) + 1.0,
                        gmix=ngmix.GMix(pars=psfs[k].ravel()))
                    ol.append(ngmix.Observation(
                        images[k], weight=weights[k],
                        jacobian=jac_from_record(jacs[k]), psf=psf_obs))
                    k += 1
                mbobs.append(ol)
            cen = priors.CenPrior(0.0, 0.0, 0.263, 0.263, rng=rng)
            gp = priors.GPriorBA(0.3, rng=rng)
            if model == "exp" and rep == 0:
                Tp = priors.Normal(0.4, 1.0, bounds=[0.01, 100.0], rng=rng)
                Fp = [priors.Normal(100.0, 100.0, bounds=[1.0, 1.0e9],
                                    rng=rng) for _ in range(nband)]
                prior = joint_prior.PriorSimpleSep(cen, gp, Tp, Fp)
            elif model == "exp":
                # tight two-sided and one-sided bounds: the transforms take
                # part in many iterations
                Tp = priors.Normal(0.5, 1.0, bounds=[0.05, 5.0], rng=rng)
                Fp = [priors.Normal(100.0, 100.0, bounds=[1.0, None],
                                    rng=rng) for _ in range(nband)]
                prior = joint_prior.PriorSimpleSep(cen, gp, Tp, Fp)
            else:
                Tp = priors.FlatPrior(-1.0, 1.0e3, rng=rng)
                Fp = [priors.FlatPrior(-100.0, 1.0e9, rng=rng)
                      for _ in range(nband)]
                prior = joint_prior.PriorBDFSep(
                    cen, gp, Tp,
                    priors.Normal(0.5, 0.3, rng=rng, bounds=[0.0, 1.0]), Fp)
            assert prior.bounds is not None
            fitter = Fitter(model=model, prior=prior,
                            analytic_jacobian=analytic)
            obs_in = mbobs if nband > 1 else mbobs[0]
            r0 = fitter.go(obs=obs_in, guess=guess.copy())
            lsb._minpack = _OneBasedMinpack(real)
            try:
                r = fitter.go(obs=obs_in, guess=guess.copy())
            finally:
                lsb._minpack = real
            assert r["nfev"] == r0["nfev"]
            assert np.array_equal(r["pars"], r0["pars"])
            p = "c%d_" % icase
            out[p + "images"] = np.array(images)
            out[p + "weights"] = np.array(weights)
            out[p + "jac"], out[p + "psf"] = np.array(jacs), np.array(psfs)
            out[p + "bands"] = np.array(bands, dtype="i4")
            out[p + "guess"] = guess
            out[p + "model"] = np.array(model)
            out[p + "nband"] = np.array(nband)
            out[p + "analytic"] = np.array(int(analytic))
            out[p + "bounds"] = np.array(
                [[np.nan if v is None else v for v in bb]
                 for bb in prior.bounds])
            out[p + "flags_asrun"] = np.array(r0["flags"])
            for key in ("flags", "nfev", "ier", "pars", "pars_cov",
                        "pars_err"):
                out[p + key] = np.array(r[key])
            if r["flags"] == 0:
                for key in ("lnprob", "s2n", "chi2per", "dof", "npix", "T",
                            "T_err"):
                    out[p + key] = np.array(r[key])
                out[p + "flux"] = np.atleast_1d(r["flux"])
                out[p + "flux_err"] = np.atleast_1d(r["flux_err"])
            print(icase, model, nband, nep, "analytic", analytic, "flags",
                  r["flags"], "asrun", r0["flags"], "nfev", r["nfev"], "ier",
                  r["ier"])
            icase += 1
    out["ncases"] = np.array(icase)
    save("lm_multi_bounds", **out)


def gen_psfflux_template():
    """PSFFluxFitter.go with pixel templates (obs.psf.template / obs.template,
    fitting/results.py:834-864) on single observations and ObsLists"""
    from ngmix.fitting import PSFFluxFitter
    b = sims.make_lm_batch(8, seed0=9450, noise=0.05, dims=(24, 24))
    rng = np.random.RandomState(95)
    b["weights"][1][rng.uniform(size=b["weights"][1].shape) < 0.2] = 0.0
    # psf templates: the psf mixture rendered on the object's own grid
    templates = np.zeros_like(b["images"])
    for i in range(8):
        gm = ngmix.GMix(pars=b["psf"][i].ravel())
        templates[i] = gm.make_image((24, 24),
                                     jacobian=jac_from_record(b["jac"][i]))
        templates[i] *= rng.uniform(0.5, 2.0)
    out = {k: b[k] for k in ("images", "weights", "jac")}
    out["templates"] = templates
    groups = [[0], [1], [2, 3, 4], [5, 6, 7]]
    out["ngroups"] = np.array(len(groups))
    for do_psf in (True, False):
        for normalize in (True, False):
            tag = "p%d_n%d_" % (int(do_psf), int(normalize))
            res = {k: np.zeros(len(groups)) for k in
                   ("flags", "flux", "flux_err", "chi2per", "dof")}
            for ig, grp in enumerate(groups):
                ol = ngmix.ObsList()
                for i in grp:
                    obs = ngmix.Observation(
                        b["images"][i], weight=b["weights"][i],
                        jacobian=jac_from_record(b["jac"][i]))
                    if do_psf:
                        psf_obs = ngmix.Observation(np.zeros((5, 5)) + 1.0)
                        psf_obs.template = templates[i]
                        obs.set_psf(psf_obs)
                    else:
                        obs.template = templates[i]
                    ol.append(obs)
                obs_in = ol if len(grp) > 1 else ol[0]
                r = PSFFluxFitter(do_psf=do_psf,
                                  normalize_psf=normalize).go(obs=obs_in)
                for k in res:
                    res[k][ig] = r[k]
            for k, v in res.items():
                out[tag + k] = v
            print(tag, res["flux"], res["flags"])
    for ig, grp in enumerate(groups):
        out["g%d" % ig] = np.array(grp)
    save("psfflux_template", **out)


GENERATORS = {
    "fastexp": gen_fastexp,
    "gmix": gen_gmix,
    "pixels": gen_pixels,
    "render": gen_render,
    "lm_exp": gen_lm_exp,
    "lm_other": gen_lm_other,
    "em_admom": gen_em_admom,
    "lm_bdf": gen_lm_bdf,
    "lm_bd": gen_lm_bd,
    "guessers": gen_guessers,
    "noise_cov": gen_noise_cov,
    "shape_moments": gen_shape_moments,
    "lm_priors2": gen_lm_priors2,
    "lm_coellip": gen_lm_coellip,
    "gaussap": gen_gaussap,
    "moments": gen_moments,
    "lm_multi": gen_lm_multi,
    "lm_multi_wide": gen_lm_multi_wide,
    "psfflux": gen_psfflux,
    "boot": gen_boot,
    "lm_bounds": gen_lm_bounds,
    "lm_multi_xwide": gen_lm_multi_xwide,
    "lm_multi_fd": gen_lm_multi_fd,
    "lm_multi_bounds": gen_lm_multi_bounds,
    "psfflux_template": gen_psfflux_template,
}

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", nargs="*", default=None)
    args = ap.parse_args()
    os.makedirs(GOLDEN, exist_ok=True)
    for name, fn in GENERATORS.items():
        if args.only and name not in args.only:
            continue
        print("== generating", name)
        fn()
