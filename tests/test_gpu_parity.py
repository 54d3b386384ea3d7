"""
GPU parity tests: the CUDA path (through the C ABI, ngmix_b200/_lib.py)
against the golden vectors of the reference and against the CPU oracle on
seeded inputs.

Tolerances (BASELINE.json north_star): rendered images and log-likelihoods
1e-6 relative (measured here: ~1e-13, asserted at 1e-11), fitted g1/g2/T/flux
1e-5 relative, flags and EM/admom iteration counts bit-exact.
"""
import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu

IMG_TOL = 1e-11       # contract: 1e-6
PAR_TOL = 1e-5        # contract: 1e-5


@pytest.fixture(scope="module")
def ngx():
    from ngmix_b200 import _lib
    _lib.init(0)
    return _lib


def relerr(a, b, floor=1e-300):
    a = np.asarray(a, dtype="f8")
    b = np.asarray(b, dtype="f8")
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), floor))


def recs(gm):
    names = ["p", "row", "col", "irr", "irc", "icc", "det",
             "drr", "drc", "dcc", "norm", "pnorm"]
    return np.array([[gm[n][i] for n in names] for i in range(gm.size)])


# ------------------------------------------------------------------ mixtures
@pytest.mark.parametrize("model", ["gauss", "turb", "exp", "dev", "bdf",
                                   "bd", "coellip"])
def test_gmix_fill(ngx, model):
    g = load_golden("gmix")
    gm, st = ngx.gmix_fill(model, g["pars_" + model])
    assert np.all(st == 0)
    assert np.all(ngx.gmix_set_norms(gm) == 0)
    for i in range(gm.shape[0]):
        assert relerr(recs(gm[i]), g["recs_" + model][i]) < 1e-14
        assert np.all(gm[i]["norm_set"] == 1)


def test_gmix_cm_convolve_and_errors(ngx):
    g = load_golden("gmix")
    cm = np.column_stack([g["cm_in"][:, 0], g["cm_in"][:, 1],
                          g["cm_tfactor"]])
    gm, st = ngx.gmix_fill_cm(cm, g["cm_in"][:, 2:])
    ngx.gmix_set_norms(gm)
    for i in range(gm.shape[0]):
        assert relerr(recs(gm[i]), g["recs_cm"][i]) < 1e-14
    gm0, _ = ngx.gmix_fill("exp", g["conv_obj"])
    gp, _ = ngx.gmix_fill("full", g["conv_psf"])
    gc = ngx.gmix_convolve(gm0, gp)
    ngx.gmix_set_norms(gc)
    for i in range(gc.shape[0]):
        assert relerr(recs(gc[i]), g["conv_recs"][i]) < 1e-14
    # GMixRangeError cases
    _, st = ngx.gmix_fill("exp", [[0, 0, 0.8, 0.7, 1.0, 1.0]])
    assert st[0] == ngx.ST_RANGE_ERROR
    gm, st = ngx.gmix_fill("exp", [[0, 0, 0.1, 0.1, -1.0, 1.0]])
    assert st[0] == 0
    assert ngx.gmix_set_norms(gm)[0] == ngx.ST_RANGE_ERROR


def test_pixels(ngx):
    g = load_golden("pixels")
    for tag, ign in (("ign", True), ("all", False)):
        pix, npix = ngx.make_pixels(g["image"], g["weight"], g["jac"], ign)
        want = g["pixels_" + tag]
        assert npix[0] == want.shape[0]
        got = np.array([pix[0][n][:npix[0]]
                        for n in ("u", "v", "area", "val", "ierr")]).T
        assert np.array_equal(got, want)          # bit exact
    co = ngx.make_coords(g["image"].shape, g["jac"])
    got = np.array([co[0][n] for n in ("u", "v", "area")]).T
    assert np.array_equal(got, g["coords"])


# ------------------------------------------------------------------ passes
def _case_gmix(ngx, g, k):
    p = "c%d_" % k
    gm0, _ = ngx.gmix_fill(str(g[p + "model"]), g[p + "pars"])
    gp, _ = ngx.gmix_fill("full", g[p + "psf"].ravel())
    gm = ngx.gmix_convolve(gm0, gp)
    return gm


def test_render_loglike_fdiff_golden(ngx):
    g = load_golden("render")
    for k in range(int(g["ncases"])):
        p = "c%d_" % k
        gm = _case_gmix(ngx, g, k)
        dims = tuple(g[p + "dims"])
        fast, st = ngx.render(gm, g[p + "jac"], dims, fast_exp=True)
        assert st[0] == 0
        scale = g[p + "fast"].max()
        assert np.abs(fast[0] - g[p + "fast"]).max() / scale < IMG_TOL
        nz = g[p + "fast"] > 1e-12 * scale
        assert relerr(fast[0][nz], g[p + "fast"][nz]) < IMG_TOL
        # the apodized evaluation is exactly zero where the reference's is
        assert np.array_equal(fast[0] == 0, g[p + "fast"] == 0)
        exact, _ = ngx.render(gm, g[p + "jac"], dims, fast_exp=False)
        assert relerr(exact[0], g[p + "exact"]) < IMG_TOL
        out4, fdiff, st = ngx.loglike(gm, g[p + "image"], g[p + "weight"],
                                      g[p + "jac"], want_fdiff=True)
        assert st[0] == 0
        assert relerr(out4[0][:3], g[p + "loglike"][:3]) < IMG_TOL
        assert out4[0][3] == g[p + "loglike"][3]
        assert relerr(out4[0][2], g[p + "s2n_sum"]) < IMG_TOL
        want = g[p + "fdiff"][13:]
        n = int(out4[0][3])
        assert np.abs(fdiff[0][:n] - want[:n]).max() \
            / np.abs(want).max() < IMG_TOL
        # the reference's own calling convention: pre-built pixel records
        pix, npix = ngx.make_pixels(g[p + "image"], g[p + "weight"],
                                    g[p + "jac"])
        out4b, fdb, _ = ngx.loglike_pixels(gm, pix, npix, want_fdiff=True)
        assert relerr(out4b[0][:3], g[p + "loglike"][:3]) < IMG_TOL
        assert np.abs(fdb[0][:n] - want[:n]).max() \
            / np.abs(want).max() < IMG_TOL


def test_appendix_b_known_answers(ngx):
    g = load_golden("render")
    gm = _case_gmix(ngx, g, 0)
    im, _ = ngx.render(gm, g["c0_jac"], (48, 48), fast_exp=True)
    im = im[0]
    assert abs(im.sum() / 99.99852896955666 - 1) < 1e-13
    assert abs(im.max() / 4.278568261192419 - 1) < 1e-13
    assert abs(im[24, 22] / 3.953826189964537 - 1) < 1e-13
    assert im[0, 0] == 0.0
    assert abs(im[10, 40] / 1.3303719076748628e-16 - 1) < 1e-11
    out4, _, _ = ngx.loglike(gm, np.zeros((48, 48)), np.zeros((48, 48)) + 4.0,
                             g["c0_jac"])
    assert abs(out4[0][0] / -378.46427897109817 - 1) < 1e-13
    assert out4[0][1] == 0.0 and out4[0][3] == 2304


def test_deriv_images(ngx):
    g = load_golden("render")
    for k in range(int(g["ncases"])):
        p = "c%d_" % k
        out = ngx.deriv_images(g[p + "gpars"], g[p + "dcov"], g[p + "jac"],
                               tuple(g[p + "dims"]))
        want = g[p + "deriv"]
        for a in range(6):
            assert np.abs(out[0][a] - want[a]).max() \
                / np.abs(want[a]).max() < IMG_TOL


def test_render_vs_oracle_random(ngx, ora):
    """seeded random mixtures, odd sizes, non-diagonal jacobians"""
    rng = np.random.RandomState(11)
    for dims in ((25, 25), (31, 47), (64, 64), (5, 3)):
        n = 6
        pars = np.zeros((n, 6))
        pars[:, 0:2] = rng.uniform(-0.5, 0.5, (n, 2))
        pars[:, 2:4] = rng.uniform(-0.4, 0.4, (n, 2))
        pars[:, 4] = rng.uniform(0.1, 3.0, n)
        pars[:, 5] = rng.uniform(1, 100, n)
        jac = np.zeros((n, 8))
        for i in range(n):
            a, b, c, d = 0.26 + 0.02 * rng.randn(), 0.02 * rng.randn(), \
                0.02 * rng.randn(), 0.26 + 0.02 * rng.randn()
            det = a * d - b * c
            jac[i] = [dims[0] / 2 + rng.uniform(-1, 1),
                      dims[1] / 2 + rng.uniform(-1, 1), a, b, c, d, det,
                      np.sqrt(abs(det))]
        gm, _ = ngx.gmix_fill("dev", pars)
        im, st = ngx.render(gm, jac, dims, fast_exp=True)
        for i in range(n):
            og, _ = ora.gmix_fill("dev", pars[i])
            oim, _ = ora.render(og, dims, jac[i], fast_exp=True)
            sc = oim.max()
            if sc > 0:
                assert np.abs(im[i] - oim).max() / sc < IMG_TOL


@pytest.mark.parametrize("dims", [(48, 48), (31, 47), (5, 3), (64, 64)])
@pytest.mark.parametrize("mask", ["none", "some", "negative_kept"])
def test_loglike_fdiff_vs_oracle_masks(ngx, ora, dims, mask):
    """fdiff order and values with no masked pixels (direct row-major
    positions), with several masked pixels (compaction tables) and with
    ignore_zero_weight=False, on stamp sizes with padding tiles"""
    rng = np.random.RandomState(5 + dims[0])
    n = 4
    pars = np.zeros((n, 6))
    pars[:, 0:2] = rng.uniform(-0.3, 0.3, (n, 2))
    pars[:, 2:4] = rng.uniform(-0.3, 0.3, (n, 2))
    pars[:, 4] = rng.uniform(0.2, 2.0, n)
    pars[:, 5] = rng.uniform(10, 100, n)
    jac = np.zeros((n, 8))
    for i in range(n):
        jac[i] = [(dims[0] - 1) / 2 + rng.uniform(-0.5, 0.5),
                  (dims[1] - 1) / 2 + rng.uniform(-0.5, 0.5),
                  0.263, 0.0, 0.0, 0.263, 0.263 ** 2, 0.263]
    gm, _ = ngx.gmix_fill("exp", pars)
    ngx.gmix_set_norms(gm)
    images = rng.normal(0.0, 0.1, (n,) + dims)
    weights = np.full((n,) + dims, 100.0)
    izw = True
    if mask != "none":
        nbad = max(1, dims[0] * dims[1] // 7)
        for i in range(n):
            idx = rng.choice(dims[0] * dims[1], nbad, replace=False)
            weights[i].reshape(-1)[idx] = 0.0
            weights[i].reshape(-1)[idx[: nbad // 2]] = -1.0
        izw = mask == "some"
    out4, fdiff, st = ngx.loglike(gm, images, weights, jac,
                                  ignore_zero_weight=izw, want_fdiff=True)
    assert not st.any()
    for i in range(n):
        og, _ = ora.gmix_fill("exp", pars[i])
        ora.gmix_set_norms(og)
        pix = ora.make_pixels(images[i], weights[i], jac[i],
                              ignore_zero_weight=izw)
        want = np.zeros(pix.size)
        ora.fill_fdiff(og, pix, want)
        o4, _ = ora.get_loglike(og, pix)
        assert out4[i][3] == pix.size == o4[3]
        assert relerr(out4[i][:3], o4[:3]) < IMG_TOL
        assert np.abs(fdiff[i][:pix.size] - want).max() \
            / np.abs(want).max() < IMG_TOL
        assert not fdiff[i][pix.size:].any()


# ------------------------------------------------------------------ LM
def _prior(ngx, kind):
    pr = ngx.Prior()
    if kind == "none":
        return pr
    pr.present = 1
    pr.g_kind = 1
    pr.cen1, pr.cen2, pr.sigma1, pr.sigma2 = 0.0, 0.0, 0.263, 0.263
    pr.g_sigma = 0.3
    if kind == "flat":
        pr.T = ngx.Prior1D(1, 0, -1.0, 1.0e3, 0, 0)
        pr.F = ngx.Prior1D(1, 0, -100.0, 1.0e9, 0, 0)
    elif kind == "erf":
        pr.T = ngx.Prior1D(2, 0, -1.0, 0.1, 1.0e3, 1.0)
        pr.F = ngx.Prior1D(2, 0, -1.0e2, 1.0, 1.0e9, 1.0e8)
    elif kind == "bdf":
        pr.T = ngx.Prior1D(1, 0, -1.0, 1.0e3, 0, 0)
        pr.fracdev = ngx.Prior1D(3, 0, 0.5, 0.1, 0, 0)
        pr.F = ngx.Prior1D(1, 0, -100.0, 1.0e9, 0, 0)
    elif kind == "bd":
        pr.T = ngx.Prior1D(1, 0, -1.0, 1.0e3, 0, 0)
        pr.logTratio = ngx.Prior1D(3, 0, 0.0, 0.5, 0, 0)
        pr.fracdev = ngx.Prior1D(3, 0, 0.5, 0.1, 0, 0)
        pr.F = ngx.Prior1D(1, 0, -100.0, 1.0e9, 0, 0)
    return pr


def _check_lm(ngx, g, tag, pk, model, analytic=True, allow_mismatch=0):
    pre = "%s_%s_" % (tag, pk)
    res = ngx.lm_fit(model, g[tag + "_images"], g[tag + "_weights"],
                     g[tag + "_jac"], g[tag + "_psf"], g[tag + "_guess"],
                     prior=_prior(ngx, pk), analytic_jacobian=analytic)
    n = res.size
    np_ = g[pre + "pars"].shape[1]
    # flags bit exact (contract); ier/nfev are diagnostics that must also
    # agree except where a convergence test lands within rounding of its
    # threshold
    assert np.array_equal(res["flags"], g[pre + "flags"])
    same = (res["nfev"] == g[pre + "nfev"]) & (res["ier"] == g[pre + "ier"])
    assert (~same).sum() <= allow_mismatch, \
        (res["nfev"], g[pre + "nfev"], res["ier"], g[pre + "ier"])
    for i in range(n):
        if res["flags"][i] != 0:
            continue
        ref = g[pre + "pars"][i]
        got = res["pars"][i][:np_]
        # g1, g2, T, flux to 1e-5 relative (small |g| measured against the
        # parameter error, as relative error of a value near zero is moot)
        err = g[pre + "pars_err"][i]
        tol = np.maximum(PAR_TOL * np.abs(ref), 1e-3 * err)
        assert np.all(np.abs(got - ref) <= tol), (tag, pk, i, got, ref)
        if not same[i]:
            continue
        assert relerr(res["pars_err"][i][:np_], g[pre + "pars_err"][i]) < 1e-4
        for k in ("lnprob", "s2n", "chi2per", "T_err", "flux_err"):
            if pre + k in g.files:
                assert relerr(res[k][i], g[pre + k][i]) < 1e-6, (k, i)
        if pre + "npix" in g.files:
            assert res["npix"][i] == g[pre + "npix"][i]
            assert res["dof"][i] == g[pre + "dof"][i]
    return res


@pytest.mark.parametrize("tag,pk", [("c1", "none"), ("hi", "none"),
                                    ("hi", "flat"), ("lo", "none"),
                                    ("lo", "flat"), ("lo", "erf")])
def test_lm_exp_golden(ngx, tag, pk):
    g = load_golden("lm_exp")
    _check_lm(ngx, g, tag, pk, "exp")


@pytest.mark.parametrize("tag,pk,model,analytic", [
    ("gauss", "none", "gauss", True), ("dev", "flat", "dev", True),
    ("expfd", "none", "exp", False), ("bad", "none", "exp", True),
    ("bad", "flat", "exp", True)])
def test_lm_other_golden(ngx, tag, pk, model, analytic):
    g = load_golden("lm_other")
    _check_lm(ngx, g, tag, pk, model, analytic=analytic)


def test_lm_bdf_golden(ngx):
    g = load_golden("lm_bdf")
    res = ngx.lm_fit("bdf", g["images"], g["weights"], g["jac"], g["psf"],
                     g["guess"], prior=_prior(ngx, "bdf"))
    assert np.array_equal(res["flags"], g["res_flags"])
    for i in range(res.size):
        ref = g["res_pars"][i]
        err = g["res_pars_err"][i]
        tol = np.maximum(PAR_TOL * np.abs(ref), 1e-2 * err)
        assert np.all(np.abs(res["pars"][i][:7] - ref) <= tol), i


@pytest.mark.parametrize("pk", ["none", "bd"])
def test_lm_bd_golden(ngx, pk):
    """bd model (8 parameters, forward differences), PriorBDSep rows"""
    g = load_golden("lm_bd")
    res = ngx.lm_fit("bd", g["images"], g["weights"], g["jac"], g["psf"],
                     g["guess"], prior=_prior(ngx, pk))
    assert np.array_equal(res["flags"], g[pk + "_flags"])
    assert np.array_equal(res["nfev"], g[pk + "_nfev"])
    assert np.array_equal(res["ier"], g[pk + "_ier"])
    for i in range(res.size):
        ref = g[pk + "_pars"][i]
        err = g[pk + "_pars_err"][i]
        tol = np.maximum(PAR_TOL * np.abs(ref), 1e-2 * err)
        assert np.all(np.abs(res["pars"][i][:8] - ref) <= tol), i
        assert relerr(res["lnprob"][i], g[pk + "_lnprob"][i]) < 1e-6
        assert res["dof"][i] == g[pk + "_dof"][i]


def test_lm_vs_oracle_batch(ngx, ora):
    """config-2 style batch vs the oracle, incl. zero-weight pixels"""
    from ngmix_b200 import sims
    b = sims.make_lm_batch(48, seed0=77000, noise=0.05)
    b["weights"][:, 10:14, 20:23] = 0.0          # masked pixels are dropped
    b["weights"][3] = 0.0                        # a stamp with no data
    res = ngx.lm_fit("exp", b["images"], b["weights"], b["jac"], b["psf"],
                     b["guess"], prior=_prior(ngx, "flat"))
    assert res["flags"][3] == 1                  # NO_ATTEMPT
    ok = np.ones(48, bool)
    ok[3] = False
    ores = ora.lm_fit_batch("exp", b["images"][ok], b["weights"][ok],
                            b["jac"][ok], b["psf"][ok], b["guess"][ok],
                            prior=ora.make_prior("flat"), nthreads=4)
    r = res[ok]
    assert np.array_equal(r["flags"], ores["flags"])
    assert np.array_equal(r["npix"], ores["npix"])
    assert np.mean((r["nfev"] == ores["nfev"])
                   & (r["ier"] == ores["ier"])) > 0.95
    for i in range(r.size):
        tol = np.maximum(PAR_TOL * np.abs(ores["pars"][i][:6]),
                         1e-3 * ores["pars_err"][i][:6])
        assert np.all(np.abs(r["pars"][i][:6] - ores["pars"][i][:6]) <= tol)


# ------------------------------------------------------------------ EM
def test_em_golden(ngx):
    g = load_golden("em_admom")
    res, pars = ngx.em(g["images"], g["weights"], g["jac"], g["em_guess"])
    # flags and iteration counts bit exact
    assert np.array_equal(res["flags"], g["em_flags"])
    assert np.array_equal(res["numiter"], g["em_numiter"])
    assert relerr(res["sky"], g["em_sky"]) < 1e-12
    n = res.size
    assert relerr(pars.reshape(n, -1), g["em_pars"], floor=1e-6) < PAR_TOL


@pytest.mark.parametrize("tag", ["psf", "fixcen", "fixcov", "fluxonly",
                                 "varysky", "zerowt", "sky"])
def test_em_variants_golden(ngx, tag):
    g = load_golden("em_admom")
    psf = g["emv_psf"]
    variant = tag if tag in ("fixcen", "fixcov", "fluxonly") else "full"
    images = g["images"][:8].copy()
    weights = g["weights"][:8].copy()
    kw = {}
    if tag == "zerowt":
        weights[:, 5:8, 6:9] = 0.0
        kw["ignore_zero_weight"] = False
    if tag == "varysky":
        kw["vary_sky"] = True
    if tag == "sky":
        images = np.abs(images) + 1e-3
        kw["sky"] = 0.0
    use_psf = tag in ("psf", "fixcen", "fixcov", "fluxonly")
    psf_pars = np.repeat(psf[None], 8, axis=0) if use_psf else None
    res, pars = ngx.em(images, weights, g["jac"][:8], g["em_guess"][:8, :2],
                       psf_pars=psf_pars, variant=variant, **kw)
    assert np.array_equal(res["flags"], g["emv_%s_flags" % tag])
    good = (res["flags"] & 128) == 0
    assert np.array_equal(res["numiter"][good],
                          g["emv_%s_numiter" % tag][good])
    if good.any():
        assert relerr(res["sky"][good], g["emv_%s_sky" % tag][good]) < 1e-11
        assert relerr(pars.reshape(8, -1)[good],
                      g["emv_%s_pars" % tag][good], floor=1e-6) < PAR_TOL


# ------------------------------------------------------------------ admom
def test_admom_golden(ngx):
    g = load_golden("em_admom")
    res = ngx.admom(g["images"], g["weights"], g["jac"], g["am_guess"])
    assert np.array_equal(res["numiter"], g["am_numiter"])
    assert np.array_equal(res["npix"], g["am_npix"])
    assert np.all(res["flags"] == 0)
    assert relerr(res["pars"][:, :5], g["am_pars"][:, :5], floor=1e-6) \
        < 1e-9
    assert relerr(res["sums"], g["am_sums"], floor=1e-12) < 1e-8
    assert relerr(res["wsum"], g["am_wsum"]) < 1e-11
    assert relerr(res["rho4"], g["am_rho4"]) < 1e-9
    for i in range(res.size):
        sc = g["am_sums_cov"][i]
        assert np.abs(res["sums_cov"][i] - sc).max() / np.abs(sc).max() < 1e-9
    # failure modes
    resf = ngx.admom(g["amf_images"], np.repeat(g["weights"][:1], 4, axis=0),
                     g["jac"][:4], g["amf_guess"])
    assert np.array_equal(resf["flags"], g["amf_flags"] & ~64)
    assert np.array_equal(resf["numiter"], g["amf_numiter"])


def test_em_admom_vs_oracle_batch(ngx, ora):
    from ngmix_b200 import sims
    b = sims.make_psf_batch(128, seed0=555000)
    res, pars = ngx.em(b["images"], b["weights"], b["jac"], b["em_guess"])
    ores, opars = ora.em_go_batch(b["images"], b["weights"], b["jac"],
                                  b["em_guess"], nthreads=4)
    assert np.array_equal(res["flags"], ores["flags"])
    assert np.array_equal(res["numiter"], ores["numiter"])
    ares = ngx.admom(b["images"], b["weights"], b["jac"], b["am_guess"])
    oares = ora.admom_go_batch(b["images"], b["weights"], b["jac"],
                               b["am_guess"], nthreads=4)
    assert np.array_equal(ares["flags"], oares["flags"])
    assert np.array_equal(ares["numiter"], oares["numiter"])
    assert relerr(ares["pars"][:, :5], oares["pars"][:, :5], floor=1e-6) \
        < 1e-9
