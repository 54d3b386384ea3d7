"""
TEST INFRASTRUCTURE - ctypes wrapper of the CPU oracle (oracle/liboracle.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module.  The product package
(ngmix_b200) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIBPATH = os.path.join(HERE, "liboracle.so")

# numpy views of the reference's record layouts
GAUSS_DTYPE = np.dtype([
    ("p", "f8"), ("row", "f8"), ("col", "f8"), ("irr", "f8"), ("irc", "f8"),
    ("icc", "f8"), ("det", "f8"), ("norm_set", "i8"), ("drr", "f8"),
    ("drc", "f8"), ("dcc", "f8"), ("norm", "f8"), ("pnorm", "f8")])
PIXEL_DTYPE = np.dtype([
    ("u", "f8"), ("v", "f8"), ("area", "f8"), ("val", "f8"), ("ierr", "f8"),
    ("fdiff", "f8")])
COORD_DTYPE = np.dtype([("u", "f8"), ("v", "f8"), ("area", "f8")])

MODELS = {"full": 0, "gauss": 1, "turb": 2, "exp": 3, "dev": 4, "bdf": 6,
          "coellip": 7, "cm": 9, "bd": 10}
NGAUSS = {"gauss": 1, "turb": 3, "exp": 6, "dev": 10, "bdf": 16, "bd": 16,
          "cm": 16}
NPARS = {"gauss": 6, "turb": 6, "exp": 6, "dev": 6, "cm": 6, "bdf": 7,
         "bd": 8}


class Prior1D(C.Structure):
    _fields_ = [("kind", C.c_int32), ("pad", C.c_int32), ("a", C.c_double),
                ("b", C.c_double), ("c", C.c_double), ("d", C.c_double)]


class Prior(C.Structure):
    _fields_ = [("present", C.c_int32), ("g_kind", C.c_int32),
                ("cen1", C.c_double), ("cen2", C.c_double),
                ("sigma1", C.c_double), ("sigma2", C.c_double),
                ("g_sigma", C.c_double), ("T", Prior1D),
                ("fracdev", Prior1D), ("F", Prior1D),
                ("logTratio", Prior1D), ("ncoellip", C.c_int32),
                ("pad2", C.c_int32)]


class LMResult(C.Structure):
    _fields_ = [("flags", C.c_int32), ("nfev", C.c_int32), ("ier", C.c_int32),
                ("npix", C.c_int32), ("dof", C.c_int32), ("njev", C.c_int32),
                ("lnprob", C.c_double), ("s2n_numer", C.c_double),
                ("s2n_denom", C.c_double), ("chi2per", C.c_double),
                ("s2n", C.c_double), ("T", C.c_double), ("T_err", C.c_double),
                ("flux", C.c_double), ("flux_err", C.c_double),
                ("pars", C.c_double * 20), ("pars_err", C.c_double * 20),
                ("pars_cov0", C.c_double * 400),
                ("pars_cov", C.c_double * 400)]


LM_RESULT_DTYPE = np.dtype([
    ("flags", "i4"), ("nfev", "i4"), ("ier", "i4"), ("npix", "i4"),
    ("dof", "i4"), ("njev", "i4"), ("lnprob", "f8"), ("s2n_numer", "f8"),
    ("s2n_denom", "f8"), ("chi2per", "f8"), ("s2n", "f8"), ("T", "f8"),
    ("T_err", "f8"), ("flux", "f8"), ("flux_err", "f8"), ("pars", "f8", 20),
    ("pars_err", "f8", 20), ("pars_cov0", "f8", 400),
    ("pars_cov", "f8", 400)])

EM_RESULT_DTYPE = np.dtype([("flags", "i4"), ("numiter", "i4"),
                            ("fdiff", "f8"), ("sky", "f8")])

ADMOM_CONF_DTYPE = np.dtype([("maxiter", "i4"), ("pad", "i4"),
                             ("shiftmax", "f8"), ("etol", "f8"),
                             ("Ttol", "f8"), ("cenonly", "i4"),
                             ("pad2", "i4")])
ADMOM_RESULT_DTYPE = np.dtype([
    ("flags", "i4"), ("numiter", "i4"), ("npix", "i4"), ("pad", "i4"),
    ("wsum", "f8"), ("sums", "f8", 7), ("sums_cov", "f8", (7, 7)),
    ("pars", "f8", 6), ("rho4", "f8"), ("F", "f8", 7)])

_lib = None


def build():
    """compile liboracle.so (gcc) if missing or stale"""
    srcs = [os.path.join(HERE, f) for f in os.listdir(HERE)
            if f.endswith(".c") or f.endswith(".h")]
    if os.path.exists(LIBPATH):
        mt = os.path.getmtime(LIBPATH)
        if all(os.path.getmtime(s) <= mt for s in srcs):
            return
    subprocess.check_call(["make", "-C", HERE, "liboracle.so"],
                          stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIBPATH):
            build()
        _lib = C.CDLL(LIBPATH)
        _lib.ora_fexp.restype = C.c_double
        _lib.ora_fexp.argtypes = [C.c_double]
        _lib.ora_get_cm_Tfactor.restype = C.c_double
        _lib.ora_get_cm_Tfactor.argtypes = [C.c_double, C.c_double]
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _f8(a):
    return np.ascontiguousarray(a, dtype="f8")


# ------------------------------------------------------------------ fastexp
def fexp(x):
    x = _f8(np.atleast_1d(x))
    out = np.zeros_like(x)
    lib().ora_fexp_arr(C.c_int(x.size), _p(x), _p(out))
    return out


def apod(chi2):
    chi2 = _f8(np.atleast_1d(chi2))
    w = np.zeros_like(chi2)
    dw = np.zeros_like(chi2)
    lib().ora_apod_arr(C.c_int(chi2.size), _p(chi2), _p(w), _p(dw))
    return w, dw


# ------------------------------------------------------------------ gmix
def gmix_fill(model, pars, ngauss=None):
    pars = _f8(pars)
    if model == "coellip":
        ngauss = (pars.size - 4) // 2
    elif model == "full":
        ngauss = pars.size // 6
    else:
        ngauss = NGAUSS[model]
    gm = np.zeros(ngauss, dtype=GAUSS_DTYPE)
    st = lib().ora_gmix_fill(C.c_int(MODELS[model]), _p(pars),
                             C.c_int(pars.size), _p(gm), C.c_int(ngauss))
    return gm, st


def gmix_fill_cm(fracdev, TdByTe, pars):
    pars = _f8(pars)
    gm = np.zeros(16, dtype=GAUSS_DTYPE)
    tf = lib().ora_get_cm_Tfactor(fracdev, TdByTe)
    st = lib().ora_gmix_fill_cm(C.c_double(fracdev), C.c_double(TdByTe),
                                C.c_double(tf), _p(pars), _p(gm))
    return gm, tf, st


def gmix_set_norms(gm):
    return lib().ora_gmix_set_norms(_p(gm), C.c_int(gm.size))


def gmix_convolve(gm, psf):
    out = np.zeros(gm.size * psf.size, dtype=GAUSS_DTYPE)
    lib().ora_gmix_convolve_fill(_p(out), _p(gm), C.c_int(gm.size), _p(psf),
                                 C.c_int(psf.size))
    return out


def gm_records(gm):
    names = ["p", "row", "col", "irr", "irc", "icc", "det",
             "drr", "drc", "dcc", "norm", "pnorm"]
    return np.array([[gm[n][i] for n in names] for i in range(gm.size)])


# ------------------------------------------------------------------ pixels
def make_pixels(image, weight, jac, ignore_zero_weight=True):
    image = _f8(image)
    weight = _f8(weight)
    jac = _f8(jac)
    if ignore_zero_weight:
        npix = int((weight > 0).sum())
    else:
        npix = image.size
    pixels = np.zeros(npix, dtype=PIXEL_DTYPE)
    n = lib().ora_fill_pixels(_p(pixels), C.c_int(npix), _p(image),
                              _p(weight), C.c_int(image.shape[0]),
                              C.c_int(image.shape[1]), _p(jac),
                              C.c_int(int(ignore_zero_weight)))
    assert n == npix
    return pixels


def make_coords(dims, jac):
    jac = _f8(jac)
    coords = np.zeros(dims[0] * dims[1], dtype=COORD_DTYPE)
    lib().ora_fill_coords(_p(coords), C.c_int(dims[0]), C.c_int(dims[1]),
                          _p(jac))
    return coords


# ------------------------------------------------------------------ passes
def render(gm, dims, jac, fast_exp=False):
    coords = make_coords(dims, jac)
    image = np.zeros(dims[0] * dims[1])
    gm = gm.copy()
    st = lib().ora_render(_p(gm), C.c_int(gm.size), _p(coords),
                          C.c_int(coords.size), _p(image),
                          C.c_int(int(fast_exp)))
    return image.reshape(dims), st


def fill_fdiff(gm, pixels, fdiff, start=0):
    gm = gm.copy()
    return lib().ora_fill_fdiff(_p(gm), C.c_int(gm.size), _p(pixels),
                                C.c_int(pixels.size), _p(fdiff),
                                C.c_int(start))


def get_loglike(gm, pixels):
    gm = gm.copy()
    out = np.zeros(4)
    st = lib().ora_get_loglike(_p(gm), C.c_int(gm.size), _p(pixels),
                               C.c_int(pixels.size), _p(out))
    return out, st


def get_model_s2n_sum(gm, pixels):
    gm = gm.copy()
    out = np.zeros(1)
    lib().ora_get_model_s2n_sum(_p(gm), C.c_int(gm.size), _p(pixels),
                                C.c_int(pixels.size), _p(out))
    return out[0]


def deriv_images(gpars, dcov, vv, uu, area):
    gpars = _f8(gpars)
    dcov = _f8(dcov)
    vv, uu, area = _f8(vv), _f8(uu), _f8(area)
    out = np.zeros((6, vv.size))
    lib().ora_deriv_images(_p(gpars), _p(dcov), C.c_int(gpars.shape[0]),
                           _p(vv), _p(uu), _p(area), C.c_int(vv.size),
                           _p(out))
    return out


# ------------------------------------------------------------------ priors
def _lognormal(mean, sigma, shift=0.0):
    """LogNormal(mean, sigma, shift) -> (logmean, logivar, lnprob_max,
    shift), priors/priors.py:688-716"""
    logvar = np.log(1 + sigma ** 2 / mean ** 2)
    logmean = np.log(mean) - 0.5 * logvar
    log_mode = logmean - logvar
    chi2 = (1.0 / logvar) * (log_mode - logmean) ** 2
    return Prior1D(4, 0, logmean, 1.0 / logvar, -0.5 * chi2 - log_mode, shift)


def make_prior(kind, model="exp"):
    """the prior sets of the golden fixtures (gen_golden._get_priors)"""
    pr = Prior()
    if kind == "none" or kind is None:
        pr.present = 0
        return pr
    pr.present = 1
    pr.g_kind = 1
    pr.cen1, pr.cen2, pr.sigma1, pr.sigma2 = 0.0, 0.0, 0.263, 0.263
    pr.g_sigma = 0.3
    if kind == "flat":
        pr.T = Prior1D(1, 0, -1.0, 1.0e3, 0, 0)
        pr.F = Prior1D(1, 0, -100.0, 1.0e9, 0, 0)
    elif kind == "erf":
        pr.T = Prior1D(2, 0, -1.0, 0.1, 1.0e3, 1.0)
        pr.F = Prior1D(2, 0, -1.0e2, 1.0, 1.0e9, 1.0e8)
    elif kind == "bdf":
        pr.T = Prior1D(1, 0, -1.0, 1.0e3, 0, 0)
        pr.fracdev = Prior1D(3, 0, 0.5, 0.1, 0, 0)
        pr.F = Prior1D(1, 0, -100.0, 1.0e9, 0, 0)
    elif kind == "lognormal":
        pr.T = _lognormal(0.5, 0.3)
        pr.F = _lognormal(100.0, 50.0, shift=-5.0)
    elif kind == "trunc":
        pr.T = Prior1D(5, 0, 0.5, 0.4, 0.0, 5.0)
        pr.F = _lognormal(100.0, 50.0)
    elif kind == "bdf2":
        pr.T = _lognormal(0.5, 0.3)
        pr.fracdev = Prior1D(5, 0, 0.5, 0.1, 0.0, 1.0)
        pr.F = _lognormal(100.0, 50.0)
    elif kind == "bd":
        pr.T = Prior1D(1, 0, -1.0, 1.0e3, 0, 0)
        pr.logTratio = Prior1D(3, 0, 0.0, 0.5, 0, 0)
        pr.fracdev = Prior1D(3, 0, 0.5, 0.1, 0, 0)
        pr.F = Prior1D(1, 0, -100.0, 1.0e9, 0, 0)
    else:
        raise ValueError(kind)
    return pr


# ------------------------------------------------------------------ LM
class Bounds(C.Structure):
    """prior.bounds of the reference (joint_prior.py:37-61)"""
    _fields_ = [("present", C.c_int32), ("type", C.c_int32 * 20),
                ("pad", C.c_int32),
                ("lo", C.c_double * 20), ("hi", C.c_double * 20)]


def make_bounds(bounds):
    b = Bounds()
    b.present = 0
    if bounds is None:
        return b
    b.present = 1
    for i, (lo, hi) in enumerate(bounds):
        b.type[i] = (1 if lo is not None else 0) + (2 if hi is not None else 0)
        b.lo[i] = 0.0 if lo is None else lo
        b.hi[i] = 0.0 if hi is None else hi
    return b


def lm_fit(model, image, weight, jac, psf_pars, guess, prior=None,
           maxfev=4000, ftol=1.0e-5, xtol=1.0e-5, analytic_jacobian=True,
           ignore_zero_weight=True, bounds=None):
    image, weight, jac, guess = _f8(image), _f8(weight), _f8(jac), _f8(guess)
    if psf_pars is None:
        npsf = 0
        psf_ptr = None
    else:
        psf_pars = _f8(psf_pars)
        npsf = psf_pars.size // 6
        psf_ptr = _p(psf_pars)
    pr = prior if prior is not None else make_prior("none")
    if model == "coellip":
        pr.ncoellip = (guess.size - 4) // 2
    res = np.zeros(1, dtype=LM_RESULT_DTYPE)
    bnd = make_bounds(bounds)
    st = lib().ora_lm_fit_bounded(
        C.c_int(MODELS[model]), _p(image), _p(weight),
        C.c_int(image.shape[0]), C.c_int(image.shape[1]), _p(jac), psf_ptr,
        C.c_int(npsf), _p(guess), C.byref(pr), C.byref(bnd), C.c_int(maxfev),
        C.c_double(ftol), C.c_double(xtol), C.c_int(int(analytic_jacobian)),
        C.c_int(int(ignore_zero_weight)), _p(res))
    assert st == 0, st
    return res[0]


def lm_fit_batch(model, images, weights, jacs, psf_pars, guess, prior=None,
                 nthreads=1, maxfev=4000, ftol=1.0e-5, xtol=1.0e-5,
                 analytic_jacobian=True):
    images, weights = _f8(images), _f8(weights)
    jacs, guess = _f8(jacs), _f8(guess)
    n, H, W = images.shape
    if psf_pars is None:
        npsf = 0
        psf_ptr = None
    else:
        psf_pars = _f8(psf_pars)
        npsf = psf_pars.shape[1]
        psf_ptr = _p(psf_pars)
    pr = prior if prior is not None else make_prior("none")
    res = np.zeros(n, dtype=LM_RESULT_DTYPE)
    lib().ora_lm_fit_batch(
        C.c_int(n), C.c_int(nthreads), C.c_int(MODELS[model]), _p(images),
        _p(weights), C.c_int(H), C.c_int(W), _p(jacs), psf_ptr, C.c_int(npsf),
        _p(guess), C.byref(pr), C.c_int(maxfev), C.c_double(ftol),
        C.c_double(xtol), C.c_int(int(analytic_jacobian)), _p(res))
    return res


def lm_calc_fdiff(model, pixels, psf_pars, prior, pars):
    psf_pars = _f8(psf_pars)
    pars = _f8(pars)
    nprior = 0 if not prior.present else (7 if model == "bd" else 6)
    m = pixels.size + nprior
    fdiff = np.zeros(m)
    st = lib().ora_lm_calc_fdiff(
        C.c_int(MODELS[model]), _p(pixels), C.c_int(pixels.size),
        _p(psf_pars), C.c_int(psf_pars.size // 6), C.byref(prior), _p(pars),
        _p(fdiff), C.c_int(m))
    return fdiff, st


def lm_calc_jacobian(model, pixels, psf_pars, prior, pars):
    psf_pars = _f8(psf_pars)
    pars = _f8(pars)
    nprior = 0 if not prior.present else (7 if model == "bd" else 6)
    m = pixels.size + nprior
    jac = np.zeros((pars.size, m))          # column major m x npars
    st = lib().ora_lm_calc_jacobian(
        C.c_int(MODELS[model]), _p(pixels), C.c_int(pixels.size),
        _p(psf_pars), C.c_int(psf_pars.size // 6), C.byref(prior), _p(pars),
        _p(jac), C.c_int(m))
    return jac.T.copy(), st


# ------------------------------------------------------------------ EM
EM_VARIANTS = {"full": 0, "fixcen": 1, "fixcov": 2, "fluxonly": 3}


def em_go(image, weight, jac, guess_pars, psf_pars=None, variant="full",
          tol=1.0e-5, miniter=None, maxiter=500, vary_sky=False, sky=None,
          ignore_zero_weight=True):
    image, weight, jac = _f8(image), _f8(weight), _f8(jac)
    guess_pars = _f8(guess_pars).ravel()
    ngauss = guess_pars.size // 6
    if miniter is None:
        miniter = 20 if variant == "fluxonly" else 40
    if psf_pars is None:
        npsf, psf_ptr = 0, None
    else:
        psf_pars = _f8(psf_pars).ravel()
        npsf, psf_ptr = psf_pars.size // 6, _p(psf_pars)
    pars_out = np.zeros(6 * ngauss)
    res = np.zeros(1, dtype=EM_RESULT_DTYPE)
    lib().ora_em_go(
        C.c_int(EM_VARIANTS[variant]), _p(image), _p(weight),
        C.c_int(image.shape[0]), C.c_int(image.shape[1]), _p(jac),
        _p(guess_pars), C.c_int(ngauss), psf_ptr, C.c_int(npsf),
        C.c_double(tol), C.c_int(miniter), C.c_int(maxiter),
        C.c_int(int(vary_sky)), C.c_int(int(sky is not None)),
        C.c_double(0.0 if sky is None else sky),
        C.c_int(int(ignore_zero_weight)), _p(pars_out), _p(res))
    return res[0], pars_out


def em_go_batch(images, weights, jacs, guess_pars, nthreads=1, tol=1.0e-5,
                miniter=40, maxiter=500):
    images, weights, jacs = _f8(images), _f8(weights), _f8(jacs)
    guess_pars = _f8(guess_pars)
    n, H, W = images.shape
    ngauss = guess_pars.shape[1]
    pars_out = np.zeros((n, ngauss, 6))
    res = np.zeros(n, dtype=EM_RESULT_DTYPE)
    lib().ora_em_go_batch(
        C.c_int(n), C.c_int(nthreads), C.c_int(0), _p(images), _p(weights),
        C.c_int(H), C.c_int(W), _p(jacs), _p(guess_pars), C.c_int(ngauss),
        C.c_double(tol), C.c_int(miniter), C.c_int(maxiter), _p(pars_out),
        _p(res))
    return res, pars_out


# ------------------------------------------------------------------ admom
def admom_conf(maxiter=200, shiftmax=5.0, etol=1.0e-5, Ttol=1.0e-3,
               cenonly=False):
    conf = np.zeros(1, dtype=ADMOM_CONF_DTYPE)
    conf["maxiter"] = maxiter
    conf["shiftmax"] = shiftmax
    conf["etol"] = etol
    conf["Ttol"] = Ttol
    conf["cenonly"] = int(cenonly)
    return conf


def admom_go(image, weight, jac, guess_pars, **kw):
    image, weight, jac = _f8(image), _f8(weight), _f8(jac)
    guess_pars = _f8(guess_pars)
    conf = admom_conf(**kw)
    res = np.zeros(1, dtype=ADMOM_RESULT_DTYPE)
    norm = np.zeros(1)
    st = lib().ora_admom_go(_p(image), _p(weight), C.c_int(image.shape[0]),
                            C.c_int(image.shape[1]), _p(jac), _p(guess_pars),
                            _p(conf), _p(res), _p(norm))
    assert st == 0
    return res[0], norm[0]


def admom_go_batch(images, weights, jacs, guess_pars, nthreads=1, **kw):
    images, weights, jacs = _f8(images), _f8(weights), _f8(jacs)
    guess_pars = _f8(guess_pars)
    n, H, W = images.shape
    conf = admom_conf(**kw)
    res = np.zeros(n, dtype=ADMOM_RESULT_DTYPE)
    lib().ora_admom_go_batch(C.c_int(n), C.c_int(nthreads), _p(images),
                             _p(weights), C.c_int(H), C.c_int(W), _p(jacs),
                             _p(guess_pars), _p(conf), _p(res))
    return res


# ------------------------------------------------------------ weighted moments
MOM_RESULT_DTYPE = np.dtype([
    ("flags", "i4"), ("npix", "i4"), ("wsum", "f8"), ("sums", "f8", 17),
    ("sums_cov", "f8", 17 * 17)], align=True)


def weighted_sums(image, weight, jac, wt_pars, maxrad, nmom=6,
                  ignore_zero_weight=True):
    """get_weighted_sums / get_higher_order_weighted_sums
    (gmix_nb.py:681-821) from image/weight/jacobian with the weight mixture
    given as full parameters; returns dict(sums, sums_cov, wsum, npix)"""
    image, weight, jac = _f8(image), _f8(weight), _f8(jac)
    wt_pars = _f8(wt_pars).ravel()
    res = np.zeros(1, dtype=MOM_RESULT_DTYPE)
    st = lib().ora_weighted_sums_go(
        _p(image), _p(weight), C.c_int(image.shape[0]),
        C.c_int(image.shape[1]), _p(jac), _p(wt_pars),
        C.c_int(wt_pars.size // 6), C.c_double(maxrad), C.c_int(nmom),
        C.c_int(int(ignore_zero_weight)), _p(res))
    assert st == 0
    r = res[0]
    return dict(sums=r["sums"][:nmom].copy(),
                sums_cov=r["sums_cov"][:nmom * nmom].reshape(nmom, nmom).copy(),
                wsum=float(r["wsum"]), npix=int(r["npix"]))


# ------------------------------------------------- multi-epoch / multi-band LM
def lm_fit_multi(model, images, weights, jacs, psf_pars, bands, nband, guess,
                 prior=None, Fpriors=None, maxfev=4000, ftol=1.0e-5,
                 xtol=1.0e-5, ignore_zero_weight=True,
                 analytic_jacobian=True, bounds=None):
    """Fitter.go on the stamps of ONE object (ObsList / MultiBandObsList):
    images [ns,H,W] ordered band by band, bands [ns] int, guess [5+nband],
    Fpriors a list of nband Prior1D (used with `prior`)"""
    images, weights, jacs = _f8(images), _f8(weights), _f8(jacs)
    guess = _f8(guess)
    ns, H, W = images.shape
    bands = np.ascontiguousarray(bands, dtype="i4")
    if psf_pars is None:
        npsf, psf_ptr = 0, None
    else:
        psf_pars = _f8(psf_pars).reshape(ns, -1, 6)
        npsf, psf_ptr = psf_pars.shape[1], _p(psf_pars)
    pr = prior if prior is not None else make_prior("none")
    fp = (Prior1D * nband)()
    for b in range(nband):
        fp[b] = Fpriors[b] if Fpriors is not None else pr.F
    res = np.zeros(1, dtype=LM_RESULT_DTYPE)
    bnd = make_bounds(bounds)
    st = lib().ora_lm_fit_multi_bounded(
        C.c_int(MODELS[model]), C.c_int(ns), _p(images), _p(weights),
        C.c_int(H), C.c_int(W), _p(jacs), psf_ptr, C.c_int(npsf), _p(bands),
        C.c_int(nband), _p(guess), C.byref(pr), fp, C.byref(bnd),
        C.c_int(maxfev),
        C.c_double(ftol), C.c_double(xtol), C.c_int(int(analytic_jacobian)),
        C.c_int(int(ignore_zero_weight)), _p(res))
    assert st == 0, st
    return res[0]


# ------------------------------------------------------------ template flux
FLUX_RESULT_DTYPE = np.dtype([("flags", "i4"), ("pad", "i4"), ("flux", "f8"),
                              ("flux_err", "f8"), ("chi2per", "f8"),
                              ("dof", "f8")])


def template_flux(images, weights, jacs, gm_pars, normalize=True,
                  ignore_zero_weight=True):
    """PSFFluxFitModel.go on the stamps of one object; gm_pars
    [ns, ngauss, 6]"""
    images, weights, jacs = _f8(images), _f8(weights), _f8(jacs)
    if images.ndim == 2:
        images, weights = images[None], weights[None]
    ns, H, W = images.shape
    gm_pars = _f8(gm_pars).reshape(ns, -1, 6)
    res = np.zeros(1, dtype=FLUX_RESULT_DTYPE)
    st = lib().ora_template_flux(
        C.c_int(ns), _p(images), _p(weights), C.c_int(H), C.c_int(W),
        _p(jacs), _p(gm_pars), C.c_int(gm_pars.shape[1]),
        C.c_int(int(normalize)), C.c_int(int(ignore_zero_weight)), _p(res))
    assert st == 0, st
    return res[0]


# ------------------------------------------------------------------ noise cov
def _full_pars(gm):
    return np.array([[gm[k][i] for k in ("p", "row", "col", "irr", "irc",
                                         "icc")] for i in range(gm.size)])


def noise_cov(model, pars, pars_cov0, weight, noise, jac, psf_pars=None):
    """calc_noise_cov (ngmix/fitting/noise_cov.py:90-137), single epoch and
    band, restated with numpy's fft exactly as the reference writes it:
    B_ab = sum_q conj(G_a) G_b |fft2(noise)|^2 / N^2, G_a = fft2(weight *
    dmodel/dp_a); returns pars_cov0 B pars_cov0 and B.  Derivative images:
    deriv_images for gauss/exp/dev (:160-201), central differences of fast
    renders otherwise (:203-226, steps fitting/results.py:927-952)."""
    pars = _f8(pars)
    weight = _f8(weight)
    noise = _f8(noise)
    jac = _f8(jac)
    dims = weight.shape
    npars = pars.size
    psf = None
    if psf_pars is not None:
        psf, _ = gmix_fill("full", _f8(psf_pars).ravel())

    def composed(p):
        gm0, st = gmix_fill(model, p)
        assert st == 0
        return gm0, (gmix_convolve(gm0, psf) if psf is not None else gm0)

    if model in ("gauss", "exp", "dev") and pars[5] != 0.0:
        gm0, gmc = composed(pars)
        gpars = _full_pars(gmc)
        modpars = _full_pars(gm0)
        npsf = gpars.shape[0] // modpars.shape[0]
        modcov = np.repeat(modpars[:, 3:6], npsf, axis=0)
        g1, g2, T, flux = pars[2:6]
        gsq = g1 * g1 + g2 * g2
        f = 2.0 / (1.0 + gsq)
        dfac = -f / (1.0 + gsq)
        de1dg1 = f + 2.0 * g1 * g1 * dfac
        de1dg2 = 2.0 * g1 * g2 * dfac
        de2dg1 = de1dg2
        de2dg2 = f + 2.0 * g2 * g2 * dfac
        Tk = modcov[:, 0] + modcov[:, 2]
        dcov = np.zeros((gpars.shape[0], 3, 3))
        for i, (de1, de2) in enumerate(((de1dg1, de2dg1), (de1dg2, de2dg2))):
            dcov[:, i, 0] = -0.5 * Tk * de1
            dcov[:, i, 1] = 0.5 * Tk * de2
            dcov[:, i, 2] = 0.5 * Tk * de1
        dcov[:, 2, :] = modcov / T
        coords = make_coords(dims, jac)
        out = deriv_images(gpars, dcov, coords["v"], coords["u"],
                           coords["area"])
        images = [out[k].reshape(dims) for k in (1, 2, 3, 4, 5)]
        images.append(out[0].reshape(dims) / flux)
    else:
        images = []
        for ipar in range(npars):
            if ipar < 2:
                step = 1.0e-3
            elif ipar < 4:
                step = 1.0e-4
            elif ipar < npars - 1:
                step = max(1.0e-4, 1.0e-3 * abs(pars[ipar]))
            else:
                step = max(1.0e-6, 1.0e-3 * abs(pars[ipar]))
            ims = []
            for sign in (1, -1):
                p = pars.copy()
                p[ipar] += sign * step
                _, gm = composed(p)
                im, st = render(gm, dims, jac, fast_exp=True)
                assert st == 0
                ims.append(im)
            images.append((ims[0] - ims[1]) / (2 * step))
    kernels = [np.fft.fft2(weight * dim) for dim in images]
    power = np.abs(np.fft.fft2(noise)) ** 2
    n = weight.size
    B = np.zeros((npars, npars))
    for ia in range(npars):
        for ib in range(ia, npars):
            val = np.sum(np.conj(kernels[ia]) * kernels[ib] * power).real \
                / n ** 2
            B[ia, ib] = val
            B[ib, ia] = val
    pars_cov0 = _f8(pars_cov0)
    return pars_cov0 @ B @ pars_cov0, B
