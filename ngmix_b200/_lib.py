"""
ctypes binding of the engine's C ABI (include/ngx.h, built into
ngmix_b200/csrc/libngx.so by ngmix_b200/csrc/Makefile).

There is NO CPU fallback: if the library is missing, or no CUDA device is
visible when a compute entry point is called, an error is raised.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBPATH = os.path.join(CSRC, "libngx.so")


class NgxError(RuntimeError):
    """API misuse or CUDA failure reported by libngx"""


# ---- record layouts (byte-identical to the reference's numpy dtypes)
GAUSS_DTYPE = np.dtype([
    ("p", "f8"), ("row", "f8"), ("col", "f8"), ("irr", "f8"), ("irc", "f8"),
    ("icc", "f8"), ("det", "f8"), ("norm_set", "i8"), ("drr", "f8"),
    ("drc", "f8"), ("dcc", "f8"), ("norm", "f8"), ("pnorm", "f8")])
PIXEL_DTYPE = np.dtype([
    ("u", "f8"), ("v", "f8"), ("area", "f8"), ("val", "f8"), ("ierr", "f8"),
    ("fdiff", "f8")])
COORD_DTYPE = np.dtype([("u", "f8"), ("v", "f8"), ("area", "f8")])
JACOBIAN_DTYPE = np.dtype([
    ("row0", "f8"), ("col0", "f8"), ("dvdrow", "f8"), ("dvdcol", "f8"),
    ("dudrow", "f8"), ("dudcol", "f8"), ("det", "f8"), ("scale", "f8")])

LM_RESULT_DTYPE = np.dtype([
    ("flags", "i4"), ("nfev", "i4"), ("ier", "i4"), ("npix", "i4"),
    ("dof", "i4"), ("njev", "i4"), ("lnprob", "f8"), ("s2n_numer", "f8"),
    ("s2n_denom", "f8"), ("chi2per", "f8"), ("s2n", "f8"), ("T", "f8"),
    ("T_err", "f8"), ("flux", "f8"), ("flux_err", "f8"), ("pars", "f8", 8),
    ("pars_err", "f8", 8), ("pars_cov0", "f8", 64), ("pars_cov", "f8", 64)])
# ngx_lm_result_wide: the same record with 14 parameter slots (4-7 bands,
# coelliptical fits)
LM_RESULT_WIDE_DTYPE = np.dtype(
    [(n, LM_RESULT_DTYPE.fields[n][0]) for n in LM_RESULT_DTYPE.names[:15]]
    + [("pars", "f8", 14), ("pars_err", "f8", 14), ("pars_cov0", "f8", 196),
       ("pars_cov", "f8", 196)])
# ngx_lm_result_xwide: 20 parameter slots (8-12 bands)
LM_RESULT_XWIDE_DTYPE = np.dtype(
    [(n, LM_RESULT_DTYPE.fields[n][0]) for n in LM_RESULT_DTYPE.names[:15]]
    + [("pars", "f8", 20), ("pars_err", "f8", 20), ("pars_cov0", "f8", 400),
       ("pars_cov", "f8", 400)])
EM_RESULT_DTYPE = np.dtype([("flags", "i4"), ("numiter", "i4"),
                            ("fdiff", "f8"), ("sky", "f8")])
ADMOM_RESULT_DTYPE = np.dtype([
    ("flags", "i4"), ("numiter", "i4"), ("npix", "i4"), ("pad", "i4"),
    ("wsum", "f8"), ("sums", "f8", 7), ("sums_cov", "f8", (7, 7)),
    ("pars", "f8", 6), ("rho4", "f8"), ("F", "f8", 7), ("wt_norm", "f8")])

assert GAUSS_DTYPE.itemsize == 104 and PIXEL_DTYPE.itemsize == 48
assert LM_RESULT_DTYPE.itemsize == 1248 and ADMOM_RESULT_DTYPE.itemsize == 592

MODELS = {"full": 0, "gauss": 1, "turb": 2, "exp": 3, "dev": 4, "bdf": 6,
          "coellip": 7, "cm": 9, "bd": 10}
# parameters of the models ngx_lm_fit takes (gmix.py model tables)
MODEL_NPARS = {"gauss": 6, "turb": 6, "exp": 6, "dev": 6, "bdf": 7, "bd": 8}
EM_VARIANTS = {"full": 0, "fixcen": 1, "fixcov": 2, "fluxonly": 3}

ST_OK, ST_RANGE_ERROR, ST_ZERO_DIV, ST_FATAL = 0, 1, 2, 3


class Prior1D(C.Structure):
    _fields_ = [("kind", C.c_int32), ("pad", C.c_int32), ("a", C.c_double),
                ("b", C.c_double), ("c", C.c_double), ("d", C.c_double)]


class Prior(C.Structure):
    _fields_ = [("present", C.c_int32), ("g_kind", C.c_int32),
                ("cen1", C.c_double), ("cen2", C.c_double),
                ("sigma1", C.c_double), ("sigma2", C.c_double),
                ("g_sigma", C.c_double), ("T", Prior1D),
                ("fracdev", Prior1D), ("F", Prior1D),
                ("nband", C.c_int32), ("ncoellip", C.c_int32),
                ("F_more", Prior1D * 2), ("logTratio", Prior1D),
                ("F_extra", Prior1D * 4), ("F_xwide", Prior1D * 5)]


class LMPars(C.Structure):
    _fields_ = [("maxfev", C.c_int32), ("analytic_jacobian", C.c_int32),
                ("ftol", C.c_double), ("xtol", C.c_double),
                ("nbounds", C.c_int32), ("btype", C.c_int32 * 20),
                ("pad_", C.c_int32),
                ("blo", C.c_double * 20), ("bhi", C.c_double * 20)]

    def set_bounds(self, bounds):
        """bounds: None or [(min | None, max | None)] per parameter, the
        reference's prior.bounds (joint_prior.py:37-61)"""
        self.nbounds = 0
        if bounds is None:
            return self
        if len(bounds) > 20:
            raise ValueError("at most 20 parameters")
        self.nbounds = len(bounds)
        for i, (lo, hi) in enumerate(bounds):
            self.btype[i] = (1 if lo is not None else 0) \
                + (2 if hi is not None else 0)
            self.blo[i] = 0.0 if lo is None else lo
            self.bhi[i] = 0.0 if hi is None else hi
        return self


class EMConf(C.Structure):
    _fields_ = [("tol", C.c_double), ("maxiter", C.c_int32),
                ("miniter", C.c_int32), ("sky", C.c_double),
                ("vary_sky", C.c_int32), ("have_sky", C.c_int32)]


class AdmomConf(C.Structure):
    _fields_ = [("maxiter", C.c_int32), ("cenonly", C.c_int32),
                ("shiftmax", C.c_double), ("etol", C.c_double),
                ("Ttol", C.c_double)]


# every symbol include/ngx.h declares
SYMBOLS = [
    "ngx_version", "ngx_last_error", "ngx_init", "ngx_device_info",
    "ngx_gmix_fill", "ngx_gmix_fill_cm", "ngx_gmix_set_norms",
    "ngx_gmix_convolve", "ngx_make_pixels", "ngx_make_coords", "ngx_render",
    "ngx_loglike", "ngx_loglike_pixels", "ngx_deriv_images", "ngx_lm_fit",
    "ngx_em", "ngx_admom",
    "ngx_gmix_fill_host", "ngx_gmix_fill_cm_host", "ngx_gmix_set_norms_host",
    "ngx_gmix_convolve_host", "ngx_make_pixels_host", "ngx_make_coords_host",
    "ngx_render_host", "ngx_loglike_host", "ngx_loglike_pixels_host",
    "ngx_deriv_images_host", "ngx_lm_fit_host", "ngx_em_host",
    "ngx_admom_host", "ngx_peak_dfma", "ngx_peak_dmma",
    "ngx_weighted_sums", "ngx_weighted_sums_host",
    "ngx_lm_fit_multi", "ngx_lm_fit_multi_host",
    "ngx_template_flux", "ngx_template_flux_host",
    "ngx_noise_b", "ngx_noise_b_host", "ngx_noise_cov_host",
    "ngx_lm_fit_host_f32", "ngx_lm_fit_host_compact",
    "ngx_host_alloc", "ngx_host_free", "ngx_host_register",
    "ngx_host_unregister", "ngx_bootstrap", "ngx_bootstrap_workspace_bytes",
    "ngx_chain_draws", "ngx_em_todo", "ngx_lm_fit_todo",
    "ngx_template_flux_images", "ngx_template_flux_images_host",
    "ngx_expand_compact",
    "ngx_lm_fit_multi_wide", "ngx_lm_fit_multi_wide_host",
    "ngx_lm_fit_multi_xwide", "ngx_lm_fit_multi_xwide_host",
    "ngx_lm_fit_coellip", "ngx_lm_fit_coellip_host",
]

_lib = None
_inited_device = None


def build(force=False):
    """compile libngx.so for sm_100a (nvcc cross-compiles without a GPU)"""
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC)
            if f.endswith((".cu", ".cuh"))]
    srcs.append(os.path.join(os.path.dirname(HERE), "include", "ngx.h"))
    if not force and os.path.exists(LIBPATH):
        mt = os.path.getmtime(LIBPATH)
        if all(os.path.getmtime(s) <= mt for s in srcs):
            return LIBPATH
    subprocess.check_call(["make", "-j", str(min(8, os.cpu_count() or 1)),
                           "-C", CSRC, "libngx.so"])
    return LIBPATH


def lib():
    """load libngx.so; raises if it is missing (no fallback)"""
    global _lib
    if _lib is None:
        if not os.path.exists(LIBPATH):
            raise NgxError(
                "libngx.so is not built (%s); run `python -c 'import "
                "__graft_entry__ as g; g.build()'` or `make -C "
                "ngmix_b200/csrc`.  There is no CPU fallback." % LIBPATH)
        _lib = C.CDLL(LIBPATH)
        _lib.ngx_last_error.restype = C.c_char_p
        for name in SYMBOLS:
            getattr(_lib, name)      # AttributeError if a symbol is missing
    return _lib


def check(rc):
    if rc != 0:
        raise NgxError("libngx error %d: %s"
                       % (rc, lib().ngx_last_error().decode()))


def init(device=None):
    """select the CUDA device; raises NgxError when there is none"""
    global _inited_device
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0")) \
            if _inited_device is None else _inited_device
    if _inited_device != device:
        check(lib().ngx_init(C.c_int(device)))
        _inited_device = device
    return device


def _p(a):
    if a is None:
        return None
    return a.ctypes.data_as(C.c_void_p)


def _f8(a):
    return np.ascontiguousarray(a, dtype="f8")


def as_jac(jac, n=None):
    """[n,8] float or structured jacobian records -> contiguous f8 [n,8]"""
    if isinstance(jac, np.ndarray) and jac.dtype == JACOBIAN_DTYPE:
        jac = jac.view("f8").reshape(-1, 8)
    jac = _f8(jac).reshape(-1, 8)
    if n is not None:
        assert jac.shape[0] == n, "one jacobian per object"
    return jac


# ---------------------------------------------------------------- mixtures
def gmix_fill(model, pars):
    """pars [n, npars] -> (gmix [n, ngauss] GAUSS_DTYPE, status [n])"""
    init()
    pars = np.atleast_2d(_f8(pars))
    n, npars = pars.shape
    if model == "coellip":
        ngauss = (npars - 4) // 2
    elif model == "full":
        ngauss = npars // 6
    else:
        ngauss = {"gauss": 1, "turb": 3, "exp": 6, "dev": 10, "bdf": 16,
                  "bd": 16}[model]
    gm = np.zeros((n, ngauss), dtype=GAUSS_DTYPE)
    st = np.zeros(n, dtype="i4")
    check(lib().ngx_gmix_fill_host(C.c_int(n), C.c_int(MODELS[model]),
                                   C.c_int(ngauss), _p(pars), C.c_int(npars),
                                   _p(gm), _p(st)))
    return gm, st


def gmix_fill_cm(cm, pars):
    init()
    cm = np.atleast_2d(_f8(cm))
    pars = np.atleast_2d(_f8(pars))
    n = pars.shape[0]
    gm = np.zeros((n, 16), dtype=GAUSS_DTYPE)
    st = np.zeros(n, dtype="i4")
    check(lib().ngx_gmix_fill_cm_host(C.c_int(n), _p(cm), _p(pars), _p(gm),
                                      _p(st)))
    return gm, st


def gmix_set_norms(gm):
    """in place on gm [n, ngauss]; returns status [n]"""
    init()
    assert gm.dtype == GAUSS_DTYPE and gm.flags.c_contiguous
    g2 = gm.reshape(-1, gm.shape[-1])
    st = np.zeros(g2.shape[0], dtype="i4")
    check(lib().ngx_gmix_set_norms_host(C.c_int(g2.shape[0]),
                                        C.c_int(g2.shape[1]), _p(g2), _p(st)))
    return st


def gmix_convolve(obj, psf):
    init()
    obj = np.ascontiguousarray(obj).reshape(-1, obj.shape[-1])
    psf = np.ascontiguousarray(psf).reshape(-1, psf.shape[-1])
    n = obj.shape[0]
    out = np.zeros((n, obj.shape[1] * psf.shape[1]), dtype=GAUSS_DTYPE)
    check(lib().ngx_gmix_convolve_host(C.c_int(n), C.c_int(obj.shape[1]),
                                       C.c_int(psf.shape[1]), _p(obj),
                                       _p(psf), _p(out)))
    return out


# ---------------------------------------------------------------- pixels
def make_pixels(images, weights, jac, ignore_zero_weight=True):
    """images/weights [n,H,W] -> (pixels [n,H*W] PIXEL_DTYPE, npix [n])"""
    init()
    images = _f8(images)
    weights = _f8(weights)
    if images.ndim == 2:
        images, weights = images[None], weights[None]
    n, H, W = images.shape
    jac = as_jac(jac, n)
    pixels = np.zeros((n, H * W), dtype=PIXEL_DTYPE)
    npix = np.zeros(n, dtype="i4")
    check(lib().ngx_make_pixels_host(
        C.c_int(n), C.c_int(H), C.c_int(W), _p(images), _p(weights), _p(jac),
        C.c_int(int(ignore_zero_weight)), _p(pixels), _p(npix)))
    return pixels, npix


def make_coords(dims, jac):
    init()
    jac = as_jac(jac)
    n = jac.shape[0]
    coords = np.zeros((n, dims[0] * dims[1]), dtype=COORD_DTYPE)
    check(lib().ngx_make_coords_host(C.c_int(n), C.c_int(dims[0]),
                                     C.c_int(dims[1]), _p(jac), _p(coords)))
    return coords


# ---------------------------------------------------------------- passes
def render(gm, jac, dims, fast_exp=False, images=None):
    """adds gm [n, ngauss] into images [n,H,W] (zeros if None)"""
    init()
    gm = np.ascontiguousarray(gm).reshape(-1, gm.shape[-1])
    n = gm.shape[0]
    jac = as_jac(jac, n)
    if images is None:
        images = np.zeros((n, dims[0], dims[1]))
    else:
        images = _f8(images).reshape(n, dims[0], dims[1])
    st = np.zeros(n, dtype="i4")
    check(lib().ngx_render_host(C.c_int(n), C.c_int(gm.shape[1]), _p(gm),
                                _p(jac), C.c_int(dims[0]), C.c_int(dims[1]),
                                C.c_int(int(fast_exp)), _p(images), _p(st)))
    return images, st


def loglike(gm, images, weights, jac, ignore_zero_weight=True,
            want_fdiff=False):
    """fused fdiff/loglike/s2n pass; returns (out4 [n,4], fdiff|None, status)"""
    init()
    gm = np.ascontiguousarray(gm).reshape(-1, gm.shape[-1])
    n = gm.shape[0]
    images = _f8(images).reshape(n, *np.shape(images)[-2:])
    weights = _f8(weights).reshape(images.shape)
    H, W = images.shape[1:]
    jac = as_jac(jac, n)
    out4 = np.zeros((n, 4))
    fdiff = np.zeros((n, H * W)) if want_fdiff else None
    st = np.zeros(n, dtype="i4")
    check(lib().ngx_loglike_host(
        C.c_int(n), C.c_int(gm.shape[1]), _p(gm), _p(images), _p(weights),
        _p(jac), C.c_int(H), C.c_int(W), C.c_int(int(ignore_zero_weight)),
        _p(out4), _p(fdiff), _p(st)))
    return out4, fdiff, st


def loglike_pixels(gm, pixels, npix=None, want_fdiff=False):
    """same pass on pre-built pixel records [n, maxpix]"""
    init()
    gm = np.ascontiguousarray(gm).reshape(-1, gm.shape[-1])
    n = gm.shape[0]
    pixels = np.ascontiguousarray(pixels).reshape(n, -1)
    maxpix = pixels.shape[1]
    if npix is None:
        npix = np.full(n, maxpix, dtype="i4")
    npix = np.ascontiguousarray(npix, dtype="i4")
    out4 = np.zeros((n, 4))
    fdiff = np.zeros((n, maxpix)) if want_fdiff else None
    st = np.zeros(n, dtype="i4")
    check(lib().ngx_loglike_pixels_host(
        C.c_int(n), C.c_int(gm.shape[1]), _p(gm), _p(pixels), C.c_int(maxpix),
        _p(npix), _p(out4), _p(fdiff), _p(st)))
    return out4, fdiff, st


def deriv_images(gpars, dcov, jac, dims):
    init()
    gpars = _f8(gpars)
    if gpars.ndim == 2:
        gpars = gpars[None]
    n, ngauss = gpars.shape[:2]
    dcov = _f8(dcov).reshape(n, ngauss, 3, 3)
    jac = as_jac(jac, n)
    out = np.zeros((n, 6, dims[0] * dims[1]))
    check(lib().ngx_deriv_images_host(C.c_int(n), C.c_int(ngauss), _p(gpars),
                                      _p(dcov), _p(jac), C.c_int(dims[0]),
                                      C.c_int(dims[1]), _p(out)))
    return out


def noise_b(dimages, weights, noise):
    """B [n,nk,nk] of the noise-power sandwich covariance from caller-made
    derivative images [n,nk,H*W] (ngx_noise_b_host)"""
    init()
    dimages = _f8(dimages)
    n, nk = dimages.shape[:2]
    weights = _f8(weights).reshape(n, *np.shape(weights)[-2:])
    noise = _f8(noise).reshape(weights.shape)
    H, W = weights.shape[1:]
    assert dimages.shape[2] == H * W
    B = np.zeros((n, nk, nk))
    check(lib().ngx_noise_b_host(C.c_int(n), C.c_int(nk), _p(dimages),
                                 _p(weights), _p(noise), C.c_int(H),
                                 C.c_int(W), _p(B)))
    return B


def noise_cov_b(gpars, dcov, flux, jac, weights, noise):
    """B [n,6,6] for the models with analytic derivative images
    (ngx_noise_cov_host); gpars [n,ngauss,6], dcov [n,ngauss,3,3]"""
    init()
    gpars = _f8(gpars)
    n, ngauss = gpars.shape[:2]
    dcov = _f8(dcov).reshape(n, ngauss, 3, 3)
    flux = _f8(flux).reshape(n)
    jac = as_jac(jac, n)
    weights = _f8(weights).reshape(n, *np.shape(weights)[-2:])
    noise = _f8(noise).reshape(weights.shape)
    H, W = weights.shape[1:]
    B = np.zeros((n, 6, 6))
    check(lib().ngx_noise_cov_host(C.c_int(n), C.c_int(ngauss), _p(gpars),
                                   _p(dcov), _p(flux), _p(jac), _p(weights),
                                   _p(noise), C.c_int(H), C.c_int(W), _p(B)))
    return B


# ---------------------------------------------------------------- LM
def lm_fit(model, images, weights, jac, psf_pars, guess, prior=None,
           maxfev=4000, ftol=1.0e-5, xtol=1.0e-5, analytic_jacobian=True,
           ignore_zero_weight=True, out=None, bounds=None):
    """
    batched Fitter.go through the host-buffer entry point.
    images/weights [n,H,W]; psf_pars [n,npsf,6] or None; guess [n,npars];
    bounds: None or the reference's prior.bounds, [(min | None, max | None)]
    per parameter (leastsqbound's transformed fit).
    Returns LM_RESULT_DTYPE [n] (written into `out` when given, e.g. a view
    of pinned host memory, which lets the D2H copy run asynchronously).
    """
    init()
    # float32 stamps (the dtype survey pixels have on disk) go through the
    # f32 entry point: half the bytes over PCIe, widened on the device
    f32 = getattr(images, "dtype", None) == np.float32 \
        and getattr(weights, "dtype", None) == np.float32
    if f32:
        images = np.ascontiguousarray(images)
        weights = np.ascontiguousarray(weights)
    else:
        images = _f8(images)
        weights = _f8(weights)
    if images.ndim == 2:
        images, weights = images[None], weights[None]
    n, H, W = images.shape
    if weights.shape != images.shape:
        raise ValueError("weights %s and images %s differ in shape"
                         % (weights.shape, images.shape))
    jac = as_jac(jac, n)
    guess = np.atleast_2d(_f8(guess))
    if guess.shape != (n, MODEL_NPARS[model]):
        raise ValueError("guess has shape %s, expected %s for model %s"
                         % (guess.shape, (n, MODEL_NPARS[model]), model))
    if psf_pars is None:
        npsf = 0
    else:
        psf_pars = _f8(psf_pars)
        npsf = psf_pars.shape[1] if psf_pars.ndim == 3 \
            else psf_pars.size // max(6 * n, 1)
        if psf_pars.size != n * npsf * 6:
            raise ValueError("psf_pars has %d values, expected n*npsf*6"
                             % psf_pars.size)
        psf_pars = psf_pars.reshape(n, npsf, 6)
    pr = prior if prior is not None else Prior()
    lmp = LMPars(maxfev, int(analytic_jacobian), ftol, xtol)
    if bounds is not None:
        if len(bounds) != guess.shape[1]:
            raise ValueError("length of bounds != number of parameters")
        lmp.set_bounds(bounds)
    if out is None:
        res = np.empty(n, dtype=LM_RESULT_DTYPE)
    else:
        res = out
        assert res.dtype == LM_RESULT_DTYPE and res.shape == (n,) \
            and res.flags.c_contiguous
    entry = lib().ngx_lm_fit_host_f32 if f32 else lib().ngx_lm_fit_host
    check(entry(
        C.c_int(n), C.c_int(MODELS[model]), _p(images), _p(weights), _p(jac),
        C.c_int(H), C.c_int(W), _p(psf_pars), C.c_int(npsf), _p(guess),
        C.byref(pr), C.byref(lmp), C.c_int(int(ignore_zero_weight)), _p(res)))
    return res


def lm_fit_compact(model, images, weight_scalar, jac, psf_pars, guess,
                   mask=None, prior=None, maxfev=4000, ftol=1.0e-5,
                   xtol=1.0e-5, analytic_jacobian=True, out=None, bounds=None):
    """
    batched Fitter.go from the compact survey format: images [n,H,W] float32,
    weight_scalar [n] (one weight per stamp), mask [n,H,W] bool (True = use
    the pixel; None = all) or already packed bits [n, ceil(H*W/8)] uint8.
    Equivalent to lm_fit(images.astype(f8), where(mask, weight_scalar, 0)).
    """
    init()
    images = np.ascontiguousarray(images, dtype=np.float32)
    if images.ndim == 2:
        images = images[None]
    n, H, W = images.shape
    ws = np.ascontiguousarray(np.broadcast_to(
        np.asarray(weight_scalar, dtype="f8"), (n,)))
    bits = None
    if mask is not None:
        mask = np.asarray(mask)
        if mask.dtype == np.uint8 and mask.ndim == 2:
            bits = np.ascontiguousarray(mask)
        else:
            if mask.shape != images.shape:
                raise ValueError("mask %s and images %s differ in shape"
                                 % (mask.shape, images.shape))
            bits = np.packbits(mask.reshape(n, -1).astype(bool), axis=1)
        if bits.shape != (n, (H * W + 7) // 8):
            raise ValueError("packed mask has shape %s" % (bits.shape,))
    jac = as_jac(jac, n)
    guess = np.atleast_2d(_f8(guess))
    if guess.shape != (n, MODEL_NPARS[model]):
        raise ValueError("guess has shape %s, expected %s for model %s"
                         % (guess.shape, (n, MODEL_NPARS[model]), model))
    if psf_pars is None:
        npsf = 0
    else:
        psf_pars = _f8(psf_pars)
        npsf = psf_pars.shape[1] if psf_pars.ndim == 3 \
            else psf_pars.size // max(6 * n, 1)
        psf_pars = psf_pars.reshape(n, npsf, 6)
    pr = prior if prior is not None else Prior()
    lmp = LMPars(maxfev, int(analytic_jacobian), ftol, xtol)
    lmp.set_bounds(bounds)
    if out is None:
        res = np.empty(n, dtype=LM_RESULT_DTYPE)
    else:
        res = out
        assert res.dtype == LM_RESULT_DTYPE and res.shape == (n,) \
            and res.flags.c_contiguous
    check(lib().ngx_lm_fit_host_compact(
        C.c_int(n), C.c_int(MODELS[model]), _p(images), _p(ws), _p(bits),
        _p(jac), C.c_int(H), C.c_int(W), _p(psf_pars), C.c_int(npsf),
        _p(guess), C.byref(pr), C.byref(lmp), C.c_int(1), _p(res)))
    return res


def lm_fit_coellip(ngauss, images, weights, jac, psf_pars, guess, prior=None,
                   maxfev=4000, ftol=1.0e-5, xtol=1.0e-5,
                   ignore_zero_weight=True):
    """batched CoellipFitter.go: guess [n, 4 + 2 ngauss]; returns
    LM_RESULT_WIDE_DTYPE [n]"""
    init()
    images = _f8(images)
    weights = _f8(weights)
    if images.ndim == 2:
        images, weights = images[None], weights[None]
    n, H, W = images.shape
    jac = as_jac(jac, n)
    guess = np.atleast_2d(_f8(guess))
    assert guess.shape == (n, 4 + 2 * ngauss)
    if psf_pars is None:
        npsf = 0
    else:
        psf_pars = _f8(psf_pars).reshape(n, -1, 6)
        npsf = psf_pars.shape[1]
    pr = prior if prior is not None else Prior()
    lmp = LMPars(maxfev, 0, ftol, xtol)
    res = np.zeros(n, dtype=LM_RESULT_WIDE_DTYPE)
    check(lib().ngx_lm_fit_coellip_host(
        C.c_int(n), C.c_int(ngauss), _p(images), _p(weights), _p(jac),
        C.c_int(H), C.c_int(W), _p(psf_pars), C.c_int(npsf), _p(guess),
        C.byref(pr), C.byref(lmp), C.c_int(int(ignore_zero_weight)), _p(res)))
    return res


def lm_fit_multi(model, stamp_start, stamp_band, nband, images, weights, jac,
                 psf_pars, guess, prior=None, maxfev=4000, ftol=1.0e-5,
                 xtol=1.0e-5, ignore_zero_weight=True,
                 analytic_jacobian=True, bounds=None):
    """
    Fitter.go on ObsList / MultiBandObsList input for a batch of objects:
    object i owns stamps stamp_start[i]:stamp_start[i+1] of images/weights
    [ns,H,W] (ordered band by band), stamp_band [ns]; guess
    [nobj, nshared + nband] with 5 (gauss, exp, dev, turb), 6 (bdf) or 7 (bd)
    shared parameters.  bdf, bd, turb and analytic_jacobian=False run
    MINPACK's forward differences.  Returns LM_RESULT_DTYPE [nobj] for up to
    8 parameters, LM_RESULT_WIDE_DTYPE (14 parameter slots) up to 7 bands and
    14 parameters, LM_RESULT_XWIDE_DTYPE (20 slots) up to 12 bands.
    """
    init()
    images = _f8(images)
    weights = _f8(weights)
    ns, H, W = images.shape
    jac = as_jac(jac, ns)
    stamp_start = np.ascontiguousarray(stamp_start, dtype="i4")
    stamp_band = np.ascontiguousarray(stamp_band, dtype="i4")
    nobj = stamp_start.size - 1
    assert stamp_start[0] == 0 and stamp_start[-1] == ns
    assert stamp_band.size == ns
    npars = MODEL_NPARS[model] - 1 + nband
    if npars > 20 or nband > 12:
        raise ValueError("at most 12 bands and 20 parameters")
    guess = _f8(guess).reshape(nobj, npars)
    if psf_pars is None:
        npsf = 0
    else:
        psf_pars = _f8(psf_pars).reshape(ns, -1, 6)
        npsf = psf_pars.shape[1]
    pr = prior if prior is not None else Prior()
    lmp = LMPars(maxfev, int(analytic_jacobian), ftol, xtol)
    lmp.set_bounds(bounds)
    if npars > 14 or nband > 7:
        dtype, entry = LM_RESULT_XWIDE_DTYPE, lib().ngx_lm_fit_multi_xwide_host
    elif npars > 8:
        dtype, entry = LM_RESULT_WIDE_DTYPE, lib().ngx_lm_fit_multi_wide_host
    else:
        dtype, entry = LM_RESULT_DTYPE, lib().ngx_lm_fit_multi_host
    res = np.zeros(nobj, dtype=dtype)
    check(entry(
        C.c_int(nobj), C.c_int(MODELS[model]), _p(stamp_start),
        _p(stamp_band), C.c_int(nband), _p(images), _p(weights), _p(jac),
        C.c_int(H), C.c_int(W), _p(psf_pars), C.c_int(npsf), _p(guess),
        C.byref(pr), C.byref(lmp), C.c_int(int(ignore_zero_weight)), _p(res)))
    return res


def lm_fit_coellip_device(ngauss, n, d_images, d_weights, d_jac, nrow, ncol,
                          d_psf, npsf, d_guess, d_res, prior=None,
                          maxfev=4000, ftol=1.0e-5, xtol=1.0e-5,
                          ignore_zero_weight=True, stream=0):
    """device-pointer coelliptical fit (asynchronous); d_res holds
    LM_RESULT_WIDE_DTYPE records"""
    pr = prior if prior is not None else Prior()
    pr.ncoellip = ngauss
    lmp = LMPars(maxfev, 0, ftol, xtol)
    check(lib().ngx_lm_fit_coellip(
        C.c_int(n), C.c_int(ngauss), C.c_void_p(d_images),
        C.c_void_p(d_weights), C.c_void_p(d_jac), C.c_int(nrow), C.c_int(ncol),
        C.c_void_p(d_psf) if npsf else None, C.c_int(npsf),
        C.c_void_p(d_guess), C.byref(pr), C.byref(lmp),
        C.c_int(int(ignore_zero_weight)), C.c_void_p(d_res),
        C.c_void_p(stream)))


def lm_fit_multi_device(model, nobj, d_stamp_start, d_stamp_band, nband,
                        d_images, d_weights, d_jac, nrow, ncol, d_psf, npsf,
                        d_guess, d_res, prior=None, maxfev=4000, ftol=1.0e-5,
                        xtol=1.0e-5, ignore_zero_weight=True, stream=0):
    """device-pointer multi-epoch / multi-band fit (asynchronous)"""
    pr = prior if prior is not None else Prior()
    lmp = LMPars(maxfev, 1, ftol, xtol)
    check(lib().ngx_lm_fit_multi(
        C.c_int(nobj), C.c_int(MODELS[model]), C.c_void_p(d_stamp_start),
        C.c_void_p(d_stamp_band), C.c_int(nband), C.c_void_p(d_images),
        C.c_void_p(d_weights), C.c_void_p(d_jac), C.c_int(nrow),
        C.c_int(ncol), C.c_void_p(d_psf) if npsf else None, C.c_int(npsf),
        C.c_void_p(d_guess), C.byref(pr), C.byref(lmp),
        C.c_int(int(ignore_zero_weight)), C.c_void_p(d_res),
        C.c_void_p(stream)))


def lm_fit_device(model, n, d_images, d_weights, d_jac, nrow, ncol, d_psf,
                  npsf, d_guess, d_res, prior=None, maxfev=4000, ftol=1.0e-5,
                  xtol=1.0e-5, analytic_jacobian=True, ignore_zero_weight=True,
                  stream=0, bounds=None):
    """
    device-pointer entry point: d_* are integer device addresses (for example
    torch tensors' data_ptr()); enqueues on `stream` and returns at once.
    """
    pr = prior if prior is not None else Prior()
    lmp = LMPars(maxfev, int(analytic_jacobian), ftol, xtol)
    lmp.set_bounds(bounds)
    check(lib().ngx_lm_fit(
        C.c_int(n), C.c_int(MODELS[model]), C.c_void_p(d_images),
        C.c_void_p(d_weights), C.c_void_p(d_jac), C.c_int(nrow),
        C.c_int(ncol), C.c_void_p(d_psf) if npsf else None, C.c_int(npsf),
        C.c_void_p(d_guess), C.byref(pr), C.byref(lmp),
        C.c_int(int(ignore_zero_weight)), C.c_void_p(d_res),
        C.c_void_p(stream)))


# ---------------------------------------------------------------- EM / admom
def em(images, weights, jac, guess, psf_pars=None, variant="full",
       tol=1.0e-5, miniter=None, maxiter=500, vary_sky=False, sky=None,
       ignore_zero_weight=True):
    """batched EMFitter*.go; guess [n,ngauss,6]; returns (res, pars_out)"""
    init()
    images = _f8(images)
    weights = _f8(weights)
    if images.ndim == 2:
        images, weights = images[None], weights[None]
    n, H, W = images.shape
    jac = as_jac(jac, n)
    guess = _f8(guess)
    ngauss = guess.shape[1] if guess.ndim == 3 \
        else guess.size // max(6 * n, 1)
    guess = guess.reshape(n, ngauss, 6)
    if psf_pars is None:
        npsf = 0
    else:
        psf_pars = _f8(psf_pars)
        npsf = psf_pars.shape[1] if psf_pars.ndim == 3 \
            else psf_pars.size // max(6 * n, 1)
        psf_pars = psf_pars.reshape(n, npsf, 6)
    if miniter is None:
        miniter = 20 if variant == "fluxonly" else 40
    conf = EMConf(tol, maxiter, miniter, 0.0 if sky is None else sky,
                  int(vary_sky), int(sky is not None))
    pars_out = np.zeros((n, ngauss, 6))
    res = np.zeros(n, dtype=EM_RESULT_DTYPE)
    check(lib().ngx_em_host(
        C.c_int(n), C.c_int(EM_VARIANTS[variant]), _p(images), _p(weights),
        _p(jac), C.c_int(H), C.c_int(W), _p(guess), C.c_int(ngauss),
        _p(psf_pars), C.c_int(npsf), C.byref(conf),
        C.c_int(int(ignore_zero_weight)), _p(pars_out), _p(res)))
    return res, pars_out


def admom(images, weights, jac, guess, maxiter=200, shiftmax=5.0,
          etol=1.0e-5, Ttol=1.0e-3, cenonly=False):
    """batched admom; guess [n,6] full pars of the weight gaussian"""
    init()
    images = _f8(images)
    weights = _f8(weights)
    if images.ndim == 2:
        images, weights = images[None], weights[None]
    n, H, W = images.shape
    jac = as_jac(jac, n)
    guess = np.atleast_2d(_f8(guess))
    conf = AdmomConf(maxiter, int(cenonly), shiftmax, etol, Ttol)
    res = np.zeros(n, dtype=ADMOM_RESULT_DTYPE)
    check(lib().ngx_admom_host(C.c_int(n), _p(images), _p(weights), _p(jac),
                               C.c_int(H), C.c_int(W), _p(guess),
                               C.byref(conf), _p(res)))
    return res


MOM_RESULT_DTYPE = np.dtype([
    ("flags", "i4"), ("npix", "i4"), ("wsum", "f8"), ("sums", "f8", 17),
    ("sums_cov", "f8", 17 * 17)], align=True)
assert MOM_RESULT_DTYPE.itemsize == 2464


def weighted_sums(gm, images, weights, jac, maxrad, nmom=6,
                  ignore_zero_weight=True):
    """batched get_weighted_sums / get_higher_order_weighted_sums: gm
    [n, ngauss] weight mixtures (GAUSS_DTYPE), maxrad [n]; returns
    MOM_RESULT_DTYPE [n] (sums_cov is [nmom, nmom] in its leading entries)"""
    init()
    images = _f8(images)
    weights = _f8(weights)
    if images.ndim == 2:
        images, weights = images[None], weights[None]
    n, H, W = images.shape
    jac = as_jac(jac, n)
    gm = np.ascontiguousarray(gm, dtype=GAUSS_DTYPE).reshape(n, -1)
    maxrad = np.ascontiguousarray(np.broadcast_to(_f8(maxrad), (n,)))
    res = np.zeros(n, dtype=MOM_RESULT_DTYPE)
    check(lib().ngx_weighted_sums_host(
        C.c_int(n), C.c_int(gm.shape[1]), _p(gm), _p(images), _p(weights),
        _p(jac), C.c_int(H), C.c_int(W), _p(maxrad), C.c_int(nmom),
        C.c_int(int(ignore_zero_weight)), _p(res)))
    return res


def weighted_sums_device(n, ngauss, d_gmix, d_images, d_weights, d_jac, nrow,
                         ncol, d_maxrad, nmom, d_res, ignore_zero_weight=True,
                         stream=0):
    """device-pointer weighted-moment pass (asynchronous on `stream`)"""
    check(lib().ngx_weighted_sums(
        C.c_int(n), C.c_int(ngauss), C.c_void_p(d_gmix),
        C.c_void_p(d_images), C.c_void_p(d_weights), C.c_void_p(d_jac),
        C.c_int(nrow), C.c_int(ncol), C.c_void_p(d_maxrad), C.c_int(nmom),
        C.c_int(int(ignore_zero_weight)), C.c_void_p(d_res),
        C.c_void_p(stream)))


FLUX_RESULT_DTYPE = np.dtype([("flags", "i4"), ("pad", "i4"), ("flux", "f8"),
                              ("flux_err", "f8"), ("chi2per", "f8"),
                              ("dof", "f8")])


def template_flux(stamp_start, images, weights, jac, gm_pars, normalize=True,
                  ignore_zero_weight=True):
    """batched PSFFluxFitModel.go: object i owns stamps
    stamp_start[i]:stamp_start[i+1]; gm_pars [ns, ngauss, 6] template
    mixtures; returns FLUX_RESULT_DTYPE [nobj]"""
    init()
    images = _f8(images)
    weights = _f8(weights)
    ns, H, W = images.shape
    jac = as_jac(jac, ns)
    stamp_start = np.ascontiguousarray(stamp_start, dtype="i4")
    nobj = stamp_start.size - 1
    assert stamp_start[0] == 0 and stamp_start[-1] == ns
    gm_pars = _f8(gm_pars).reshape(ns, -1, 6)
    res = np.zeros(nobj, dtype=FLUX_RESULT_DTYPE)
    check(lib().ngx_template_flux_host(
        C.c_int(nobj), _p(stamp_start), _p(images), _p(weights), _p(jac),
        C.c_int(H), C.c_int(W), _p(gm_pars), C.c_int(gm_pars.shape[1]),
        C.c_int(int(normalize)), C.c_int(int(ignore_zero_weight)), _p(res)))
    return res


def template_flux_images(stamp_start, images, weights, templates,
                         normalize=True, ignore_zero_weight=True):
    """batched PSFFluxFitModel.go with pixel templates [ns, H, W]"""
    init()
    images, weights, templates = _f8(images), _f8(weights), _f8(templates)
    ns, H, W = images.shape
    if templates.shape != images.shape or weights.shape != images.shape:
        raise ValueError("images, weights and templates differ in shape")
    stamp_start = np.ascontiguousarray(stamp_start, dtype="i4")
    nobj = stamp_start.size - 1
    assert stamp_start[0] == 0 and stamp_start[-1] == ns
    res = np.zeros(nobj, dtype=FLUX_RESULT_DTYPE)
    check(lib().ngx_template_flux_images_host(
        C.c_int(nobj), _p(stamp_start), _p(images), _p(weights),
        _p(templates), C.c_int(H), C.c_int(W), C.c_int(int(normalize)),
        C.c_int(int(ignore_zero_weight)), _p(res)))
    return res


def peak_dfma(iters=20000):
    """measured fp64 FMA lane-ops per second of the current device"""
    init()
    out = C.c_double()
    ms = C.c_double()
    check(lib().ngx_peak_dfma(C.c_int(iters), C.byref(out), C.byref(ms)))
    return out.value, ms.value


def peak_dmma(iters=20000, mode=0):
    """measured warp-level fp64 m8n8k4 MMA instructions per second"""
    init()
    out = C.c_double()
    ms = C.c_double()
    check(lib().ngx_peak_dmma(C.c_int(iters), C.c_int(mode), C.byref(out),
                              C.byref(ms)))
    return out.value, ms.value


def em_device(n, d_images, d_weights, d_jac, nrow, ncol, d_guess, ngauss,
              d_pars_out, d_res, d_psf=0, npsf=0, variant="full", tol=1.0e-5,
              miniter=40, maxiter=500, vary_sky=False, sky=None,
              ignore_zero_weight=True, stream=0):
    """device-pointer EM entry point (asynchronous on `stream`)"""
    conf = EMConf(tol, maxiter, miniter, 0.0 if sky is None else sky,
                  int(vary_sky), int(sky is not None))
    check(lib().ngx_em(
        C.c_int(n), C.c_int(EM_VARIANTS[variant]), C.c_void_p(d_images),
        C.c_void_p(d_weights), C.c_void_p(d_jac), C.c_int(nrow),
        C.c_int(ncol), C.c_void_p(d_guess), C.c_int(ngauss),
        C.c_void_p(d_psf) if npsf else None, C.c_int(npsf), C.byref(conf),
        C.c_int(int(ignore_zero_weight)), C.c_void_p(d_pars_out),
        C.c_void_p(d_res), C.c_void_p(stream)))


def admom_device(n, d_images, d_weights, d_jac, nrow, ncol, d_guess, d_res,
                 maxiter=200, shiftmax=5.0, etol=1.0e-5, Ttol=1.0e-3,
                 cenonly=False, stream=0):
    """device-pointer admom entry point (asynchronous on `stream`)"""
    conf = AdmomConf(maxiter, int(cenonly), shiftmax, etol, Ttol)
    check(lib().ngx_admom(
        C.c_int(n), C.c_void_p(d_images), C.c_void_p(d_weights),
        C.c_void_p(d_jac), C.c_int(nrow), C.c_int(ncol), C.c_void_p(d_guess),
        C.byref(conf), C.c_void_p(d_res), C.c_void_p(stream)))


def loglike_device(n, ngauss, d_gmix, d_images, d_weights, d_jac, nrow, ncol,
                   d_out4, d_fdiff=0, d_status=0, ignore_zero_weight=True,
                   stream=0):
    """device-pointer fused fdiff/loglike/s2n pass (asynchronous)"""
    check(lib().ngx_loglike(
        C.c_int(n), C.c_int(ngauss), C.c_void_p(d_gmix),
        C.c_void_p(d_images), C.c_void_p(d_weights), C.c_void_p(d_jac),
        C.c_int(nrow), C.c_int(ncol), C.c_int(int(ignore_zero_weight)),
        C.c_void_p(d_out4), C.c_void_p(d_fdiff) if d_fdiff else None,
        C.c_void_p(d_status) if d_status else None, C.c_void_p(stream)))


# -------------------------------------------------- device-pointer mixtures
MODEL_NGAUSS = {"gauss": 1, "turb": 3, "exp": 6, "dev": 10, "bdf": 16,
                "bd": 16}


def gmix_fill_device(model, n, d_pars, npars, ngauss, d_gmix, d_status,
                     stream=0):
    """ngx_gmix_fill on device pointers: pars [n, npars] -> gmix [n, ngauss]"""
    check(lib().ngx_gmix_fill(
        C.c_int(n), C.c_int(MODELS[model]), C.c_int(ngauss),
        C.c_void_p(d_pars), C.c_int(npars), C.c_void_p(d_gmix),
        C.c_void_p(d_status), C.c_void_p(stream)))


def gmix_convolve_device(n, nobj, npsf, d_obj, d_psf, d_out, stream=0):
    check(lib().ngx_gmix_convolve(
        C.c_int(n), C.c_int(nobj), C.c_int(npsf), C.c_void_p(d_obj),
        C.c_void_p(d_psf), C.c_void_p(d_out), C.c_void_p(stream)))


def render_device(n, ngauss, d_gmix, d_jac, nrow, ncol, fast_exp, d_images,
                  d_status, stream=0):
    """adds the mixtures into images [n, nrow, ncol] (asynchronous)"""
    check(lib().ngx_render(
        C.c_int(n), C.c_int(ngauss), C.c_void_p(d_gmix), C.c_void_p(d_jac),
        C.c_int(nrow), C.c_int(ncol), C.c_int(int(fast_exp)),
        C.c_void_p(d_images), C.c_void_p(d_status), C.c_void_p(stream)))


# ------------------------------------------------------------ pinned memory
class _PinnedBlock(object):
    """owner of one cudaHostAlloc block; freed when the last array viewing
    it is collected"""

    def __init__(self, nbytes):
        self.ptr = C.c_void_p()
        check(lib().ngx_host_alloc(C.c_size_t(nbytes), C.byref(self.ptr)))
        self.nbytes = nbytes

    def __del__(self):
        try:
            if self.ptr:
                lib().ngx_host_free(self.ptr)
        except Exception:
            pass


def pinned_empty(shape, dtype="f8"):
    """numpy array in page-locked host memory (cudaHostAlloc): the input
    format the host-buffer calls (fit_batch, em_batch, ...) stream from at
    full PCIe rate, with the copy of one chunk overlapping the fit of the
    previous one"""
    init()
    dtype = np.dtype(dtype)
    shape = tuple(np.atleast_1d(shape).astype(int))
    nbytes = int(np.prod(shape)) * dtype.itemsize
    block = _PinnedBlock(max(nbytes, 1))
    buf = (C.c_char * max(nbytes, 1)).from_address(block.ptr.value)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape)))
    arr = arr.reshape(shape)
    _pinned_owners[id(buf)] = block
    import weakref
    weakref.finalize(buf, _pinned_owners.pop, id(buf), None)
    return arr


_pinned_owners = {}


def pinned_copy(a):
    """a copy of `a` in page-locked host memory"""
    a = np.asarray(a)
    out = pinned_empty(a.shape, a.dtype)
    out[...] = a
    return out


class pinned(object):
    """context manager that page-locks arrays the caller already owns
    (cudaHostRegister) for the duration of the block:

        with ngmix_b200.pinned(images, weights):
            recs = ngmix_b200.fit_batch(...)

    Registration costs about as much as one pageable copy, so it pays when
    the arrays are used for more than one call."""

    def __init__(self, *arrays):
        self.arrays = [a for a in arrays if a is not None]
        self.done = []

    def __enter__(self):
        init()
        for a in self.arrays:
            if not a.flags.c_contiguous:
                raise ValueError("only contiguous arrays can be pinned")
            check(lib().ngx_host_register(C.c_void_p(a.ctypes.data),
                                          C.c_size_t(a.nbytes)))
            self.done.append(a)
        return self

    def __exit__(self, *exc):
        for a in self.done:
            lib().ngx_host_unregister(C.c_void_p(a.ctypes.data))
        self.done = []
        return False


# ----------------------------------------------------------- fitting chain


def bootstrap_device(n, d_psf_images, d_psf_weights, d_psf_jac, psf_nrow,
                     psf_ncol, psf_ngauss, psf_ntry, d_images, d_weights,
                     d_jac, nrow, ncol, model, d_T_guess, ntry, seed,
                     d_psf_pars, d_psf_res, d_psf_flux, d_guess, d_res,
                     d_ntry_used, d_workspace, workspace_bytes, prior=None,
                     maxfev=4000, ftol=1.0e-5, xtol=1.0e-5,
                     analytic_jacobian=True, em_tol=1.0e-5, em_miniter=40,
                     em_maxiter=500, bounds=None, stream=0):
    """ngx_bootstrap on device pointers (asynchronous on `stream`)"""
    conf = EMConf(em_tol, em_maxiter, em_miniter, 0.0, 0, 0)
    pr = prior if prior is not None else Prior()
    lmp = LMPars(maxfev, int(analytic_jacobian), ftol, xtol)
    lmp.set_bounds(bounds)
    check(lib().ngx_bootstrap(
        C.c_int(n), C.c_void_p(d_psf_images), C.c_void_p(d_psf_weights),
        C.c_void_p(d_psf_jac), C.c_int(psf_nrow), C.c_int(psf_ncol),
        C.c_int(psf_ngauss), C.c_int(psf_ntry), C.byref(conf),
        C.c_void_p(d_images), C.c_void_p(d_weights), C.c_void_p(d_jac),
        C.c_int(nrow), C.c_int(ncol), C.c_int(MODELS[model]),
        C.c_void_p(d_T_guess) if d_T_guess else None, C.c_int(ntry),
        C.byref(pr), C.byref(lmp), C.c_ulonglong(seed),
        C.c_void_p(d_psf_pars), C.c_void_p(d_psf_res),
        C.c_void_p(d_psf_flux), C.c_void_p(d_guess), C.c_void_p(d_res),
        C.c_void_p(d_ntry_used), C.c_void_p(d_workspace),
        C.c_size_t(workspace_bytes), C.c_void_p(stream)))


def bootstrap_workspace_bytes(n, psf_ngauss):
    f = lib().ngx_bootstrap_workspace_bytes
    f.restype = C.c_size_t
    return int(f(C.c_int(n), C.c_int(psf_ngauss)))


def chain_draws_device(n, seed, tr, stage, ndraw, d_out, stream=0):
    check(lib().ngx_chain_draws(C.c_int(n), C.c_ulonglong(seed), C.c_int(tr),
                                C.c_int(stage), C.c_int(ndraw),
                                C.c_void_p(d_out), C.c_void_p(stream)))


def expand_compact_device(n, nrow, ncol, d_images32, d_wscalar, d_mask,
                          d_images, d_weights, stream=0):
    """ngx_expand_compact on device pointers (asynchronous)"""
    check(lib().ngx_expand_compact(
        C.c_int(n), C.c_int(nrow), C.c_int(ncol), C.c_void_p(d_images32),
        C.c_void_p(d_wscalar), C.c_void_p(d_mask) if d_mask else None,
        C.c_void_p(d_images), C.c_void_p(d_weights), C.c_void_p(stream)))
