"""
Sandwich covariance of LM fits under stationary correlated noise
(ngmix/fitting/noise_cov.py): Cov = pars_cov0 B pars_cov0, with B accumulated
from the noise images attached to the observations.

B is computed on the device (csrc/ngx_noise.cu): the reference's sum over
Fourier modes is, by Parseval, the Gram matrix of the circular
cross-correlations of weight * dmodel/dp_a with the noise image.  The models
with analytic derivative images (gauss, exp, dev) get them from
ngx_deriv_images; the others use the reference's central differences of fast
renders (noise_cov.py:203-226).  The small matrix algebra and the covariance
tests (leastsqbound.py:165-180) run on the host per object.
"""
import numpy as np

from . import _lib
from .defaults import CDEF
from .flags import LM_NEG_COV_EIG, LM_NEG_COV_DIAG, EIG_NOTFINITE

SIMPLE_ANALYTIC_MODELS = ("gauss", "exp", "dev")

# steps of the numerical derivative images, fitting/results.py:927-952
STEP_CEN = 1.0e-3
STEP_SHAPE = 1.0e-4
STEP_STRUCT_MIN = 1.0e-4
STEP_FLUX_MIN = 1.0e-6
STEP_FRAC = 1.0e-3


def get_step(pars, ipar, nband):
    nshape = pars.size - nband
    if ipar < 2:
        return STEP_CEN
    if ipar < 4:
        return STEP_SHAPE
    if ipar < nshape:
        return max(STEP_STRUCT_MIN, STEP_FRAC * abs(pars[ipar]))
    return max(STEP_FLUX_MIN, STEP_FRAC * abs(pars[ipar]))


def _full(gm):
    """[n, ngauss] gaussian records -> [n, ngauss, 6] (p,row,col,irr,irc,icc)"""
    return np.stack([gm[k] for k in ("p", "row", "col", "irr", "irc", "icc")],
                    axis=-1)


def model_deriv_data(model, pars, psf_pars):
    """get_model_deriv_data (fitting/results.py:955-1010) for n objects:
    the composed gaussians [n, ng, 6] and d(irr, irc, icc)/d(g1, g2, T)
    [n, ng, 3, 3]"""
    pars = np.atleast_2d(np.asarray(pars, dtype="f8"))
    gm0, st = _lib.gmix_fill(model, pars)
    if st.any():
        raise ValueError("parameters outside the model's domain")
    if psf_pars is not None:
        n = pars.shape[0]
        psf, _ = _lib.gmix_fill("full", np.asarray(psf_pars).reshape(n, -1))
        gmc = _lib.gmix_convolve(gm0, psf)
        npsf = psf.shape[1]
    else:
        gmc, npsf = gm0, 1
    gpars = _full(gmc)
    modcov = np.repeat(_full(gm0)[:, :, 3:6], npsf, axis=1)
    g1, g2, T = pars[:, 2, None], pars[:, 3, None], pars[:, 4, None]
    gsq = g1 * g1 + g2 * g2
    f = 2.0 / (1.0 + gsq)
    dfac = -f / (1.0 + gsq)
    de1dg1 = f + 2.0 * g1 * g1 * dfac
    de1dg2 = 2.0 * g1 * g2 * dfac
    de2dg1 = de1dg2
    de2dg2 = f + 2.0 * g2 * g2 * dfac
    Tk = modcov[:, :, 0] + modcov[:, :, 2]
    dcov = np.zeros(gpars.shape[:2] + (3, 3))
    for i, (de1, de2) in enumerate(((de1dg1, de2dg1), (de1dg2, de2dg2))):
        dcov[:, :, i, 0] = -0.5 * Tk * de1
        dcov[:, :, i, 1] = 0.5 * Tk * de2
        dcov[:, :, i, 2] = 0.5 * Tk * de1
    dcov[:, :, 2, :] = modcov / T[:, :, None]
    return gpars, dcov


def _fd_images(model, pars, psf_pars, jac, dims, per_object=False):
    """central-difference derivative images of the convolved model with the
    fast exp, one per parameter (noise_cov.py:203-226): [n, npars, H*W].
    A stepped point outside the model's domain raises for a single object
    (as the reference does); with per_object the offender's images become
    NaN instead, so only that object's covariance is flagged."""
    n, npars = pars.shape
    out = np.zeros((n, npars, dims[0] * dims[1]))
    bad = np.zeros(n, dtype=bool)
    psf = None
    if psf_pars is not None:
        psf, _ = _lib.gmix_fill("full", np.asarray(psf_pars).reshape(n, -1))
    for ipar in range(npars):
        step = np.array([get_step(p, ipar, 1) for p in pars])
        ims = []
        for sign in (1.0, -1.0):
            p = pars.copy()
            p[:, ipar] += sign * step
            gm, st = _lib.gmix_fill(model, p)
            if st.any():
                if not per_object:
                    raise ValueError("stepped parameters outside the domain")
                bad |= st != 0
                # any in-domain stand-in keeps the launch going
                p[st != 0] = pars[st != 0]
                gm, _ = _lib.gmix_fill(model, p)
            if psf is not None:
                gm = _lib.gmix_convolve(gm, psf)
            im, st = _lib.render(gm, jac, dims, fast_exp=True)
            if st.any():
                if not per_object:
                    raise ValueError("stepped parameters outside the domain")
                bad |= st != 0
            ims.append(im.reshape(n, -1))
        out[:, ipar] = (ims[0] - ims[1]) / (2 * step[:, None])
    out[bad] = np.nan
    return out


def calc_noise_b(model, pars, psf_pars, jac, weights, noise,
                 per_object=False):
    """B [n, npars, npars] of n single-epoch, single-band fits
    (noise_cov.py:90-137); weights / noise [n, H, W].  per_object: an object
    whose derivative images cannot be formed gets a NaN matrix (flagged by
    sandwich()) instead of aborting the batch."""
    pars = np.atleast_2d(np.asarray(pars, dtype="f8"))
    n, npars = pars.shape
    weights = np.ascontiguousarray(weights, dtype="f8")
    noise = np.ascontiguousarray(noise, dtype="f8")
    dims = weights.shape[-2:]
    jac = _lib.as_jac(jac, n)
    if model in SIMPLE_ANALYTIC_MODELS and np.all(pars[:, 5] != 0.0):
        gpars, dcov = model_deriv_data(model, pars, psf_pars)
        return _lib.noise_cov_b(gpars, dcov, pars[:, 5], jac, weights, noise)
    dimages = _fd_images(model, pars, psf_pars, jac, dims, per_object)
    badobj = np.isnan(dimages).any(axis=(1, 2))
    if badobj.any():
        dimages[badobj] = 0.0
    B = _lib.noise_b(dimages, weights, noise)
    B[badobj] = np.nan
    return B


def check_cov(pcov):
    """covariance sanity flags, leastsqbound.py:165-180"""
    flags = 0
    try:
        e = np.linalg.eigvals(pcov)
        if np.any(e < 0):
            flags |= LM_NEG_COV_EIG
        if np.any(np.diag(pcov) < 0):
            flags |= LM_NEG_COV_DIAG
    except np.linalg.LinAlgError:
        flags |= EIG_NOTFINITE
    return flags


def sandwich(pars_cov0, B):
    """(cov, flags): pars_cov0 B pars_cov0 and its sanity flags; a
    non-finite matrix is flagged as the reference does
    (noise_cov.py:63-74)"""
    npars = B.shape[0]
    cov = pars_cov0 @ B @ pars_cov0
    if not np.all(np.isfinite(cov)):
        return cov, check_cov(np.diag(np.full(npars, -1.0)))
    return cov, check_cov(cov)


def apply_noise_cov(result, B):
    """apply_noise_cov (noise_cov.py:38-87) on a result dict, in place"""
    if result["flags"] != 0:
        return
    pcov0 = result.get("pars_cov0")
    if pcov0 is None or not np.all(np.isfinite(pcov0)):
        return
    npars = result["pars"].size
    cov, cflags = sandwich(pcov0, B)
    if cflags != 0:
        result["flags"] |= cflags
        result["errmsg"] = "bad noise covariance matrix"
        result["pars_cov"] = np.zeros((npars, npars)) + CDEF
        result["pars_err"] = np.zeros(npars) + CDEF
    else:
        result["pars_cov"] = cov
        result["pars_err"] = np.sqrt(np.diag(cov))


def apply_noise_cov_batch(recs, B, npars):
    """the same on the structured records of fit_batch (matrices stored in
    8x8 slots, row major), in place"""
    for i in range(recs.size):
        if recs["flags"][i] != 0:
            continue
        n = npars
        pcov0 = recs["pars_cov0"][i][:n * n].reshape(n, n)
        if not np.all(np.isfinite(pcov0)):
            continue
        cov, cflags = sandwich(pcov0, B[i])
        if cflags != 0:
            recs["flags"][i] |= cflags
            recs["pars_cov"][i][:n * n] = CDEF
            recs["pars_err"][i][:n] = CDEF
            # no fit statistics on a flagged fit (results.py:45-72)
            for k in ("lnprob", "s2n_numer", "s2n_denom", "chi2per", "s2n",
                      "T", "T_err", "flux", "flux_err"):
                recs[k][i] = np.nan
            recs["npix"][i] = 0
            recs["dof"][i] = 0
        else:
            recs["pars_cov"][i][:n * n] = cov.ravel()
            recs["pars_err"][i][:n] = np.sqrt(np.diag(cov))
            recs["T_err"][i] = np.sqrt(cov[4, 4])
            recs["flux_err"][i] = np.sqrt(cov[n - 1, n - 1])
