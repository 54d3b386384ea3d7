"""
ngmix.em on the B200: EMFitter, EMFitterFixCen, EMFitterFixCov,
EMFitterFluxOnly, run_em, prep_image, prep_obs and EMResult with the
signatures of ngmix/em/em.py.  The sky shift (prep_image), the psf
normalisation, the E and M steps and the convergence test all run inside one
kernel launch (ngx_em); `go` is the n = 1 case of `em_batch`.
"""
import numpy as np

from . import _lib
from .flags import EM_MAXITER, EM_RANGE_ERROR
from .gmix import GMix
from .observation import Observation

__all__ = ['run_em', 'fit_em', 'prep_image', 'prep_obs', 'EMResult',
           'EMFitter', 'EMFitterFixCen', 'EMFitterFixCov',
           'EMFitterFluxOnly', 'em_batch']

DEFAULT_TOL = 1.0e-5


def em_batch(images, weights, jacobians, guesses, psf_pars=None,
             variant="full", tol=DEFAULT_TOL, miniter=None, maxiter=500,
             vary_sky=False, sky=None, ignore_zero_weight=True):
    """
    EM fits of n stamps in one launch.  guesses [n, ngauss, 6] are the full
    parameters of the pre-psf starting mixtures; psf_pars [n, npsf, 6] or
    None.  Returns (res, pars): res has fields flags, numiter, fdiff, sky and
    pars [n, ngauss, 6] the fitted mixtures.
    """
    return _lib.em(images, weights, jacobians, guesses, psf_pars=psf_pars,
                   variant=variant, tol=tol, miniter=miniter, maxiter=maxiter,
                   vary_sky=vary_sky, sky=sky,
                   ignore_zero_weight=ignore_zero_weight)


def prep_image(im0):
    """
    shift the image so its minimum is 0.001 (max-min) (em.py:95-120); the
    fitters apply the same shift on the device
    """
    im = np.array(im0, dtype="f8")
    im_min, im_max = im.min(), im.max()
    sky = 0.001 * (im_max - im_min) - im_min
    im += sky
    return im, sky


def prep_obs(obs):
    imsky, sky = prep_image(obs.image)
    newobs = obs.copy()
    newobs.image = imsky
    return newobs, sky


class EMResult(dict):
    """result of an EM fit: flags, numiter, fdiff, sky, message"""

    def __init__(self, obs, result, gm=None, gm_conv=None):
        self._obs = obs
        self.update(result)
        if gm is not None and gm_conv is not None:
            self._gm = gm
            self._gm_conv = gm_conv

    def has_gmix(self):
        return hasattr(self, '_gm')

    def get_gmix(self):
        if not self.has_gmix():
            raise RuntimeError('no gmix set')
        return self._gm.copy()

    def get_convolved_gmix(self):
        if not self.has_gmix():
            raise RuntimeError('no gmix set')
        return self._gm_conv.copy()

    def make_image(self):
        gm = self.get_convolved_gmix()
        return gm.make_image(self._obs.image.shape,
                             jacobian=self._obs.jacobian)


class EMFitter(object):
    """
    Fit an image with a gaussian mixture using the EM algorithm.

    Parameters
    ----------
    miniter, maxiter: int
        default 40, 500
    tol: float
        fractional change in the log likelihood that implies convergence
    vary_sky: bool
        fit for the sky level
    """
    _variant = "full"

    def __init__(self, miniter=40, maxiter=500, tol=DEFAULT_TOL,
                 vary_sky=False):
        self.miniter = miniter
        self.maxiter = maxiter
        self.tol = tol
        self.vary_sky = vary_sky

    def go(self, obs, guess, sky=None):
        if not isinstance(obs, Observation):
            raise ValueError('input obs must be an instance of Observation')
        psf_pars = None
        gmix_psf = None
        if obs.has_psf() and obs.psf.has_gmix():
            gmix_psf = obs.psf.gmix
            gmix_psf.set_flux(1.0)
            psf_pars = gmix_psf.get_full_pars().reshape(1, -1, 6)
        gpars = guess.get_full_pars().reshape(1, -1, 6)
        res, pars = em_batch(
            obs.image[None], obs.weight[None], obs._jacobian._data, gpars,
            psf_pars=psf_pars, variant=self._variant, tol=self.tol,
            miniter=self.miniter, maxiter=self.maxiter,
            vary_sky=self.vary_sky, sky=sky,
            ignore_zero_weight=obs.ignore_zero_weight)
        r = res[0]
        if r["flags"] & EM_RANGE_ERROR:
            result = {'flags': EM_RANGE_ERROR,
                      'message': 'range error in em_run'}
            return EMResult(obs=obs, result=result)
        gm = GMix(pars=pars[0].ravel())
        if gmix_psf is None:
            gm_conv = gm.copy()
        else:
            gm_conv = gm.convolve(gmix_psf)
        result = {
            'flags': int(r["flags"]),
            'numiter': int(r["numiter"]),
            'fdiff': float(r["fdiff"]),
            'sky': float(r["sky"]),
            'message': 'maxit' if r["flags"] & EM_MAXITER else 'OK',
        }
        return EMResult(obs=obs, result=result, gm=gm, gm_conv=gm_conv)


class EMFitterFixCen(EMFitter):
    """EM with the gaussian centres held fixed"""
    _variant = "fixcen"


class EMFitterFixCov(EMFitter):
    """EM with the gaussian covariances held fixed"""
    _variant = "fixcov"


class EMFitterFluxOnly(EMFitter):
    """EM varying only the fluxes; default miniter 20"""
    _variant = "fluxonly"

    def __init__(self, miniter=20, maxiter=500, tol=DEFAULT_TOL,
                 vary_sky=False):
        super().__init__(miniter=miniter, maxiter=maxiter, tol=tol,
                         vary_sky=vary_sky)


def run_em(obs, guess, sky=None, fixcen=False, fixcov=False, fluxonly=False,
           **kws):
    """fit the observation with EM (em.py:23-67)"""
    if fixcen:
        fitter = EMFitterFixCen(**kws)
    elif fixcov:
        fitter = EMFitterFixCov(**kws)
    elif fluxonly:
        fitter = EMFitterFluxOnly(**kws)
    else:
        fitter = EMFitter(**kws)
    return fitter.go(obs=obs, guess=guess, sky=sky)


fit_em = run_em
