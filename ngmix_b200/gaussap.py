"""
Gaussian-aperture fluxes of fitted models (ngmix/gaussap.py): for every object
and band the flux of the model mixture under a round gaussian weight of the
given fwhm.  Host arithmetic on the mixture parameters (no pixels involved).
"""
import numpy as np

from .flags import NO_ATTEMPT, GMIX_RANGE_ERROR
from .gexceptions import GMixRangeError
from .gmix import GMixCM, GMixModel

DEFAULT_FLUX = np.nan


def get_gaussap_flux(pars, model, weight_fwhm, fracdev=None, TdByTe=None,
                     mask=None, verbose=True):
    """
    pars [nobj, npars] (the fluxes of all bands at the end), model name
    ('cm' takes fracdev / TdByTe arrays); objects with mask False are not
    attempted.  Returns (flux [nobj, nband], flags [nobj, nband]); T is
    floored at 1e-4 as in the reference (gaussap.py:17-170).
    """
    pars = np.array(pars, dtype="f8", ndmin=2)
    nobj = pars.shape[0]
    mask = np.ones(nobj, dtype=bool) if mask is None \
        else np.array(mask, dtype=bool, ndmin=1)
    assert mask.shape[0] == nobj, "mask and pars must be same length"
    if model == "cm":
        fracdev = np.array(fracdev, dtype="f8", ndmin=1)
        TdByTe = np.array(TdByTe, dtype="f8", ndmin=1)
        assert fracdev.size == nobj and TdByTe.size == nobj
    nshape = {"bdf": 6, "bd": 7}.get(model, 5)
    nband = pars.shape[1] - nshape
    flux = np.full((nobj, nband), DEFAULT_FLUX)
    flags = np.zeros((nobj, nband), dtype="i4")
    for i in range(nobj):
        if not mask[i]:
            flags[i, :] = NO_ATTEMPT
            continue
        for band in range(nband):
            tpars = np.append(pars[i, :nshape], pars[i, nshape + band])
            tpars[4] = max(tpars[4], 0.0001)
            try:
                if model == "cm":
                    gm = GMixCM(fracdev[i], TdByTe[i], tpars)
                else:
                    gm = GMixModel(tpars, model)
                flux[i, band] = gm.get_gaussap_flux(fwhm=weight_fwhm)
            except GMixRangeError:
                flags[i, band] = GMIX_RANGE_ERROR
    return flux, flags
