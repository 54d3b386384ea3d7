"""
Gaussian-weighted moments (ngmix/gaussmom.py:7-95): a round gaussian weight of
given FWHM centred on the jacobian's origin, peak normalised to 1.  The pixel
pass runs on the device (ngx_weighted_sums); GaussMom.go handles one
Observation as the reference does, gaussmom_batch handles stacked stamps in
one launch.
"""
import numpy as np

from . import _lib, moments
from .gmix import GMixModel, get_weighted_moments_stats, moments_record


class GaussMom(object):
    """
    measure gaussian weighted moments

    parameters
    ----------
    fwhm: float
        The FWHM of the Gaussian weight function.
    with_higher_order: bool, optional
        If True the 17 moments of ngmix.moments.MOMENTS_NAME_MAP are returned
        in sums / sums_cov instead of the first 6.
    """
    kind = "wmom"

    def __init__(self, fwhm, with_higher_order=False):
        self.fwhm = fwhm
        self.with_higher_order = with_higher_order
        self._set_mompars()

    def go(self, obs):
        """moments of one Observation -> result dict"""
        res = self.weight.get_weighted_moments(
            obs=obs, with_higher_order=self.with_higher_order)
        if res["flags"] != 0:
            return res
        _to_flux_units(res, 1.0 / obs.jacobian.area)
        return res

    def _set_mompars(self):
        T = moments.fwhm_to_T(self.fwhm)
        weight = GMixModel([0.0, 0.0, 0.0, 0.0, T, 1.0], "gauss")
        # unit peak: flux = 1 / norm
        weight.set_norms()
        weight.set_flux(1.0 / weight.get_data()["norm"][0])
        self.weight = weight


def _to_flux_units(res, fac):
    """the sums are in surface-brightness units; take the pixel area out
    (gaussmom.py:62-72)"""
    res["flux"] *= fac
    res["flux_err"] *= fac
    res["pars"][5] *= fac
    res["sums"] *= fac
    res["sums_cov"] *= fac ** 2
    res["sums_norm"] *= fac
    res["wsum"] *= fac
    res["sums_err"] *= fac


def gaussmom_batch(images, weights, jacobians, fwhm, with_higher_order=False,
                   ignore_zero_weight=True):
    """
    GaussMom.go for n stacked stamps in one launch; returns a list of result
    dicts (the statistics of make_mom_result are host arithmetic on 6 or 17
    numbers per stamp).
    """
    images = np.ascontiguousarray(images, dtype="f8")
    n = images.shape[0]
    jac = _lib.as_jac(jacobians, n)
    gm = GaussMom(fwhm, with_higher_order).weight
    wt = np.repeat(gm.get_data()[None, :], n, axis=0)
    nmom = 17 if with_higher_order else 6
    maxrad = np.full(n, 100.0 * np.sqrt(gm.get_T() / 2.0))
    recs = _lib.weighted_sums(wt, images, weights, jac, maxrad, nmom,
                              ignore_zero_weight=ignore_zero_weight)
    out = []
    for i in range(n):
        res = get_weighted_moments_stats(moments_record(recs[i], nmom))
        if res["flags"] == 0:
            _to_flux_units(res, 1.0 / jac[i, 7] ** 2)
        out.append(res)
    return out
