"""
ngmix.admom on the B200: AdmomFitter, run_admom, find_cen_admom, AdmomResult
and get_result with the signatures of ngmix/admom/admom.py.  The iteration
(censums, momsums, convergence test, deweighting) runs inside one kernel
launch (ngx_admom); `go` is the n = 1 case of `admom_batch`.  get_result, the
scalar post-processing of the sums into flux/T/e/rho4 and their errors, is
vectorised over the batch on the host.
"""
import numpy as np

from . import _lib, flags as _flags
from .gexceptions import GMixRangeError  # noqa: F401 (re-export)
from .gmix import GMix, GMixModel
from .observation import Observation
from .shape import e1e2_to_g1g2, fwhm_to_T

__all__ = ['run_admom', 'find_cen_admom', 'AdmomFitter', 'AdmomResult',
           'admom_batch', 'get_result']

DEFAULT_MAXITER = 200
DEFAULT_SHIFTMAX = 5.0
DEFAULT_ETOL = 1.0e-5
DEFAULT_TTOL = 1.0e-3
GMIX_LOW_DETVAL = 1.0e-200


def admom_batch(images, weights, jacobians, guesses,
                maxiter=DEFAULT_MAXITER, shiftmax=DEFAULT_SHIFTMAX,
                etol=DEFAULT_ETOL, Ttol=DEFAULT_TTOL, cenonly=False):
    """
    adaptive moments of n stamps in one launch.  guesses [n, 6] are the full
    parameters (p,row,col,irr,irc,icc) of the starting weight gaussians.
    Returns the raw result records (flags, numiter, npix, wsum, sums,
    sums_cov, pars, rho4, wt_norm).
    """
    return _lib.admom(images, weights, jacobians, guesses, maxiter=maxiter,
                      shiftmax=shiftmax, etol=etol, Ttol=Ttol,
                      cenonly=cenonly)


def get_ratio_error(a, b, var_a, var_b, cov_ab):
    """error on a/b (ngmix/util.py:57-80)"""
    if b == 0:
        raise ValueError("zero in denominator")
    rsq = (a / b) ** 2
    with np.errstate(divide="ignore", invalid="ignore"):
        var = rsq * (var_a / a ** 2 + var_b / b ** 2 - 2 * cov_ab / (a * b))
    return np.sqrt(np.clip(var, 0.0, np.inf))


def get_result(ares, jac_area, wgt_norm):
    """
    copy one raw admom record to a dict and derive flux, T, e, rho4 with
    errors and sub-flags (admom.py:406-568)
    """
    res = {n: (ares[n].copy() if isinstance(ares[n], np.ndarray) else ares[n])
           for n in ("flags", "numiter", "npix", "wsum", "sums", "sums_cov",
                     "pars", "rho4", "F")}
    res["flags"] = int(res["flags"])
    res["numiter"] = int(res["numiter"])
    res["npix"] = int(res["npix"])
    res["sums_cov"] = res["sums_cov"].reshape(7, 7)
    res["sums_norm"] = ares["wsum"]
    for k in ("flux_flags", "T_flags", "rho4_flags"):
        res[k] = 0
    nan = np.nan
    for k in ("flux", "flux_mean", "flux_err", "T", "T_err", "rho4",
              "rho4_err", "s2n", "e1", "e2", "e1err", "e2err"):
        res[k] = nan
    res["e"] = np.array([nan, nan])
    res["e_err"] = np.array([nan, nan])
    res["e_cov"] = np.diag([nan, nan])
    sums, cov = res["sums"], res["sums_cov"]

    if res["flags"] == 0:
        res["T"] = res["pars"][4]
        res["rho4"] = ares["rho4"]
        res["flux_mean"] = sums[5] / res["wsum"]
        res["pars"][5] = res["flux_mean"]
        if res["T"] > GMIX_LOW_DETVAL:
            fnorm = jac_area * wgt_norm * res["wsum"]
            res["flux"] = sums[5] / fnorm
            if cov[5, 5] > 0:
                res["flux_err"] = np.sqrt(cov[5, 5]) / fnorm
                res["s2n"] = res["flux"] / res["flux_err"]
            else:
                res["flux_flags"] |= _flags.NONPOS_VAR
        else:
            res["flux_flags"] |= _flags.NONPOS_SIZE
        for name, idx in (("T", 4), ("rho4", 6)):
            if cov[idx, idx] > 0 and cov[5, 5] > 0:
                if sums[5] > 0:
                    if name == "rho4":
                        res["rho4"] = sums[6] / sums[5]
                    res[name + "_err"] = 4 * get_ratio_error(
                        sums[idx], sums[5], cov[idx, idx], cov[5, 5],
                        cov[idx, 5])
                else:
                    res[name + "_flags"] |= _flags.NONPOS_FLUX
            else:
                res[name + "_flags"] |= _flags.NONPOS_VAR
    else:
        res["flux_flags"] |= res["flags"]
        res["T_flags"] |= res["flags"]
        res["rho4_flags"] |= res["flags"]

    if not np.all(np.diagonal(cov[2:, 2:]) > 0):
        res["flags"] |= _flags.NONPOS_VAR

    if res["flags"] == 0:
        if sums[5] > 0:
            if res["T"] > 0.0:
                res["e"][:] = res["pars"][2:4] / res["T"]
                res["e1"], res["e2"] = res["e"]
                res["e1err"] = 2 * get_ratio_error(
                    sums[2], sums[4], cov[2, 2], cov[4, 4], cov[2, 4])
                res["e2err"] = 2 * get_ratio_error(
                    sums[3], sums[4], cov[3, 3], cov[4, 4], cov[3, 4])
                if not (np.isfinite(res["e1err"])
                        and np.isfinite(res["e2err"])):
                    res["e1err"] = res["e2err"] = nan
                    res["flags"] |= _flags.NONPOS_SHAPE_VAR
                else:
                    res["e_cov"] = np.diag([res["e1err"] ** 2,
                                            res["e2err"] ** 2])
                    res["e_err"] = np.array([res["e1err"], res["e2err"]])
            else:
                res["flags"] |= _flags.NONPOS_SIZE
        else:
            res["flags"] |= _flags.NONPOS_FLUX

    for k in ("flags", "flux_flags", "T_flags", "rho4_flags"):
        res[k.replace("flags", "flagstr")] = _flags.get_flags_str(res[k])
    return res


class AdmomResult(dict):
    """dict of the adaptive-moment results with mixture helpers"""

    def __init__(self, obs, result):
        self._obs = obs
        self.update(result)

    def get_gmix(self):
        if self['flags'] != 0:
            raise RuntimeError('cannot create gmix, fit failed')
        pars = self['pars'].copy()
        pars[5] = 1.0
        g1, g2 = e1e2_to_g1g2(pars[2] / pars[4], pars[3] / pars[4])
        pars[2], pars[3] = g1, g2
        return GMixModel(pars, "gauss")

    def make_image(self):
        if self['flags'] != 0:
            raise RuntimeError('cannot create image, fit failed')
        gm = self.get_gmix()
        gm.set_flux(self._obs.image.sum())
        return gm.make_image(self._obs.image.shape,
                             jacobian=self._obs.jacobian)


class AdmomFitter(object):
    """
    Measure adaptive moments for the input observation.

    Parameters
    ----------
    maxiter: int, default 200
    shiftmax: float, default 5.0
        largest allowed centroid shift relative to the guess
    etol, Ttol: float, default 1e-5, 1e-3
        convergence tolerances on e1/e2 (absolute) and T (relative)
    cenonly: bool
        only vary the centre
    rng: np.random.RandomState
        for generating a guess from a T value
    """
    kind = "am"

    def __init__(self, maxiter=DEFAULT_MAXITER, shiftmax=DEFAULT_SHIFTMAX,
                 etol=DEFAULT_ETOL, Ttol=DEFAULT_TTOL, cenonly=False,
                 rng=None):
        self.conf = dict(maxiter=maxiter, shiftmax=shiftmax, etol=etol,
                         Ttol=Ttol, cenonly=cenonly)
        self.rng = rng

    def go(self, obs, guess):
        if not isinstance(obs, Observation):
            raise ValueError("input obs must be an Observation")
        guess_gmix = self._get_guess(obs=obs, guess=guess)
        gd = guess_gmix._data
        # admom works on the caller's weight mixture in place: mirror that
        gpars = guess_gmix.get_full_pars()[:6]
        rec = admom_batch(obs.image[None], obs.weight[None],
                          obs._jacobian._data, gpars[None], **self.conf)[0]
        result = get_result(rec, obs.jacobian.area, rec["wt_norm"])
        del gd
        return AdmomResult(obs=obs, result=result)

    def _get_guess(self, obs, guess):
        if isinstance(guess, GMix):
            return guess
        return self._generate_guess(obs=obs, Tguess=guess)

    def _get_rng(self):
        if self.rng is None:
            self.rng = np.random.RandomState()
        return self.rng

    def _generate_guess(self, obs, Tguess):
        rng = self._get_rng()
        scale = obs.jacobian.get_scale()
        pars = np.zeros(6)
        pars[0:2] = rng.uniform(low=-0.5 * scale, high=0.5 * scale, size=2)
        pars[2:4] = rng.uniform(low=-0.3, high=0.3, size=2)
        pars[4] = Tguess * (1.0 + rng.uniform(low=-0.1, high=0.1))
        pars[5] = 1.0
        return GMixModel(pars, "gauss")


def run_admom(obs, guess, maxiter=DEFAULT_MAXITER, shiftmax=DEFAULT_SHIFTMAX,
              etol=DEFAULT_ETOL, Ttol=DEFAULT_TTOL, cenonly=False, rng=None):
    """run adaptive moments on the observation (admom.py:20-71)"""
    am = AdmomFitter(maxiter=maxiter, shiftmax=shiftmax, etol=etol, Ttol=Ttol,
                     cenonly=cenonly, rng=rng)
    return am.go(obs=obs, guess=guess)


def find_cen_admom(obs, fwhm=None, gmix=None, maxiter=DEFAULT_MAXITER,
                   shiftmax=DEFAULT_SHIFTMAX, etol=DEFAULT_ETOL,
                   Ttol=DEFAULT_TTOL, ntry=1, rng=None):
    """centre from adaptive moments with a fixed weight (admom.py:74-164)"""
    if ntry > 1 and rng is None:
        raise ValueError('send a random number generator rng= when trying '
                         'more than once')
    if gmix is not None:
        wt = gmix.copy()
    elif fwhm is not None:
        wt = GMixModel([0.0, 0.0, 0.0, 0.0, fwhm_to_T(fwhm), 1.0], 'gauss')
    else:
        raise ValueError('send gmix= or fwhm=')
    scale = obs.jacobian.scale
    am = AdmomFitter(maxiter=maxiter, shiftmax=shiftmax, etol=etol, Ttol=Ttol,
                     cenonly=True)
    for itry in range(ntry):
        res = am.go(obs=obs, guess=wt)
        if res['flags'] == 0:
            break
        if ntry > 1:
            drow, dcol = rng.uniform(low=-scale / 2, high=scale / 2, size=2)
            wt.set_cen(row=drow, col=dcol)
    if res['flags'] == 0:
        res['cen'] = res.get_gmix().get_cen()
    else:
        res['cen'] = np.zeros(2) + np.nan
    return res
