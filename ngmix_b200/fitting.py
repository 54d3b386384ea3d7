"""
ngmix.fitting.Fitter on the B200: same constructor and `go(obs, guess)` as
ngmix/fitting/fitters.py:20-117, same result keys as run_leastsq
(fitting/leastsqbound.py:118-126) + FitModel.set_fit_result
(fitting/results.py:45-72).  The whole fit - residuals, analytic or
forward-difference Jacobian, MINPACK iteration, covariance and fit statistics
- runs in one kernel launch (ngx_lm_fit); `go` is the n = 1 case of
`fit_batch`, which takes stacked arrays and is the fast path.
"""
import numpy as np

from . import _lib
from .defaults import DEFAULT_LM_PARS
from .gexceptions import GMixFatalError, GMixRangeError
from .gmix import (get_model_name, get_model_npars, get_model_num,
                   make_gmix_model)
from .observation import Observation, get_mb_obs

__all__ = ['Fitter', 'CoellipFitter', 'FitResult', 'FitModel', 'fit_batch',
           'fit_batch_coellip',
           'SIMPLE_ANALYTIC_MODELS', 'get_band_pars', 'get_lm_n_prior_pars']

# models with an analytic Jacobian (fitting/results.py:921)
SIMPLE_ANALYTIC_MODELS = ('gauss', 'exp', 'dev')
_LM_MODELS = ('gauss', 'exp', 'dev', 'turb', 'bdf', 'bd')


def _prior_record(prior):
    if prior is None:
        return None
    if isinstance(prior, _lib.Prior):
        return prior
    if hasattr(prior, "to_ngx"):
        return prior.to_ngx()
    raise TypeError("unsupported prior %s" % type(prior))


def fit_batch(model, images, weights, jacobians, psf_pars, guesses,
              prior=None, fit_pars=None, analytic_jacobian=True,
              ignore_zero_weight=True, noise_images=None, mask=None):
    """
    Levenberg-Marquardt fits of n stamps in one launch.

    Parameters
    ----------
    model: str
        gauss, exp, dev, turb, bdf or bd
    images, weights: [n, H, W] arrays
    jacobians: [n, 8] array (or Jacobian records)
        row0, col0, dvdrow, dvdcol, dudrow, dudcol, det, scale per stamp
    psf_pars: [n, npsf, 6] array or None
        full gaussian parameters of each stamp's psf mixture
    guesses: [n, npars] array
    mask: [n, H, W] bool array, optional
        True = use the pixel.  With float32 `images` and ONE weight per stamp
        (`weights` a scalar or an [n] array) this is the compact survey format
        (ngx_lm_fit_host_compact): 4.1 instead of 16 bytes per pixel cross
        PCIe; the results equal those of the expanded arrays
        images.astype(f8), where(mask, weight, 0) bit for bit.
    prior: PriorSimpleSep / PriorBDFSep, optional
    fit_pars: dict with maxfev, ftol, xtol, optional
    noise_images: [n, H, W] array, optional
        independent noise realisations; if sent, pars_cov / pars_err become
        the noise-power sandwich covariance (Fitter(use_noise_image=True),
        fitting/noise_cov.py)

    Returns
    -------
    structured array [n] with fields flags, nfev, ier, npix, dof, lnprob,
    s2n_numer, s2n_denom, chi2per, s2n, T, T_err, flux, flux_err, pars,
    pars_err, pars_cov0, pars_cov (matrices flattened row major, 8x8 slots)
    """
    model = get_model_name(get_model_num(model))
    if model not in _LM_MODELS:
        raise ValueError("model '%s' cannot be fit with ngx_lm_fit" % model)
    fp = dict(DEFAULT_LM_PARS)
    if fit_pars is not None:
        fp.update(fit_pars)
    w_arr = np.asarray(weights)
    if w_arr.ndim <= 1:
        # one weight per stamp (+ optional mask)
        images = np.asarray(images)
        if images.ndim == 2:
            images = images[None]
        n = images.shape[0]
        w_arr = np.broadcast_to(w_arr.astype("f8"), (n,))
        if images.dtype == np.float32 and noise_images is None \
                and ignore_zero_weight:
            return _lib.lm_fit_compact(
                model, images, w_arr, jacobians, psf_pars, guesses, mask=mask,
                prior=_prior_record(prior), maxfev=fp["maxfev"],
                ftol=fp["ftol"], xtol=fp["xtol"],
                analytic_jacobian=analytic_jacobian,
                bounds=getattr(prior, "bounds", None))
        weights = np.empty(images.shape)
        weights[...] = w_arr[:, None, None]
        if mask is not None:
            weights[~np.asarray(mask, dtype=bool)] = 0.0
    elif mask is not None:
        weights = np.where(np.asarray(mask, dtype=bool), w_arr, 0.0)
    recs = _lib.lm_fit(model, images, weights, jacobians, psf_pars, guesses,
                       prior=_prior_record(prior), maxfev=fp["maxfev"],
                       ftol=fp["ftol"], xtol=fp["xtol"],
                       analytic_jacobian=analytic_jacobian,
                       ignore_zero_weight=ignore_zero_weight,
                       bounds=getattr(prior, "bounds", None))
    if noise_images is not None:
        from . import noise_cov
        npars = get_model_npars(model)
        ok = np.flatnonzero(recs["flags"] == 0)
        if ok.size > 0:
            psf_ok = None if psf_pars is None else np.asarray(psf_pars)[ok]
            B = np.zeros((recs.size, npars, npars))
            B[ok] = noise_cov.calc_noise_b(
                model, recs["pars"][ok, :npars], psf_ok,
                _lib.as_jac(jacobians)[ok], np.asarray(weights)[ok],
                np.asarray(noise_images)[ok], per_object=True)
            noise_cov.apply_noise_cov_batch(recs, B, npars)
    return recs


def fit_batch_multi(model, stamp_start, stamp_band, nband, images, weights,
                    jacobians, psf_pars, guesses, prior=None, fit_pars=None,
                    ignore_zero_weight=True, analytic_jacobian=True):
    """
    Levenberg-Marquardt fits of objects observed in several epochs and/or
    bands (Fitter.go on ObsList / MultiBandObsList, fitting/results.py:
    212-570) in one launch.  Object i owns stamps
    stamp_start[i]:stamp_start[i+1] of the [ns, H, W] stamp arrays, ordered
    band by band; stamp_band [ns] gives each stamp's band; guesses
    [nobj, nshared + nband] = (cen1, cen2, g1, g2, T, [fracdev | logTratio,
    fracdev,] F_0 .. F_{nband-1}).  gauss / exp / dev use the analytic
    Jacobian unless analytic_jacobian=False; bdf, bd and turb always run
    MINPACK's forward differences, as Fitter.go does.  nband <= 12 and at most
    20 parameters (records with 14 parameter slots beyond 8 parameters, 20
    beyond 14 parameters or 7 bands).
    """
    model = get_model_name(get_model_num(model))
    if model not in _LM_MODELS:
        raise ValueError("model '%s' cannot be fit with ngx_lm_fit" % model)
    fp = dict(DEFAULT_LM_PARS)
    if fit_pars is not None:
        fp.update(fit_pars)
    pr = _prior_record(prior)
    if pr is not None and getattr(prior, "nband", nband) != nband:
        raise ValueError("prior has %d flux priors, data has %d bands"
                         % (prior.nband, nband))
    return _lib.lm_fit_multi(model, stamp_start, stamp_band, nband, images,
                             weights, jacobians, psf_pars, guesses, prior=pr,
                             maxfev=fp["maxfev"], ftol=fp["ftol"],
                             xtol=fp["xtol"],
                             ignore_zero_weight=ignore_zero_weight,
                             analytic_jacobian=analytic_jacobian,
                             bounds=getattr(prior, "bounds", None))


_ERRMSG = {
    1: "Both actual and predicted relative reductions in the sum of squares "
       "are at most ftol",
    2: "The relative error between two consecutive iterates is at most xtol",
    3: "Both ftol and xtol termination conditions are satisfied",
    4: "The cosine of the angle between func(x) and any column of the "
       "Jacobian is at most gtol in absolute value",
    5: "Number of calls to function has reached maxfev",
    6: "ftol is too small, no further reduction in the sum of squares is "
       "possible",
    7: "xtol is too small, no further improvement in the approximate "
       "solution is possible",
    8: "gtol is too small, func(x) is orthogonal to the columns of the "
       "Jacobian to machine precision",
}


def record_to_result_n(rec, model, npars):
    """record_to_result for models whose parameter count is not tabulated
    (coellip): single band"""
    return record_to_result(rec, model, npars, nband=1)


def record_to_result(rec, model, npars, nband=None):
    """one record of fit_batch -> the reference's result dict"""
    n = npars
    res = {
        "model": model,
        "flags": int(rec["flags"]),
        "nfev": int(rec["nfev"]),
        "ier": int(rec["ier"]),
        "errmsg": _ERRMSG.get(int(rec["ier"]), ""),
        "pars": rec["pars"][:n].copy(),
        "pars_err": rec["pars_err"][:n].copy(),
        "pars_cov": rec["pars_cov"][:n * n].reshape(n, n).copy(),
    }
    cov0 = rec["pars_cov0"][:n * n].reshape(n, n).copy()
    res["pars_cov0"] = None if np.isnan(cov0).all() else cov0
    if res["flags"] & _lib_flags("LM_SINGULAR_MATRIX"):
        res["errmsg"] = "singular covariance"
    if res["flags"] == 0:
        for k in ("lnprob", "s2n_numer", "s2n_denom", "chi2per", "s2n",
                  "T", "T_err", "flux", "flux_err"):
            res[k] = float(rec[k])
        res["npix"] = int(rec["npix"])
        res["dof"] = int(rec["dof"])
        res["s2n_w"] = res["s2n"]
        res["g"] = res["pars"][2:4].copy()
        res["g_cov"] = res["pars_cov"][2:4, 2:4].copy()
        res["g_err"] = res["pars_err"][2:4].copy()
        if nband is None:
            nband = n - get_model_npars(model) + 1
        if nband > 1:
            # _set_flux, results.py:1079-1109
            start = n - nband
            res["flux"] = res["pars"][start:].copy()
            res["flux_cov"] = res["pars_cov"][start:, start:].copy()
            res["flux_err"] = np.sqrt(np.diag(res["flux_cov"]))
    return res


_STAT_KEYS = ("lnprob", "s2n_numer", "s2n_denom", "chi2per", "s2n", "s2n_w",
              "npix", "dof", "g", "g_cov", "g_err", "T", "T_err", "flux",
              "flux_err", "flux_cov")


def refresh_result_after_noise_cov(res, nband):
    """The reference applies the noise covariance BEFORE set_fit_result
    (fitting/fitters.py:105-110), so every key derived from pars_cov comes
    from the sandwich matrix, and a fit whose sandwich matrix is bad
    (flags != 0) carries no fit statistics at all (results.py:45-72)."""
    if res["flags"] != 0:
        for k in _STAT_KEYS:
            res.pop(k, None)
        return res
    n = res["pars"].size
    cov, err = res["pars_cov"], res["pars_err"]
    res["g_cov"] = cov[2:4, 2:4].copy()
    res["g_err"] = err[2:4].copy()
    res["T_err"] = float(np.sqrt(cov[4, 4]))
    if nband > 1:
        start = n - nband
        res["flux_cov"] = cov[start:, start:].copy()
        res["flux_err"] = np.sqrt(np.diag(res["flux_cov"]))
    else:
        res["flux_err"] = float(np.sqrt(cov[n - 1, n - 1]))
    return res


def _lib_flags(name):
    from . import flags
    return getattr(flags, name)


class FitResult(dict):
    """
    The result of Fitter.go: a dict with the reference's keys plus the
    mixture helpers of FitModel (results.py:74-123).
    """

    def __init__(self, obs, model, result):
        super().__init__(result)
        self.obs = get_mb_obs(obs)
        self.model_name = model
        self.nband = len(self.obs)
        self.npars = get_model_npars(model) + self.nband - 1

    def get_band_pars(self, pars, band):
        """results.py:1025-1047 (simple models: 5 shared + the band's flux)"""
        n0 = get_model_npars(self.model_name)
        out = np.array(pars[:n0], dtype="f8")
        out[n0 - 1] = pars[n0 - 1 + band]
        return out

    def get_gmix(self, band=0):
        return make_gmix_model(self.get_band_pars(self["pars"], band),
                               self.model_name)

    def get_convolved_gmix(self, band=0, obsnum=0):
        gm = self.get_gmix(band)
        obs = self.obs[band][obsnum]
        if obs.has_psf_gmix():
            gm = gm.convolve(obs.psf.gmix)
        return gm

    def make_image(self, band=0, obsnum=0):
        gm = self.get_convolved_gmix(band=band, obsnum=obsnum)
        obs = self.obs[band][obsnum]
        return gm.make_image(obs.image.shape, jacobian=obs.jacobian)


def get_lm_n_prior_pars(model, nband):
    """residual slots reserved for the prior (fitting/results.py:1050-1076)"""
    if model == "bd":
        return 6 + nband
    if model in ("bdf", "exp", "dev", "gauss", "turb"):
        return 5 + nband
    raise ValueError("bad model: %s" % model)


def get_band_pars(model, pars, band):
    """the parameters of one band: everything up to the fluxes plus that
    band's flux (fitting/results.py:1013-1047)"""
    nshape = get_model_npars(model) - 1
    return np.append(np.asarray(pars[:nshape], dtype="f8"),
                     pars[nshape + band])


class FitModel(FitResult):
    """
    The fit model of the reference (fitting/results.py:17-625): the residual
    vector, its Jacobian and ln(prob) at arbitrary parameters, evaluated
    through the same device passes the fit uses, plus the result dict.

    calc_fdiff    prior rows, then (model - data) * ierr of every epoch, -inf
                  everywhere on a range error            (results.py:439-485)
    calc_jacobian prior rows by forward differences, pixel rows from
                  deriv_images, zero on a range error    (results.py:487-625)
    calc_lnprob   sum of get_loglike + ln(prior)         (results.py:142-210)
    """

    def __init__(self, obs, model, guess, prior=None):
        model = get_model_name(get_model_num(model))
        super().__init__(obs, model, {})
        self.prior = prior
        self.model = get_model_num(model)
        self.n_prior_pars = 0 if prior is None else \
            get_lm_n_prior_pars(model, self.nband)
        self.totpix = sum(o.pixels.size for ol in self.obs for o in ol)
        self.fdiff_size = self.totpix + self.n_prior_pars
        guess = np.array(guess, dtype="f8")
        assert guess.size == self.npars, \
            "guess has npars=%d, expected %d" % (guess.size, self.npars)
        # as FitModel._setup_fit: a guess outside the model's domain raises
        self._convolved(guess)

    def _convolved(self, pars):
        """[(obs, convolved GMix, pre-psf GMix, band)] at pars"""
        out = []
        for band, obslist in enumerate(self.obs):
            gm0 = make_gmix_model(get_band_pars(self.model_name, pars, band),
                                  self.model_name)
            for o in obslist:
                gm = gm0.convolve(o.psf.gmix) if o.has_psf_gmix() else gm0
                gm.set_norms()
                out.append((o, gm, gm0, band))
        return out

    def set_fit_result(self, result):
        self.update(result)

    def calc_fdiff(self, pars):
        from .defaults import LOWVAL
        pars = np.asarray(pars, dtype="f8")
        fdiff = np.zeros(self.fdiff_size)
        try:
            mix = self._convolved(pars)
            start = 0
            if self.prior is not None:
                start = self.prior.fill_fdiff(pars, fdiff)
            for o, gm, _, _ in mix:
                pix = o.pixels
                _, fd, st = _lib.loglike_pixels(gm._data[None], pix[None],
                                                want_fdiff=True)
                if st[0]:
                    raise GMixRangeError("det or T too low")
                fdiff[start:start + pix.size] = fd[0]
                start += pix.size
        except GMixRangeError:
            fdiff[:] = LOWVAL
        return fdiff

    def calc_lnprob(self, pars, more=False):
        from .defaults import LOWVAL, BIGVAL
        pars = np.asarray(pars, dtype="f8")
        tot = {"lnprob": 0.0, "s2n_numer": 0.0, "s2n_denom": 0.0, "npix": 0}
        try:
            if self.prior is not None:
                tot["lnprob"] += self.prior.get_lnprob_scalar(pars)
            for o, gm, _, _ in self._convolved(pars):
                r = gm.get_loglike(o, more=True)
                tot["lnprob"] += r["loglike"]
                for k in ("s2n_numer", "s2n_denom", "npix"):
                    tot[k] += r[k]
        except GMixRangeError:
            tot = {"lnprob": LOWVAL, "s2n_numer": 0.0, "s2n_denom": BIGVAL,
                   "npix": 0}
        return tot if more else tot["lnprob"]

    def calc_jacobian(self, pars):
        from .noise_cov import model_deriv_data
        if self.model_name not in SIMPLE_ANALYTIC_MODELS:
            raise ValueError("analytic jacobian is not available for model "
                             "%s" % self.model_name)
        pars = np.asarray(pars, dtype="f8")
        jac = np.zeros((self.fdiff_size, self.npars))
        try:
            start = self._fill_prior_jacobian(pars, jac)
            mix = self._convolved(pars)
            for o, _, _, band in mix:
                bp = get_band_pars(self.model_name, pars, band)
                if bp[4] == 0.0 or bp[5] == 0.0:
                    raise GMixRangeError("zero T or flux")
                psf_pars = None
                if o.has_psf_gmix():
                    psf_pars = o.psf.gmix.get_full_pars().reshape(1, -1, 6)
                gpars, dcov = model_deriv_data(self.model_name, bp[None],
                                               psf_pars)
                out = _lib.deriv_images(gpars, dcov, o._jacobian._data,
                                        o.image.shape)[0]
                # the reference evaluates on obs.pixels: the valid pixels in
                # row-major order
                keep = np.ones(o.image.size, dtype=bool)
                if o.ignore_zero_weight:
                    keep = o.weight.ravel() > 0.0
                out = out[:, keep]
                ierr = o.pixels["ierr"]
                sl = slice(start, start + ierr.size)
                for k in range(5):
                    jac[sl, k] = out[1 + k] * ierr
                jac[sl, 5 + band] = out[0] * (ierr / bp[5])
                start += ierr.size
        except (GMixRangeError, ValueError):
            jac[:] = 0.0
        return jac

    def _fill_prior_jacobian(self, pars, jac):
        """forward differences of prior.fill_fdiff, backward where the step
        leaves the prior's domain (results.py:572-625)"""
        if self.prior is None:
            return 0
        step_prior = 1.0e-8
        f0 = np.zeros(self.n_prior_pars)
        fs = np.zeros(self.n_prior_pars)
        n = self.prior.fill_fdiff(pars, f0)
        good0 = np.isfinite(f0[:n])
        for ipar in range(self.npars):
            step = step_prior * max(1.0, abs(pars[ipar]))
            p = pars.copy()
            p[ipar] = pars[ipar] + step
            try:
                self.prior.fill_fdiff(p, fs)
            except GMixRangeError:
                step = -step
                p[ipar] = pars[ipar] + step
                self.prior.fill_fdiff(p, fs)
            good = good0 & np.isfinite(fs[:n])
            d = np.zeros(n)
            d[good] = (fs[:n][good] - f0[:n][good]) / step
            jac[:n, ipar] = d
        return n


class PSFFluxFitter(object):
    """
    Template flux by zero-lag cross-correlation (fitting/fitters.py:144-181,
    results.py:676-805) through ngx_template_flux: with do_psf the psf
    mixtures of the observation(s) are the templates ("psf flux"), otherwise
    the observations' own mixtures; without a mixture the pixel template
    (obs.psf.template / obs.template) is used (ngx_template_flux_images).
    """

    def __init__(self, do_psf=True, normalize_psf=True):
        self.do_psf = do_psf
        self.normalize_psf = normalize_psf

    def go(self, obs):
        from .observation import Observation, ObsList
        if isinstance(obs, Observation):
            obslist = [obs]
        elif isinstance(obs, ObsList):
            obslist = list(obs)
        else:
            raise ValueError("obs should be Observation or ObsList")
        first = obslist[0].psf if self.do_psf else obslist[0]
        if not first.has_gmix():
            # pixel templates (results.py:834-864)
            if getattr(first, "template", None) is None:
                raise ValueError("neither gmix or template image are set")
            return self._go_templates(obslist)
        gms = []
        for o in obslist:
            src = o.psf if self.do_psf else o
            if not src.has_gmix():
                raise ValueError("neither gmix or template image are set")
            gms.append(src.gmix.get_full_pars().reshape(-1, 6))
        shape = obslist[0].image.shape
        if any(o.image.shape != shape for o in obslist) \
                or any(g.shape != gms[0].shape for g in gms):
            raise NotImplementedError(
                "all stamps must share their size and mixture size")
        # without do_psf the template is always normalised (results.py:853)
        normalize = self.normalize_psf if self.do_psf else True
        rec = _lib.template_flux(
            [0, len(obslist)], np.stack([o.image for o in obslist]),
            np.stack([o.weight for o in obslist]),
            np.stack([np.asarray(o._jacobian._data.view("f8")).reshape(8)
                      for o in obslist]),
            np.stack(gms), normalize=normalize,
            ignore_zero_weight=all(o.ignore_zero_weight for o in obslist))[0]
        if rec["flags"] & 128:
            raise GMixRangeError("det or T too low")
        return {"model": "template", "flags": int(rec["flags"]),
                "chi2per": float(rec["chi2per"]), "dof": float(rec["dof"]),
                "flux": float(rec["flux"]),
                "flux_err": float(rec["flux_err"])}


def _psfflux_go_templates(self, obslist):
    """PSFFluxFitModel with pixel templates: obs.psf.template (do_psf) or
    obs.template, through ngx_template_flux_images"""
    tps = [np.asarray((o.psf if self.do_psf else o).template, dtype="f8")
           for o in obslist]
    shape = obslist[0].image.shape
    if any(o.image.shape != shape for o in obslist) \
            or any(t.shape != shape for t in tps):
        raise NotImplementedError(
            "all stamps and templates must share their size")
    normalize = self.normalize_psf if self.do_psf else True
    rec = _lib.template_flux_images(
        [0, len(obslist)], np.stack([o.image for o in obslist]),
        np.stack([o.weight for o in obslist]), np.stack(tps),
        normalize=normalize,
        ignore_zero_weight=all(o.ignore_zero_weight for o in obslist))[0]
    return {"model": "template", "flags": int(rec["flags"]),
            "chi2per": float(rec["chi2per"]), "dof": float(rec["dof"]),
            "flux": float(rec["flux"]), "flux_err": float(rec["flux_err"])}


PSFFluxFitter._go_templates = _psfflux_go_templates


class Fitter(object):
    """
    Levenberg-Marquardt fit of a PSF-convolved model to an observation.

    Parameters
    ----------
    model: str
        The model to fit
    prior: PriorSimpleSep or PriorBDFSep, optional
    fit_pars: dict, optional
        maxfev, ftol, xtol (defaults 4000, 1e-5, 1e-5)
    use_noise_image: bool, optional
        replace the chi^2-scaled covariance by the noise-power sandwich
        covariance measured from the noise images attached to the
        observations (fitting/noise_cov.py); every observation must then
        carry a noise image.  Default False.
    analytic_jacobian: bool, optional
        analytic Jacobian for gauss/exp/dev, forward differences otherwise
        (MINPACK lmder / lmdif, as in the reference).  Default True.
    """

    def __init__(self, model, prior=None, fit_pars=None,
                 use_noise_image=False, analytic_jacobian=True):
        self.prior = prior
        self.model = get_model_num(model)
        self.model_name = get_model_name(self.model)
        self.use_noise_image = use_noise_image
        self.analytic_jacobian = analytic_jacobian
        self.fit_pars = dict(fit_pars) if fit_pars is not None \
            else dict(DEFAULT_LM_PARS)

    def go(self, obs, guess):
        """run the fit; returns a dict-like FitResult"""
        mb = get_mb_obs(obs)
        if self.use_noise_image:
            for obslist in mb:
                for tobs in obslist:
                    if not tobs.has_noise():
                        raise ValueError("obs.noise must be set when "
                                         "use_noise_image=True")
        if len(mb) != 1 or len(mb[0]) != 1:
            return self._go_multi(mb, guess)
        o = mb[0][0]
        npars = get_model_npars(self.model_name)
        guess = np.array(guess, dtype="f8")
        assert guess.size == npars, \
            "guess has npars=%d, expected %d" % (guess.size, npars)
        psf_pars = None
        if o.has_psf_gmix():
            psf_pars = o.psf.gmix.get_full_pars().reshape(1, -1, 6)
        # FitModel._setup_fit fills the mixtures at the guess and lets a
        # GMixRangeError propagate (results.py:359-375)
        gm0 = make_gmix_model(guess, self.model_name)
        del gm0
        rec = fit_batch(self.model_name, o.image[None], o.weight[None],
                        o._jacobian._data, psf_pars, guess[None],
                        prior=self.prior, fit_pars=self.fit_pars,
                        analytic_jacobian=self.analytic_jacobian,
                        ignore_zero_weight=o.ignore_zero_weight)[0]
        if rec["flags"] == 1 and rec["nfev"] == 0:
            raise GMixFatalError("no weights > 0")
        res = record_to_result(rec, self.model_name, npars)
        if self.use_noise_image:
            self._apply_noise_cov(mb, res)
        return FitResult(o, self.model_name, res)

    def _go_multi(self, mb, guess):
        """ObsList / MultiBandObsList: every epoch of every band is a stamp
        of one object (same stamp size and psf mixture size throughout)"""
        nband = len(mb)
        npars = get_model_npars(self.model_name) + nband - 1
        guess = np.array(guess, dtype="f8")
        assert guess.size == npars, \
            "guess has npars=%d, expected %d" % (guess.size, npars)
        if nband > 12 or npars > 20:
            raise NotImplementedError("at most 12 bands and 20 parameters")
        stamps = [(b, o) for b, ol in enumerate(mb) for o in ol]
        shape = stamps[0][1].image.shape
        has_psf = stamps[0][1].has_psf_gmix()
        for _, o in stamps:
            if o.image.shape != shape or o.has_psf_gmix() != has_psf:
                raise NotImplementedError(
                    "all stamps of an object must share their size and all "
                    "or none carry a psf mixture")
        images = np.stack([o.image for _, o in stamps])
        weights = np.stack([o.weight for _, o in stamps])
        jac = np.stack([np.asarray(o._jacobian._data.view("f8")).reshape(8)
                        for _, o in stamps])
        bands = np.array([b for b, _ in stamps], dtype="i4")
        psf_pars = None
        if has_psf:
            psf_pars = np.stack([o.psf.gmix.get_full_pars().reshape(-1, 6)
                                 for _, o in stamps])
        nsh = get_model_npars(self.model_name) - 1
        for b in range(nband):
            make_gmix_model(np.append(guess[:nsh], guess[nsh + b]),
                            self.model_name)
        izw = all(o.ignore_zero_weight for _, o in stamps)
        rec = fit_batch_multi(self.model_name, [0, len(stamps)], bands, nband,
                              images, weights, jac, psf_pars, guess[None],
                              prior=self.prior, fit_pars=self.fit_pars,
                              ignore_zero_weight=izw,
                              analytic_jacobian=self.analytic_jacobian)[0]
        if rec["flags"] == 1 and rec["nfev"] == 0:
            raise GMixFatalError("no weights > 0")
        res = record_to_result(rec, self.model_name, npars)
        if self.use_noise_image:
            self._apply_noise_cov(mb, res)
        return FitResult(mb, self.model_name, res)

    def _apply_noise_cov(self, mb, res):
        """calc_noise_cov + apply_noise_cov (fitting/noise_cov.py:38-137): B
        is accumulated over the epochs of every band, each epoch's block
        landing on the shape parameters and its band's flux"""
        from . import noise_cov
        if res["flags"] != 0:
            return
        pars = res["pars"]
        npars, nband = pars.size, len(mb)
        nshape = npars - nband
        B = np.zeros((npars, npars))
        for band, obslist in enumerate(mb):
            kp = list(range(nshape)) + [nshape + band]
            for o in obslist:
                psf_pars = None
                if o.has_psf_gmix():
                    psf_pars = o.psf.gmix.get_full_pars().reshape(1, -1, 6)
                try:
                    Bk = noise_cov.calc_noise_b(
                        self.model_name, pars[kp][None], psf_pars,
                        o._jacobian._data, o.weight[None], o.noise[None])[0]
                except ValueError:
                    Bk = np.full((len(kp), len(kp)), np.nan)
                B[np.ix_(kp, kp)] += Bk
        noise_cov.apply_noise_cov(res, B)
        refresh_result_after_noise_cov(res, nband)


def fit_batch_coellip(ngauss, images, weights, jacobians, guesses,
                      psf_pars=None, prior=None, fit_pars=None,
                      ignore_zero_weight=True):
    """
    Fits of n coelliptical gaussians to n stamps in one launch
    (CoellipFitter.go, fitting/fitters.py:120-141): guesses
    [n, 4 + 2 ngauss] = (cen1, cen2, g1, g2, T_1..T_n, F_1..F_n),
    forward-difference Jacobian as in the reference, ngauss <= 5.  Returns
    records with 14 parameter slots (matrices packed npars x npars).
    """
    fp = dict(DEFAULT_LM_PARS)
    if fit_pars is not None:
        fp.update(fit_pars)
    if getattr(prior, "bounds", None) is not None:
        raise NotImplementedError("bounded coelliptical fits")
    return _lib.lm_fit_coellip(ngauss, images, weights, jacobians, psf_pars,
                               guesses, prior=_prior_record(prior),
                               maxfev=fp["maxfev"], ftol=fp["ftol"],
                               xtol=fp["xtol"],
                               ignore_zero_weight=ignore_zero_weight)


class CoellipFitResult(FitResult):
    """result of CoellipFitter.go: the run_leastsq keys and fit statistics;
    as in the reference (CoellipFitModel._set_flux, results.py:646-650) no
    total flux is derived"""

    def __init__(self, obs, ngauss, result):
        dict.__init__(self, result)
        self.obs = get_mb_obs(obs)
        self.model_name = "coellip"
        self.nband = 1
        self.npars = 4 + 2 * ngauss

    def get_band_pars(self, pars, band):
        return np.array(pars, dtype="f8")

    def get_gmix(self, band=0):
        from .gmix import GMixCoellip
        return GMixCoellip(self["pars"])


class CoellipFitter(Fitter):
    """
    Levenberg-Marquardt fit of ngauss coelliptical gaussians, the psf model
    of fitting/fitters.py:120-141.

    Parameters
    ----------
    ngauss: int
        number of gaussians, at most 5
    prior: PriorCoellipSame, optional
    fit_pars: dict, optional
    """

    def __init__(self, ngauss, prior=None, fit_pars=None):
        if ngauss > 5:
            raise NotImplementedError("at most 5 coelliptical gaussians")
        self._ngauss = ngauss
        self.prior = prior
        self.model_name = "coellip"
        self.use_noise_image = False
        self.analytic_jacobian = False
        self.fit_pars = dict(fit_pars) if fit_pars is not None \
            else dict(DEFAULT_LM_PARS)

    def go(self, obs, guess):
        mb = get_mb_obs(obs)
        assert len(mb) == 1 and len(mb[0]) == 1, \
            "Coellip can only fit one band, one epoch"
        o = mb[0][0]
        npars = 4 + 2 * self._ngauss
        guess = np.array(guess, dtype="f8")
        assert guess.size == npars, \
            "guess has npars=%d, expected %d" % (guess.size, npars)
        psf_pars = None
        if o.has_psf_gmix():
            psf_pars = o.psf.gmix.get_full_pars().reshape(1, -1, 6)
        from .gmix import GMixCoellip
        GMixCoellip(guess)          # a guess outside the domain raises
        rec = fit_batch_coellip(self._ngauss, o.image[None], o.weight[None],
                                o._jacobian._data, guess[None],
                                psf_pars=psf_pars, prior=self.prior,
                                fit_pars=self.fit_pars,
                                ignore_zero_weight=o.ignore_zero_weight)[0]
        if rec["flags"] == 1 and rec["nfev"] == 0:
            raise GMixFatalError("no weights > 0")
        res = record_to_result_n(rec, "coellip", npars)
        for k in ("flux", "flux_err", "flux_cov"):
            res.pop(k, None)
        return CoellipFitResult(o, self._ngauss, res)
