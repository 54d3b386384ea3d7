"""
Guess generators (ngmix/guessers.py).  Host-side: a guess is a handful of
random draws; what matters for a drop-in is that the draws come from the
caller's RandomState in the reference's order, so a pipeline seeded the same
way starts its fits from the same points.  The psf fluxes behind
TPSFFluxGuesser come from the device (ngx_template_flux).

Built: TFluxGuesser, TPSFFluxGuesser, ParsGuesser (each with the optional
prior fix-up), TFluxAndPriorGuesser, TPSFFluxAndPriorGuesser, BDFGuesser,
BDFPSFFluxGuesser, BDGuesser, PriorGuesser, GMixPSFGuesser, SimplePSFGuesser.
CoellipPSFGuesser.  Not built: the galsim-model guessers (R50*).
"""
import numpy as np

from . import moments
from .defaults import LOWVAL
from .gexceptions import GMixRangeError, PSFFluxFailure
from .gmix import GMix, GMixModel
from .observation import get_mb_obs


def _cen_shape_T(rng, T, nrand, guess):
    """the draws every T/flux guesser starts with, in the reference's order
    (guessers.py:57-61): cen1, cen2, g1, g2, T"""
    guess[:, 0] = rng.uniform(low=-0.01, high=0.01, size=nrand)
    guess[:, 1] = rng.uniform(low=-0.01, high=0.01, size=nrand)
    guess[:, 2] = rng.uniform(low=-0.02, high=0.02, size=nrand)
    guess[:, 3] = rng.uniform(low=-0.02, high=0.02, size=nrand)
    guess[:, 4] = T * rng.uniform(low=0.9, high=1.1, size=nrand)


def _fluxes(rng, fluxes, nrand, guess):
    for band, flux in enumerate(fluxes):
        guess[:, 5 + band] = flux * rng.uniform(low=0.9, high=1.1, size=nrand)


def _allowed(prior, pars):
    """does the prior accept this parameter vector"""
    try:
        return prior.get_lnprob_scalar(pars) > LOWVAL
    except GMixRangeError:
        return False


def _fix_guess(guess, prior, ntry=4):
    """rows the prior rejects are replaced by draws from the prior, up to
    ntry times each (guessers.py:1275-1299)"""
    for row in guess:
        for _ in range(ntry):
            if _allowed(prior, row):
                break
            row[:] = prior.sample()


def _fix_guess_TFlux(guess, prior, ntry=4):
    """the same, replacing only T and what follows it
    (guessers.py:1245-1272)"""
    for row in guess:
        for _ in range(ntry):
            if _allowed(prior, row):
                break
            row[4:] = prior.sample()[4:]


class TFluxGuesser(object):
    """guesses from a T and flux(es): centre and shape tight around zero,
    T and fluxes within 10% (guessers.py:14-75)"""

    def __init__(self, rng, T, flux, prior=None):
        self.rng = rng
        self.T = T
        self.fluxes = np.array(flux, dtype="f8", ndmin=1)
        self.prior = prior

    def __call__(self, nrand=1, obs=None):
        guess = np.zeros((nrand, 5 + self.fluxes.size))
        _cen_shape_T(self.rng, self.T, nrand, guess)
        _fluxes(self.rng, self.fluxes, nrand, guess)
        if self.prior is not None:
            _fix_guess(guess, self.prior)
        return guess[0] if nrand == 1 else guess


class TPSFFluxGuesser(object):
    """like TFluxGuesser with the fluxes taken from psf template fits of the
    observations (guessers.py:78-144); the fluxes are cached per obs"""

    def __init__(self, rng, T, prior=None):
        self.rng = rng
        self.T = T
        self.prior = prior
        self._id_last = None
        self._psf_fluxes = None

    def _get_psf_fluxes(self, obs):
        if id(obs) != self._id_last:
            self._id_last = id(obs)
            self._psf_fluxes = get_psf_fluxes(rng=self.rng, obs=obs)["flux"]
        return self._psf_fluxes

    def __call__(self, obs, nrand=1):
        fluxes = self._get_psf_fluxes(obs)
        guess = np.zeros((nrand, 5 + fluxes.size))
        _cen_shape_T(self.rng, self.T, nrand, guess)
        _fluxes(self.rng, fluxes, nrand, guess)
        if self.prior is not None:
            _fix_guess(guess, self.prior)
        return guess[0] if nrand == 1 else guess


def _prior_start(prior, rng, T, nrand):
    """a draw of every parameter from the prior with T replaced by the
    caller's T within 10%: the common start of the *AndPrior / BDF / BD
    guessers (guessers.py:183-190, 311-314)"""
    guess = prior.sample(nrand)
    guess[:, 4] = T * (1.0 + rng.uniform(low=-0.1, high=0.1, size=nrand))
    return guess


class TPSFFluxAndPriorGuesser(TPSFFluxGuesser):
    """centre and shape from the prior, T within 10% of the input, fluxes
    within 10% of the psf fluxes (guessers.py:147-202)"""

    def __init__(self, rng, T, prior):
        super().__init__(rng, T, prior=prior)

    def __call__(self, obs, nrand=1):
        fluxes = self._get_psf_fluxes(obs)
        guess = _prior_start(self.prior, self.rng, self.T, nrand)
        _fluxes(self.rng, fluxes, nrand, guess)
        _fix_guess_TFlux(guess, self.prior)
        return guess[0] if nrand == 1 else guess


def _scatter_fluxes(rng, fluxes, nrand, guess, first):
    for band, flux in enumerate(fluxes):
        r = rng.uniform(low=-0.1, high=0.1, size=nrand)
        guess[:, first + band] = flux * (1.0 + r)


class TFluxAndPriorGuesser(object):
    """centre and shape from the prior, T and fluxes within 10% of the
    inputs; draws come from the centre prior's rng (guessers.py:264-322)"""

    def __init__(self, rng, T, flux, prior):
        self.T = T
        self.fluxes = np.array(flux, dtype="f8", ndmin=1)
        self.prior = prior

    def __call__(self, nrand=1, obs=None):
        rng = self.prior.cen_prior.rng
        guess = _prior_start(self.prior, rng, self.T, nrand)
        _scatter_fluxes(rng, self.fluxes, nrand, guess, 5)
        _fix_guess_TFlux(guess, self.prior)
        return guess[0] if nrand == 1 else guess


class BDFGuesser(object):
    """bdf guesses: the prior's draw with T and fluxes within 10% of the
    inputs and fracdev in [0.4, 0.6] (guessers.py:379-429)"""
    _first_flux = 6

    def __init__(self, T, flux, prior):
        self.T = T
        self.fluxes = np.array(flux, ndmin=1)
        self.prior = prior

    def _get_fluxes(self, obs):
        return self.fluxes

    def __call__(self, nrand=1, obs=None):
        rng = self.prior.cen_prior.rng
        fluxes = self._get_fluxes(obs)
        guess = _prior_start(self.prior, rng, self.T, nrand)
        # the reference writes this draw into column 5 for bd as well (where
        # column 5 is log10(Td/Te)); kept as is
        guess[:, 5] = rng.uniform(low=0.4, high=0.6, size=nrand)
        _scatter_fluxes(rng, fluxes, nrand, guess, self._first_flux)
        _fix_guess(guess, self.prior)
        return guess[0] if nrand == 1 else guess


class BDGuesser(BDFGuesser):
    """bd guesses, fluxes from column 7 (guessers.py:432-486)"""
    _first_flux = 7


class BDFPSFFluxGuesser(BDFGuesser, TPSFFluxGuesser):
    """bdf guesses with the fluxes from psf template fits
    (guessers.py:325-376)"""

    def __init__(self, T, prior):
        self.T = T
        self.prior = prior
        self.rng = prior.cen_prior.rng
        self._id_last = None
        self._psf_fluxes = None

    def _get_fluxes(self, obs):
        return self._get_psf_fluxes(obs)

    def __call__(self, obs, nrand=1):
        return BDFGuesser.__call__(self, nrand=nrand, obs=obs)


class PriorGuesser(object):
    """guesses that are plain draws from a joint prior
    (guessers.py:667-686)"""

    def __init__(self, prior):
        self.prior = prior

    def __call__(self, obs=None, nrand=None):
        return self.prior.sample(nrand)


def get_psf_fluxes(rng, obs):
    """psf flux of every band (guessers.py:205-261): flags, flux, flux_err
    arrays; bands whose fit failed get the mean of the good ones within 10%;
    PSFFluxFailure when none is good"""
    from .fitting import PSFFluxFitter
    mbobs = get_mb_obs(obs)
    nband = len(mbobs)
    flux, flux_err = np.zeros(nband), np.zeros(nband)
    flags = np.zeros(nband, dtype="i4")
    fitter = PSFFluxFitter()
    for band, obslist in enumerate(mbobs):
        res = fitter.go(obs=obslist)
        flags[band], flux[band] = res["flags"], res["flux"]
        flux_err[band] = res["flux_err"]
    good = (flags == 0) & np.isfinite(flux)
    if not good.all():
        if not good.any():
            raise PSFFluxFailure("no good psf fluxes")
        nbad = int((~good).sum())
        fac = 1.0 + rng.uniform(low=-0.1, high=0.1, size=nbad)
        flux[~good] = flux[good].mean() * fac
    return {"flags": flags, "flux": flux, "flux_err": flux_err}


class ParsGuesser(object):
    """guesses scattered around a parameter vector (guessers.py:489-567):
    centres and shapes by absolute widths, the rest by fractional widths"""

    def __init__(self, rng, pars, prior=None, widths=None):
        self.rng = rng
        self.prior = prior
        self.pars = np.array(pars, dtype="f8")
        self.np = self.pars.size
        if widths is None:
            self.widths = self.pars * 0 + 0.1
            self.widths[0:2] = 0.02
        else:
            self.widths = np.array(widths, dtype="f8")

    def __call__(self, nrand=None, obs=None):
        rng, pars, widths = self.rng, self.pars, self.widths
        is_scalar = nrand is None
        n = 1 if is_scalar else nrand
        guess = np.zeros((n, self.np))
        guess[:, 0] = pars[0] + widths[0] * _srandu(rng, n)
        guess[:, 1] = pars[1] + widths[1] * _srandu(rng, n)
        g1, g2 = _shape_guess(rng, pars[2], pars[3], n, widths[2:4],
                              gmax=0.8)
        guess[:, 2], guess[:, 3] = g1, g2
        for i in range(4, self.np):
            guess[:, i] = pars[i] * (1.0 + widths[i] * _srandu(rng, n))
        if self.prior is not None:
            _fix_guess(guess, self.prior)
        return guess[0] if is_scalar else guess


def _srandu(rng, n):
    return rng.uniform(low=-1.0, high=1.0, size=n)


def _shape_guess(rng, g1, g2, n, width, gmax=0.99):
    """shapes near (g1, g2): the (rescaled, if |g| > gmax) shape sheared by
    small random offsets, redrawn while the result is out of range
    (guessers.py:570-599)"""
    from .gexceptions import GMixRangeError
    from .shape import Shape
    g = np.sqrt(g1 ** 2 + g2 ** 2)
    if g > gmax:
        g1, g2 = g1 * (gmax / g), g2 * (gmax / g)
    shape = Shape(g1, g2)
    out1, out2 = np.zeros(n), np.zeros(n)
    for i in range(n):
        while True:
            try:
                off1 = width[0] * rng.uniform(low=-1.0, high=1.0)
                off2 = width[1] * rng.uniform(low=-1.0, high=1.0)
                new = shape.get_sheared(off1, off2)
                break
            except GMixRangeError:
                pass
        out1[i], out2[i] = new.g1, new.g2
    return out1, out2


# component (amplitude, size) fractions of the psf mixtures by ngauss, and
# which table entry each gaussian uses (the reference reuses entry 2 for the
# trailing gaussians of its 4 and 5 component guesses, guessers.py:951-1051)
_PSF_P = {
    2: [0.596510042804182, 0.4034898268889178],
    3: [0.596510042804182, 0.4034898268889178, 1.303069003078001e-07],
    4: [0.596510042804182, 0.4034898268889178, 1.303069003078001e-07, 1.0e-8],
    5: [0.59453032, 0.35671819, 0.03567182, 0.01189061, 0.00118906],
}
_PSF_F = {
    2: [0.5793612389470884, 1.621860687127999],
    3: [0.5793612389470884, 1.621860687127999, 7.019347162356363],
    4: [0.5793612389470884, 1.621860687127999, 7.019347162356363, 16.0],
    5: [0.5, 1.0, 3.0, 10.0, 20.0],
}
_PSF_ENTRY = {2: [0, 1], 3: [0, 1, 2], 4: [0, 1, 2, 2], 5: [0, 1, 2, 2, 2]}


class GMixPSFGuesser(object):
    """full gaussian-mixture guesses for psf fits with EM / admom
    (guessers.py:767-1051): flux and T from the image sum and 3.5 pixels of
    FWHM, or from deweighted gaussian moments with guess_from_moms"""

    def __init__(self, rng, ngauss, guess_from_moms=False):
        if ngauss not in (1, 2, 3, 4, 5):
            raise ValueError("bad ngauss: %d" % ngauss)
        self.rng = rng
        self.ngauss = ngauss
        self.guess_from_moms = guess_from_moms

    def __call__(self, obs):
        T, flux = self._get_T_flux(obs)
        return GMix(pars=self._draw_pars(flux, T))

    def _get_T_flux(self, obs):
        if self.guess_from_moms:
            return self._get_T_flux_from_moms(obs)
        return self._get_T_flux_default(obs)

    def _get_T_flux_default(self, obs):
        scale = obs.jacobian.scale
        return moments.fwhm_to_T(scale * 3.5), obs.image.sum()

    def _get_T_flux_from_moms(self, obs):
        scale = obs.jacobian.scale
        Tweight = moments.fwhm_to_T(scale * 3.5)
        wt = GMixModel([0.0, 0.0, 0.0, 0.0, Tweight, 1.0], "gauss")
        res = wt.get_weighted_moments(obs=obs, maxrad=1.0e9)
        if res["flags"] != 0:
            return self._get_T_flux_default(obs)
        Tmeas = res["T"]
        if moments.T_to_fwhm(Tmeas) < scale:
            return self._get_T_flux_default(obs)
        # deweight as if the profile were gaussian; ngmix is in flux units
        T = 1.0 / (1 / Tmeas - 1 / Tweight)
        flux = res["flux"] * np.pi * (Tweight + T) / scale ** 2
        if T < 0:
            T, flux = res["T"], res["flux"]
        return T, flux

    def _draw_pars(self, flux, T):
        """six numbers per gaussian (p, row, col, irr, irc, icc), drawn in
        that order"""
        rng = self.rng
        sigma2 = T / 2

        def u(lo, hi):
            return rng.uniform(low=lo, high=hi)

        n = self.ngauss
        if n == 1:
            return np.array([
                flux * u(0.9, 1.1), u(-0.1, 0.1), u(-0.1, 0.1),
                sigma2 * (1.0 + u(-0.1, 0.1)), u(-0.2 * sigma2, 0.2 * sigma2),
                sigma2 * (1.0 + u(-0.1, 0.1))])
        pars = np.zeros(6 * n)
        for i, e in enumerate(_PSF_ENTRY[n]):
            pfrac, ffrac = _PSF_P[n][e], _PSF_F[n][e]
            if n == 2:
                # fixed amplitudes and no cross term for two gaussians
                p = pfrac * flux
            else:
                p = flux * pfrac * (1.0 + u(-0.1, 0.1))
            row, col = u(-0.1, 0.1), u(-0.1, 0.1)
            irr = ffrac * sigma2 * (1.0 + u(-0.1, 0.1))
            irc = 0.0 if n == 2 else u(-0.01, 0.01)
            icc = ffrac * sigma2 * (1.0 + u(-0.1, 0.1))
            pars[6 * i:6 * i + 6] = (p, row, col, irr, irc, icc)
        return pars


class SimplePSFGuesser(GMixPSFGuesser):
    """[cen1, cen2, g1, g2, T, flux] guesses for simple-model psf fits
    (guessers.py:1054-1103)"""

    def __init__(self, rng, guess_from_moms=False):
        self.rng = rng
        self.guess_from_moms = guess_from_moms
        self.npars = 6

    def __call__(self, obs):
        rng = self.rng
        T, flux = self._get_T_flux(obs)
        guess = np.zeros(self.npars)
        guess[0:2] += rng.uniform(low=-0.01, high=0.01, size=2)
        guess[2:4] += rng.uniform(low=-0.05, high=0.05, size=2)
        guess[4] = T * rng.uniform(low=0.9, high=1.1)
        guess[5] = flux * rng.uniform(low=0.9, high=1.1)
        return guess


# size and amplitude fractions of the coelliptical guesses, by ngauss
# (moffat-like profiles, guessers.py:1219-1242; the later of the reference's
# two 5-gaussian tables is the one in effect)
_COELLIP_P = {
    1: [1.0],
    2: [0.5, 0.5],
    3: [0.27559669, 0.55817131, 0.166232],
    4: [0.44534, 0.366951, 0.10506, 0.0826497],
    5: [0.57874897, 0.32273483, 0.03327272, 0.0341253, 0.03111819],
}
_COELLIP_F = {
    1: [1.0],
    2: [0.48955064, 1.50658978],
    3: [0.36123609, 0.8426139, 2.58747785],
    4: [0.541019, 1.19701, 0.282176, 3.51086],
    5: [0.27831284, 0.9959897, 5.86989779, 5.63590429, 4.17285878],
}


class CoellipPSFGuesser(GMixPSFGuesser):
    """[cen1, cen2, g1, g2, T_1..T_n, F_1..F_n] guesses for coelliptical psf
    fits (guessers.py:1106-1216): centre and shape near zero, sizes and
    fluxes the table fractions of T and flux within 1 % (10 % for one
    gaussian), drawn T's first, then fluxes"""

    def __init__(self, rng, ngauss, guess_from_moms=False):
        super().__init__(rng, ngauss, guess_from_moms=guess_from_moms)
        self.npars = 4 + 2 * ngauss

    def __call__(self, obs):
        rng, n = self.rng, self.ngauss
        T, flux = self._get_T_flux(obs)
        guess = np.zeros(self.npars)
        guess[0:2] += rng.uniform(low=-0.01, high=0.01, size=2)
        guess[2:4] += rng.uniform(low=-0.05, high=0.05, size=2)
        lo, hi = (0.9, 1.1) if n == 1 else (0.99, 1.01)
        for i in range(n):
            guess[4 + i] = T * _COELLIP_F[n][i] * rng.uniform(low=lo, high=hi)
        for i in range(n):
            guess[4 + n + i] = flux * _COELLIP_P[n][i] \
                * rng.uniform(low=lo, high=hi)
        return guess
