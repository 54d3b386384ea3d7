"""
The priors the LM fitter evaluates on the device, with the reference's class
names and constructor signatures (ngmix/priors/priors.py, multivariate.py,
shape.py; ngmix/joint_prior.py): CenPrior, GPriorBA, FlatPrior, TwoSidedErf,
Normal and the joint PriorSimpleSep / PriorBDFSep.

Each class also evaluates ln(prob) on the host for user code (scalar
bookkeeping), and `to_ngx()` turns a joint prior into the ngx_prior record the
kernels consume.  Sampling helpers and the MCMC-oriented priors are out of
scope.
"""
from math import erf

import numpy as np

from . import _lib
from .defaults import LOWVAL
from .gexceptions import GMixRangeError


class PriorBase(object):
    def __init__(self, rng=None, bounds=None):
        self.bounds = bounds
        self.rng = rng if rng is not None else np.random.RandomState()

    def has_bounds(self):
        return hasattr(self, "bounds") and self.bounds is not None


class FlatPrior(PriorBase):
    """flat between minval and maxval; out of range raises GMixRangeError"""

    def __init__(self, minval, maxval, rng=None):
        super().__init__(rng=rng)
        self.minval = minval
        self.maxval = maxval

    def get_lnprob_scalar(self, val):
        if val < self.minval or val > self.maxval:
            raise GMixRangeError("value %s out of range: [%s,%s]"
                                 % (val, self.minval, self.maxval))
        return 0.0

    def get_prob_scalar(self, val):
        self.get_lnprob_scalar(val)
        return 1.0

    def get_fdiff(self, val):
        return self.get_lnprob_scalar(val)

    def sample(self, nrand=None):
        return self.rng.uniform(self.minval, self.maxval, size=nrand)

    def _ngx(self):
        return _lib.Prior1D(1, 0, self.minval, self.maxval, 0.0, 0.0)


class TwoSidedErf(PriorBase):
    """1 in the middle, error-function roll-off at both ends"""

    def __init__(self, minval, width_at_min, maxval, width_at_max, rng=None):
        super().__init__(rng=rng)
        self.minval = minval
        self.width_at_min = width_at_min
        self.maxval = maxval
        self.width_at_max = width_at_max

    def get_prob_scalar(self, val):
        p1 = 0.5 * erf((self.maxval - val) / self.width_at_max)
        p2 = 0.5 * erf((val - self.minval) / self.width_at_min)
        return p1 + p2

    def get_lnprob_scalar(self, val):
        p = self.get_prob_scalar(val)
        return LOWVAL if p <= 0.0 else np.log(p)

    def get_fdiff(self, val):
        p = -2 * self.get_lnprob_scalar(val)
        return np.sqrt(max(p, 0.0))

    def _ngx(self):
        return _lib.Prior1D(2, 0, self.minval, self.width_at_min, self.maxval,
                            self.width_at_max)


class Normal(PriorBase):
    """gaussian prior with peak ln(prob) = 0"""

    def __init__(self, mean, sigma, rng=None, bounds=None):
        super().__init__(rng=rng, bounds=bounds)
        self.mean = mean
        self.sigma = sigma
        self.sinv = 1.0 / sigma
        self.s2inv = 1.0 / sigma ** 2

    def get_lnprob(self, val):
        diff = self.mean - val
        return -0.5 * diff * diff * self.s2inv

    get_lnprob_scalar = get_lnprob
    get_lnprob_array = get_lnprob

    def get_fdiff(self, val):
        return (val - self.mean) * self.sinv

    def sample(self, nrand=None):
        return self.rng.normal(self.mean, self.sigma, size=nrand)

    def _ngx(self):
        return _lib.Prior1D(3, 0, self.mean, self.sigma, 0.0, 0.0)


class CenPrior(PriorBase):
    """independent gaussians on the two centre coordinates"""

    def __init__(self, cen1, cen2, sigma1, sigma2, rng=None):
        super().__init__(rng=rng)
        self.cen1 = float(cen1)
        self.cen2 = float(cen2)
        self.sigma1 = float(sigma1)
        self.sigma2 = float(sigma2)
        self.s2inv1 = 1.0 / self.sigma1 ** 2
        self.s2inv2 = 1.0 / self.sigma2 ** 2

    def get_lnprob_scalar_sep(self, x1, x2):
        d1 = self.cen1 - x1
        d2 = self.cen2 - x2
        return -0.5 * d1 * d1 * self.s2inv1, -0.5 * d2 * d2 * self.s2inv2

    def get_lnprob_scalar(self, x1, x2):
        a, b = self.get_lnprob_scalar_sep(x1, x2)
        return a + b

    def get_fdiff(self, x1, x2):
        return ((x1 - self.cen1) / self.sigma1, (x2 - self.cen2) / self.sigma2)

    def sample(self, nrand=None):
        return (self.rng.normal(self.cen1, self.sigma1, size=nrand),
                self.rng.normal(self.cen2, self.sigma2, size=nrand))


SimpleGauss2D = CenPrior


class GPriorBA(PriorBase):
    """Bernstein & Armstrong shape prior (1-g^2)^2 exp(-g^2/2 sigma^2)"""

    def __init__(self, sigma=0.3, rng=None, A=1.0):
        super().__init__(rng=rng)
        self.A = A
        self.sigma = sigma
        self.sig2inv = 1.0 / sigma ** 2
        self.gmax = 1.0

    def get_lnprob_scalar2d(self, g1, g2):
        gsq = g1 * g1 + g2 * g2
        omgsq = 1.0 - gsq
        if omgsq <= 0.0:
            raise GMixRangeError("g^2 too big: %s" % gsq)
        return 2 * np.log(omgsq) - 0.5 * gsq * self.sig2inv

    def get_fdiff(self, g1, g2):
        chi2 = -2 * self.get_lnprob_scalar2d(g1, g2)
        return np.sqrt(max(chi2, 0.0))


class PriorSimpleSep(object):
    """separate priors on centre, shape, T and flux (joint_prior.py:5-236)"""

    def __init__(self, cen_prior, g_prior, T_prior, F_prior):
        self.cen_prior = cen_prior
        self.g_prior = g_prior
        self.T_prior = T_prior
        if isinstance(F_prior, list):
            self.nband = len(F_prior)
        else:
            self.nband = 1
            F_prior = [F_prior]
        self.F_priors = F_prior
        self.bounds = None
        for p in [self.T_prior] + self.F_priors:
            if p.has_bounds():
                raise NotImplementedError(
                    "bounded fits are not in scope (SURVEY.md section 7: the "
                    "reference's bounded covariance is unpinned)")

    def get_lnprob_scalar(self, pars):
        lnp = self.cen_prior.get_lnprob_scalar(pars[0], pars[1])
        lnp += self.g_prior.get_lnprob_scalar2d(pars[2], pars[3])
        lnp += self.T_prior.get_lnprob_scalar(pars[4])
        for i, F_prior in enumerate(self.F_priors):
            lnp += F_prior.get_lnprob_scalar(pars[5 + i])
        return lnp

    def fill_fdiff(self, pars, fdiff):
        lnp1, lnp2 = self.cen_prior.get_lnprob_scalar_sep(pars[0], pars[1])
        rows = [lnp1, lnp2,
                self.g_prior.get_lnprob_scalar2d(pars[2], pars[3]),
                self.T_prior.get_lnprob_scalar(pars[4])]
        rows += [F.get_lnprob_scalar(pars[5 + i])
                 for i, F in enumerate(self.F_priors)]
        n = len(rows)
        chi2 = -2 * np.array(rows)
        chi2.clip(min=0.0, max=None, out=chi2)
        fdiff[0:n] = np.sqrt(chi2)
        return n

    def to_ngx(self):
        if self.nband > 3:
            raise NotImplementedError("at most 3 bands (8 parameters)")
        if not isinstance(self.g_prior, GPriorBA):
            raise NotImplementedError("g prior must be GPriorBA")
        pr = _lib.Prior()
        pr.present = 1
        pr.g_kind = 1
        c = self.cen_prior
        pr.cen1, pr.cen2, pr.sigma1, pr.sigma2 = c.cen1, c.cen2, c.sigma1, \
            c.sigma2
        pr.g_sigma = self.g_prior.sigma
        pr.T = self.T_prior._ngx()
        pr.F = self.F_priors[0]._ngx()
        pr.nband = self.nband
        for b in range(1, self.nband):
            pr.F_more[b - 1] = self.F_priors[b]._ngx()
        return pr


class PriorBDFSep(PriorSimpleSep):
    """separate priors for the bdf model (joint_prior.py:484-676)"""

    def __init__(self, cen_prior, g_prior, T_prior, fracdev_prior, F_prior):
        self.fracdev_prior = fracdev_prior
        super().__init__(cen_prior, g_prior, T_prior, F_prior)
        if fracdev_prior.has_bounds():
            raise NotImplementedError("bounded fits are not in scope")

    def get_lnprob_scalar(self, pars):
        lnp = self.cen_prior.get_lnprob_scalar(pars[0], pars[1])
        lnp += self.g_prior.get_lnprob_scalar2d(pars[2], pars[3])
        lnp += self.T_prior.get_lnprob_scalar(pars[4])
        lnp += self.fracdev_prior.get_lnprob_scalar(pars[5])
        for i, F_prior in enumerate(self.F_priors):
            lnp += F_prior.get_lnprob_scalar(pars[6 + i])
        return lnp

    def fill_fdiff(self, pars, fdiff):
        f1, f2 = self.cen_prior.get_fdiff(pars[0], pars[1])
        rows = [f1, f2, self.g_prior.get_fdiff(pars[2], pars[3]),
                self.T_prior.get_fdiff(pars[4]),
                self.fracdev_prior.get_fdiff(pars[5])]
        rows += [F.get_fdiff(pars[6 + i]) for i, F in enumerate(self.F_priors)]
        fdiff[0:len(rows)] = rows
        return len(rows)

    def to_ngx(self):
        pr = super().to_ngx()
        pr.fracdev = self.fracdev_prior._ngx()
        return pr


class PriorBDSep(PriorSimpleSep):
    """separate priors for the bd model, pars [cen1, cen2, g1, g2, T,
    logTratio, fracdev, F] (joint_prior.py:267-481)"""

    def __init__(self, cen_prior, g_prior, T_prior, logTratio_prior,
                 fracdev_prior, F_prior):
        self.logTratio_prior = logTratio_prior
        self.fracdev_prior = fracdev_prior
        super().__init__(cen_prior, g_prior, T_prior, F_prior)
        if logTratio_prior.has_bounds() or fracdev_prior.has_bounds():
            raise NotImplementedError("bounded fits are not in scope")

    def get_lnprob_scalar(self, pars):
        lnp = self.cen_prior.get_lnprob_scalar(pars[0], pars[1])
        lnp += self.g_prior.get_lnprob_scalar2d(pars[2], pars[3])
        lnp += self.T_prior.get_lnprob_scalar(pars[4])
        lnp += self.logTratio_prior.get_lnprob_scalar(pars[5])
        lnp += self.fracdev_prior.get_lnprob_scalar(pars[6])
        for i, F_prior in enumerate(self.F_priors):
            lnp += F_prior.get_lnprob_scalar(pars[7 + i])
        return lnp

    def fill_fdiff(self, pars, fdiff):
        f1, f2 = self.cen_prior.get_fdiff(pars[0], pars[1])
        rows = [f1, f2, self.g_prior.get_fdiff(pars[2], pars[3]),
                self.T_prior.get_fdiff(pars[4]),
                self.logTratio_prior.get_fdiff(pars[5]),
                self.fracdev_prior.get_fdiff(pars[6])]
        rows += [F.get_fdiff(pars[7 + i]) for i, F in enumerate(self.F_priors)]
        fdiff[0:len(rows)] = rows
        return len(rows)

    def to_ngx(self):
        if self.nband != 1:
            raise NotImplementedError("bd fits are single band")
        pr = super().to_ngx()
        pr.logTratio = self.logTratio_prior._ngx()
        pr.fracdev = self.fracdev_prior._ngx()
        return pr
