"""
The priors the LM fitter evaluates on the device, with the reference's class
names and constructor signatures (ngmix/priors/priors.py, multivariate.py,
shape.py; ngmix/joint_prior.py): CenPrior, GPriorBA, FlatPrior, TwoSidedErf,
Normal and the joint PriorSimpleSep / PriorBDFSep.

Each class also evaluates ln(prob) on the host for user code (scalar
bookkeeping), and `to_ngx()` turns a joint prior into the ngx_prior record the
kernels consume.  `sample()` draws from the caller's RandomState in the
reference's order (the guessers that start a fit from the prior depend on it);
the MCMC-oriented priors are out of scope.
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
        """uniform on the interval (priors/priors.py:159-187)"""
        u = self.rng.uniform(size=1 if nrand is None else nrand)
        vals = self.minval + (self.maxval - self.minval) * u
        return vals[0] if nrand is None else vals

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

    def get_prob_array(self, vals):
        vals = np.atleast_1d(np.asarray(vals, dtype="f8"))
        return np.array([self.get_prob_scalar(v) for v in vals])

    def sample(self, nrand=None):
        """rejection sampling between 5 widths either side
        (priors/priors.py:339-388): candidates and heights are drawn in
        rounds of the number still missing"""
        n = 1 if nrand is None else nrand
        lo = self.minval - 5.0 * self.width_at_min
        hi = self.maxval + 5.0 * self.width_at_max
        out = np.zeros(n)
        have = 0
        while have < n:
            need = n - have
            x = self.rng.uniform(low=lo, high=hi, size=need)
            p = self.get_prob_array(x)
            h = self.rng.uniform(size=need)
            keep = x[h < p]
            out[have:have + keep.size] = keep
            have += keep.size
        return out[0] if nrand is None else out

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


class LogNormal(PriorBase):
    """log-normal prior in terms of the mean and sigma of the distribution,
    peak ln(prob) = 0, optional shift (priors/priors.py:674-863)"""

    def __init__(self, mean, sigma, rng=None, shift=None):
        super().__init__(rng=rng)
        if mean <= 0:
            raise ValueError("mean %s is < 0" % mean)
        self.shift = shift
        self.mean = mean
        self.sigma = sigma
        self.logvar = np.log(1 + sigma ** 2 / mean ** 2)
        self.logmean = np.log(mean) - 0.5 * self.logvar
        self.logsigma = np.sqrt(self.logvar)
        self.logivar = 1.0 / self.logvar
        self.log_mode = self.logmean - self.logvar
        self.mode = np.exp(self.log_mode)
        chi2 = self.logivar * (self.log_mode - self.logmean) ** 2
        self.lnprob_max = -0.5 * chi2 - self.log_mode

    def get_lnprob_scalar(self, val):
        if self.shift is not None:
            val = val - self.shift
        if val <= 0:
            raise GMixRangeError("values of val must be > 0")
        logval = np.log(val)
        chi2 = self.logivar * (logval - self.logmean) ** 2
        return -0.5 * chi2 - logval - self.lnprob_max

    def get_prob_scalar(self, val):
        return np.exp(self.get_lnprob_scalar(val))

    def get_fdiff(self, val):
        return np.sqrt(max(-2 * self.get_lnprob_scalar(val), 0.0))

    def sample(self, nrand=None):
        r = np.exp(self.logmean + self.logsigma * self.rng.normal(size=nrand))
        return r if self.shift is None else r + self.shift

    def _ngx(self):
        return _lib.Prior1D(4, 0, self.logmean, self.logivar, self.lnprob_max,
                            0.0 if self.shift is None else self.shift)


class TruncatedGaussian(PriorBase):
    """gaussian restricted to [minval, maxval]; outside raises
    GMixRangeError (priors/priors.py:1046-1150)"""

    def __init__(self, mean, sigma, minval, maxval, rng=None):
        super().__init__(rng=rng)
        self.mean = mean
        self.sigma = sigma
        self.ivar = 1.0 / sigma ** 2
        self.sinv = 1.0 / sigma
        self.minval = minval
        self.maxval = maxval

    def _check(self, val):
        if val < self.minval or val > self.maxval:
            raise GMixRangeError("value out of range")

    def get_lnprob_scalar(self, val):
        self._check(val)
        diff = val - self.mean
        return -0.5 * diff * diff * self.ivar

    def get_fdiff(self, val):
        self._check(val)
        return (val - self.mean) * self.sinv

    def sample(self, nrand=None):
        """normal draws, keeping those strictly inside the range, in rounds
        of the number still missing (priors/priors.py:1128-1150)"""
        n = 1 if nrand is None else nrand
        out = np.zeros(n)
        have = 0
        while have < n:
            t = self.rng.normal(loc=self.mean, scale=self.sigma,
                                size=n - have)
            keep = t[(t > self.minval) & (t < self.maxval)]
            out[have:have + keep.size] = keep
            have += keep.size
        return out[0] if nrand is None else out

    def _ngx(self):
        return _lib.Prior1D(5, 0, self.mean, self.sigma, self.minval,
                            self.maxval)


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

    def get_prob_array1d(self, g):
        """p(|g|) = A 2 pi g (1-g^2)^2 exp(-g^2/2 sigma^2), 0 for g >= 1
        (priors/shape.py:608-640)"""
        g = np.atleast_1d(np.asarray(g, dtype="f8"))
        out = np.zeros(g.size)
        gsq = g * g
        omgsq = 1.0 - gsq
        w, = np.where(omgsq > 0.0)
        if w.size > 0:
            omgsq *= omgsq
            out[w] = omgsq[w] * np.exp(-0.5 * gsq[w] * self.sig2inv)
            out[w] *= self.A * 2 * np.pi * g[w]
        return out

    def get_prob_scalar1d(self, g):
        return float(self.get_prob_array1d(g)[0])

    def _set_maxval1d(self):
        """peak of p(|g|).  The reference finds it with a numerical minimiser
        started at sigma (shape.py:294-311); the stationary point of
        log p = log g + 2 log(1-g^2) - g^2/2 sigma^2 is the root in (0,1) of
        1/g - 4g/(1-g^2) - g/sigma^2, bracketed and bisected here."""
        f = lambda g: 1.0 / g - 4.0 * g / (1.0 - g * g) - g * self.sig2inv
        lo, hi = 1.0e-12, 1.0 - 1.0e-12
        for _ in range(200):
            mid = 0.5 * (lo + hi)
            if f(mid) > 0.0:
                lo = mid
            else:
                hi = mid
        self.maxval1d_loc = 0.5 * (lo + hi)
        self.maxval1d = self.get_prob_scalar1d(self.maxval1d_loc)

    def sample1d(self, nrand, maxguess=None):
        """|g| by rejection under 1.1 x the peak, candidates in [0, gmax-1e-4)
        (priors/shape.py:160-209,471-491).  The reference draws one number
        for its minimiser's starting point on every call; that draw is kept
        so the caller's random stream stays aligned."""
        rng = self.rng
        if maxguess is None:
            rng.uniform(low=-1.0, high=1.0)
        if not hasattr(self, "maxval1d"):
            self._set_maxval1d()
        top = self.maxval1d * 1.1
        gmax = self.gmax - 1.0e-4
        out = np.zeros(nrand)
        have = 0
        while have < nrand:
            need = nrand - have
            g = gmax * rng.uniform(size=need)
            h = top * rng.uniform(size=need)
            keep = g[h < self.get_prob_array1d(g)]
            out[have:have + keep.size] = keep
            have += keep.size
        return out

    def sample2d(self, nrand=None, maxguess=0.1):
        """(g1, g2): |g| from sample1d and a uniform position angle
        (priors/shape.py:211-248; its default maxguess is a number, so no
        starting-point draw is made on this path)"""
        n = 1 if nrand is None else nrand
        g = self.sample1d(n, maxguess=maxguess)
        twotheta = 2 * (self.rng.uniform(size=n) * 2 * np.pi)
        g1, g2 = g * np.cos(twotheta), g * np.sin(twotheta)
        if nrand is None:
            return g1[0], g2[0]
        return g1, g2


class PriorSimpleSep(object):
    """separate priors on centre, shape, T and flux (joint_prior.py:5-236)"""

    def __init__(self, cen_prior, g_prior, T_prior, F_prior):
        self.cen_prior = cen_prior
        self.g_prior = g_prior
        self.T_prior = T_prior
        if isinstance(F_prior, (list, tuple)):
            self.nband = len(F_prior)
        else:
            self.nband = 1
            F_prior = [F_prior]
        self.F_priors = F_prior
        self.set_bounds()

    def set_bounds(self):
        """prior.bounds of the reference (joint_prior.py:37-61): (None, None)
        for the centre and shape, then the (min, max) of every 1-d member
        that has bounds; None when no member has any.  A fit with bounds
        runs leastsqbound's transformed parameters (ngx_lm_pars.nbounds)."""
        bounds = [(None, None)] * 4
        some = False
        for p in [self.T_prior] + self._middle_priors() + self.F_priors:
            if p.has_bounds():
                some = True
                bounds.append((p.bounds[0], p.bounds[1]))
            else:
                bounds.append((None, None))
        self.bounds = bounds if some else None

    def get_lnprob_scalar(self, pars):
        lnp = self.cen_prior.get_lnprob_scalar(pars[0], pars[1])
        lnp += self.g_prior.get_lnprob_scalar2d(pars[2], pars[3])
        lnp += self.T_prior.get_lnprob_scalar(pars[4])
        for i, F_prior in enumerate(self.F_priors):
            lnp += F_prior.get_lnprob_scalar(pars[5 + i])
        return lnp

    # the 1-d priors between T and the fluxes, in parameter order
    def _middle_priors(self):
        return []

    def sample(self, nrand=None):
        """one draw per component in parameter order: centre, shape, T, the
        model's extra parameters, one flux per band (joint_prior.py:191-226,
        425-461, 621-655); [npars] or [nrand, npars]"""
        n = 1 if nrand is None else nrand
        cols = list(self.cen_prior.sample(n))
        cols += list(self.g_prior.sample2d(n))
        cols.append(self.T_prior.sample(n))
        cols += [p.sample(n) for p in self._middle_priors()]
        cols += [F.sample(n) for F in self.F_priors]
        samples = np.stack(cols, axis=1)
        return samples[0] if nrand is None else samples

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
        if self.nband > 12:
            raise NotImplementedError("at most 12 bands")
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
            if b <= 2:
                pr.F_more[b - 1] = self.F_priors[b]._ngx()
            elif b <= 6:
                pr.F_extra[b - 3] = self.F_priors[b]._ngx()
            else:
                pr.F_xwide[b - 7] = self.F_priors[b]._ngx()
        return pr


class PriorBDFSep(PriorSimpleSep):
    """separate priors for the bdf model (joint_prior.py:484-676)"""

    def __init__(self, cen_prior, g_prior, T_prior, fracdev_prior, F_prior):
        self.fracdev_prior = fracdev_prior
        super().__init__(cen_prior, g_prior, T_prior, F_prior)

    def get_lnprob_scalar(self, pars):
        lnp = self.cen_prior.get_lnprob_scalar(pars[0], pars[1])
        lnp += self.g_prior.get_lnprob_scalar2d(pars[2], pars[3])
        lnp += self.T_prior.get_lnprob_scalar(pars[4])
        lnp += self.fracdev_prior.get_lnprob_scalar(pars[5])
        for i, F_prior in enumerate(self.F_priors):
            lnp += F_prior.get_lnprob_scalar(pars[6 + i])
        return lnp

    def _middle_priors(self):
        return [self.fracdev_prior]

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

    def get_lnprob_scalar(self, pars):
        lnp = self.cen_prior.get_lnprob_scalar(pars[0], pars[1])
        lnp += self.g_prior.get_lnprob_scalar2d(pars[2], pars[3])
        lnp += self.T_prior.get_lnprob_scalar(pars[4])
        lnp += self.logTratio_prior.get_lnprob_scalar(pars[5])
        lnp += self.fracdev_prior.get_lnprob_scalar(pars[6])
        for i, F_prior in enumerate(self.F_priors):
            lnp += F_prior.get_lnprob_scalar(pars[7 + i])
        return lnp

    def _middle_priors(self):
        return [self.logTratio_prior, self.fracdev_prior]

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
        pr = super().to_ngx()
        pr.logTratio = self.logTratio_prior._ngx()
        pr.fracdev = self.fracdev_prior._ngx()
        return pr


class PriorCoellipSame(PriorSimpleSep):
    """priors for a coelliptical mixture: centre, shape, the same T prior on
    every T_i and the same flux prior on every F_i (joint_prior.py:874-1031)"""

    def __init__(self, ngauss, cen_prior, g_prior, T_prior, F_prior):
        self.ngauss = ngauss
        self.npars = 4 + 2 * ngauss
        super().__init__(cen_prior, g_prior, T_prior, F_prior)
        if self.nband != 1:
            raise ValueError("coellip only supports one band")

    def set_bounds(self):
        """(joint_prior.py:887-911) the T and flux bounds repeated for every
        gaussian; bounded coelliptical fits are not run on the device"""
        n = getattr(self, "ngauss", 1)
        bounds = [(None, None)] * 4
        some = False
        for p in [self.T_prior] * n + [self.F_priors[0]] * n:
            if p.has_bounds():
                some = True
                bounds.append((p.bounds[0], p.bounds[1]))
            else:
                bounds.append((None, None))
        self.bounds = bounds if some else None

    def _check(self, pars):
        if len(pars) != self.npars:
            raise ValueError("pars size %d expected %d"
                             % (len(pars), self.npars))

    def _lnp_rows(self, pars):
        n = self.ngauss
        lnp1, lnp2 = self.cen_prior.get_lnprob_scalar_sep(pars[0], pars[1])
        rows = [lnp1, lnp2,
                self.g_prior.get_lnprob_scalar2d(pars[2], pars[3])]
        rows += [self.T_prior.get_lnprob_scalar(pars[4 + i])
                 for i in range(n)]
        rows += [self.F_priors[0].get_lnprob_scalar(pars[4 + n + i])
                 for i in range(n)]
        return rows

    def get_lnprob_scalar(self, pars):
        self._check(pars)
        return float(np.sum(self._lnp_rows(pars)))

    def fill_fdiff(self, pars, fdiff):
        self._check(pars)
        rows = self._lnp_rows(pars)
        chi2 = -2 * np.array(rows)
        chi2.clip(min=0.0, max=None, out=chi2)
        fdiff[0:len(rows)] = np.sqrt(chi2)
        return len(rows)

    def sample(self, nrand=None):
        """(joint_prior.py:992-1031; as there, the first size column holds
        the sum of two T draws)"""
        n = 1 if nrand is None else nrand
        ng = self.ngauss
        samples = np.zeros((n, self.npars))
        samples[:, 0], samples[:, 1] = self.cen_prior.sample(n)
        samples[:, 2], samples[:, 3] = self.g_prior.sample2d(n)
        samples[:, 4] = self.T_prior.sample(n)
        for i in range(ng):
            samples[:, 4 + i] += self.T_prior.sample(n)
        for i in range(ng):
            samples[:, 4 + ng + i] = self.F_priors[0].sample(n)
        return samples[0] if nrand is None else samples

    def to_ngx(self):
        pr = super().to_ngx()
        pr.ncoellip = self.ngauss
        return pr
