"""
Host-side shape algebra used by the GMix helpers (subset of ngmix/shape.py
and ngmix/moments.py: g <-> e conversion, reduced-shear addition, sheared
second moments, unit conversions).  Scalar bookkeeping, not on the hot path.
"""
import numpy as np

from .gexceptions import GMixRangeError

ONE_MINUS_EPS = 0.9999999999999999


def shear_reduced(g1, g2, s1, s2):
    """addition formula for reduced shear (shape.py:9-35)"""
    A = 1 + g1 * s1 + g2 * s2
    B = g2 * s1 - g1 * s2
    denom_inv = 1.0 / (A * A + B * B)
    g1o = (A * (g1 + s1) + B * (g2 + s2)) * denom_inv
    g2o = (A * (g2 + s2) - B * (g1 + s1)) * denom_inv
    return g1o, g2o


def g1g2_to_e1e2(g1, g2):
    """reduced shear to distortion, e = tanh(2 atanh g) (shape.py:179-235)"""
    g = np.sqrt(g1 * g1 + g2 * g2)
    if g >= 1.0:
        raise GMixRangeError("g out of bounds: %s" % g)
    if g == 0.0:
        return 0.0, 0.0
    e = np.tanh(2 * np.arctanh(g))
    if e >= 1.0:
        e = ONE_MINUS_EPS
    fac = e / g
    return fac * g1, fac * g2


def e1e2_to_g1g2(e1, e2):
    """distortion to reduced shear (shape.py:238-293)"""
    e = np.sqrt(e1 * e1 + e2 * e2)
    if e >= 1.0:
        raise GMixRangeError("e out of bounds: %s" % e)
    if e == 0.0:
        return 0.0, 0.0
    g = np.tanh(0.5 * np.arctanh(e))
    if g >= 1.0:
        g = ONE_MINUS_EPS
    fac = g / e
    return fac * e1, fac * e2


def get_round_factor(g1, g2):
    gsq = g1 ** 2 + g2 ** 2
    return (1 - gsq) / (1 + gsq)


def get_sheared_moments(irr, irc, icc, s1, s2):
    """second moments after applying reduced shear s (moments.py:268-291)"""
    T = irr + icc
    e1 = (icc - irr) / T
    e2 = 2 * irc / T
    g1, g2 = e1e2_to_g1g2(e1, e2)
    g1s, g2s = shear_reduced(g1, g2, s1, s2)
    Tround = T * get_round_factor(g1, g2)
    gsq = g1s ** 2 + g2s ** 2
    Ts = Tround * (1 + gsq) / (1 - gsq)
    e1s, e2s = g1g2_to_e1e2(g1s, g2s)
    return (1 - e1s) * Ts / 2.0, e2s * Ts / 2.0, (1 + e1s) * Ts / 2.0


class Shape(object):
    """a reduced-shear shape (g1, g2)"""

    def __init__(self, g1, g2):
        if g1 * g1 + g2 * g2 >= 1.0:
            raise GMixRangeError("g out of range")
        self.g1 = g1
        self.g2 = g2

    @property
    def g(self):
        return np.sqrt(self.g1 ** 2 + self.g2 ** 2)

    def get_sheared(self, s1, s2=None):
        if isinstance(s1, Shape):
            s1, s2 = s1.g1, s1.g2
        g1, g2 = shear_reduced(self.g1, self.g2, s1, s2)
        return Shape(g1, g2)

    def copy(self):
        return Shape(self.g1, self.g2)

    def __repr__(self):
        return "(%.16g, %.16g)" % (self.g1, self.g2)


# unit conversions (moments.py:64-92)
def fwhm_to_sigma(fwhm):
    return fwhm / 2.3548200450309493


def fwhm_to_T(fwhm):
    return 2 * fwhm_to_sigma(fwhm) ** 2


def T_to_fwhm(T):
    return 2.3548200450309493 * np.sqrt(T / 2)
