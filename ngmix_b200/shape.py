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

    def set_g1g2(self, g1, g2):
        if g1 * g1 + g2 * g2 >= 1.0:
            raise GMixRangeError("g out of range")
        self.g1, self.g2 = g1, g2

    def get_rotated(self, theta_radians):
        return Shape(*rotate_shape(self.g1, self.g2, theta_radians))

    def rotate(self, theta_radians):
        self.set_g1g2(*rotate_shape(self.g1, self.g2, theta_radians))

    def __neg__(self):
        return Shape(-self.g1, -self.g2)

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


def rotate_shape(g1, g2, theta):
    """the shape seen after rotating the axes by theta radians: a spin-2
    rotation by 2 theta (shape.py:492-519)"""
    c, sn = np.cos(2.0 * theta), np.sin(2.0 * theta)
    return g1 * c + g2 * sn, g2 * c - g1 * sn


def _scaled(a1, a2, amax_msg, scale_of):
    """(a1, a2) rescaled so its modulus becomes scale_of(modulus); the
    modulus must be < 1; zeros stay zeros.  Scalars in, scalars out."""
    scalar = np.ndim(a1) == 0
    a1 = np.atleast_1d(np.asarray(a1, dtype="f8"))
    a2 = np.atleast_1d(np.asarray(a2, dtype="f8"))
    mod = np.sqrt(a1 * a1 + a2 * a2)
    if np.any(mod >= 1.0):
        raise GMixRangeError(amax_msg)
    fac = np.zeros(mod.size)
    w = mod > 0.0
    fac[w] = scale_of(mod[w]) / mod[w]
    o1, o2 = fac * a1, fac * a2
    return (o1[0], o2[0]) if scalar else (o1, o2)


def g1g2_to_eta1eta2(g1, g2):
    """eta = 2 atanh(g) along the same position angle (shape.py:296-347)"""
    return _scaled(g1, g2, "g out of bounds converting to eta",
                   lambda g: 2 * np.arctanh(g))


def e1e2_to_eta1eta2(e1, e2):
    """eta = atanh(e) (shape.py:350-393)"""
    return _scaled(e1, e2, "some e were out of bounds", np.arctanh)


def eta1eta2_to_g1g2(eta1, eta2):
    """g = tanh(eta / 2) (shape.py:396-440)"""
    scalar = np.ndim(eta1) == 0
    eta1 = np.atleast_1d(np.asarray(eta1, dtype="f8"))
    eta2 = np.atleast_1d(np.asarray(eta2, dtype="f8"))
    eta = np.sqrt(eta1 * eta1 + eta2 * eta2)
    g = np.tanh(0.5 * eta)
    if np.any(g >= 1.0):
        raise GMixRangeError("some g were out of bounds")
    fac = np.zeros(eta.size)
    w = eta != 0.0
    fac[w] = g[w] / eta[w]
    g1, g2 = fac * eta1, fac * eta2
    return (g1[0], g2[0]) if scalar else (g1, g2)


def dgs_by_dgo_jacob(g1, g2, s1, s2):
    """jacobian determinant of the map g -> g sheared by s
    (shape.py:443-468)"""
    ssq = s1 * s1 + s2 * s2
    den = 1 + 2 * g1 * s1 + 2 * g2 * s2 + g1 ** 2 * ssq + g2 ** 2 * ssq
    return (ssq - 1) ** 2 / den ** 2
