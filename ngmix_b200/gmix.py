"""
GMix / GMixModel / GMixCM / GMixCoellip: host-side mirror of
ngmix/gmix/gmix.py.  The record array (`get_data()`, _gauss2d_dtype) lives on
the host as in the reference; every operation the reference hands to numba
(fill, set_norms, convolve, render, fill_fdiff, get_loglike,
get_model_s2n_sum) goes through the C ABI to the CUDA kernels with n = 1.
The batched entry points in ngmix_b200.batch are the fast path.
"""
import numpy as np

from . import _lib
from .gexceptions import GMixRangeError
from .jacobian import Jacobian, UnitJacobian

__all__ = [
    'GMix', 'GMixModel', 'GMixCM', 'GMixCoellip', 'gmix_concat',
    'make_gmix_model', 'get_coellip_npars', 'get_coellip_ngauss',
    'get_model_num', 'get_model_name', 'get_model_ngauss', 'get_model_npars',
]

_gauss2d_dtype = _lib.GAUSS_DTYPE

GMIX_FULL = 0
GMIX_GAUSS = 1
GMIX_TURB = 2
GMIX_EXP = 3
GMIX_DEV = 4
GMIX_BDC = 5
GMIX_BDF = 6
GMIX_COELLIP = 7
GMIX_CM = 9
GMIX_BD = 10

_names = {GMIX_FULL: "full", GMIX_GAUSS: "gauss", GMIX_TURB: "turb",
          GMIX_EXP: "exp", GMIX_DEV: "dev", GMIX_BDC: "bdc", GMIX_BD: "bd",
          GMIX_BDF: "bdf", GMIX_CM: "cm", GMIX_COELLIP: "coellip"}
_gmix_model_dict = {}
_gmix_string_dict = {}
for _num, _name in _names.items():
    _gmix_model_dict[_num] = _num
    _gmix_model_dict[_name] = _num
    _gmix_string_dict[_num] = _name
    _gmix_string_dict[_name] = _name

_gmix_npars_dict = {GMIX_GAUSS: 6, GMIX_TURB: 6, GMIX_EXP: 6, GMIX_DEV: 6,
                    GMIX_CM: 6, GMIX_BD: 8, GMIX_BDF: 7, GMIX_BDC: 8}
_gmix_ngauss_dict = {GMIX_GAUSS: 1, "gauss": 1, GMIX_TURB: 3, "turb": 3,
                     GMIX_EXP: 6, "exp": 6, GMIX_DEV: 10, "dev": 10,
                     GMIX_CM: 16, GMIX_BD: 16, GMIX_BDF: 16, GMIX_BDC: 16}
for _k in range(1, 6):
    _gmix_ngauss_dict["em%d" % _k] = _k
    _gmix_ngauss_dict["coellip%d" % _k] = _k


def get_model_num(model):
    if model not in _gmix_model_dict:
        raise ValueError("unknown model: '%s'" % model)
    return _gmix_model_dict[model]


def get_model_name(model):
    if model not in _gmix_string_dict:
        raise ValueError("unknown model: '%s'" % model)
    return _gmix_string_dict[model]


def get_model_ngauss(model):
    if model not in _gmix_ngauss_dict:
        raise ValueError("unknown model: '%s'" % model)
    return _gmix_ngauss_dict[model]


def get_model_npars(model):
    if model not in _gmix_model_dict:
        raise ValueError("bad model: '%s'" % model)
    return _gmix_npars_dict[_gmix_model_dict[model]]


def get_coellip_npars(ngauss):
    return 4 + 2 * ngauss


def get_coellip_ngauss(npars):
    return (npars - 4) // 2


def make_gmix_model(pars, model):
    """a GMix of the right class for the model (gmix.py:37-60)"""
    model = get_model_num(model)
    if model == GMIX_COELLIP:
        return GMixCoellip(pars)
    elif model == GMIX_FULL:
        return GMix(pars=pars)
    return GMixModel(pars, model)


def _raise_on_status(status, what):
    if status == _lib.ST_RANGE_ERROR:
        raise GMixRangeError(what)
    if status == _lib.ST_ZERO_DIV:
        raise ZeroDivisionError(what)


class GMix(object):
    """
    A two-dimensional gaussian mixture.

    Parameters
    ----------
    ngauss: int, optional
        number of gaussians; the mixture is left zeroed
    pars: array-like, optional
        6*ngauss values [p1,row1,col1,irr1,irc1,icc1, p2,...]
    """

    def __init__(self, ngauss=None, pars=None):
        self._model = GMIX_FULL
        self._model_name = "full"
        if ngauss is None and pars is None:
            raise ValueError("send ngauss= or pars=")
        if pars is not None:
            npars = len(pars)
            if (npars % 6) != 0:
                raise ValueError(
                    "len(pars) must be mutiple of 6 got %s" % npars)
            self._ngauss = npars // 6
            self._npars = npars
            self.reset()
            self._fill(pars)
        else:
            self._ngauss = ngauss
            self._npars = 6 * ngauss
            self.reset()

    # ---- data access -------------------------------------------------
    def get_data(self):
        return self._data

    def get_full_pars(self):
        gm = self._data
        pars = np.zeros(self._ngauss * 6)
        for k, n in enumerate(("p", "row", "col", "irr", "irc", "icc")):
            pars[k::6] = gm[n]
        return pars

    def get_cen(self):
        gm = self._data
        psum = gm["p"].sum()
        return ((gm["row"] * gm["p"]).sum() / psum,
                (gm["col"] * gm["p"]).sum() / psum)

    def set_cen(self, row, col):
        gm = self._data
        row0, col0 = self.get_cen()
        gm["row"] += row - row0
        gm["col"] += col - col0

    def get_e1e2T(self):
        gm = self._data
        row, col = self.get_cen()
        rowdiff = gm["row"] - row
        coldiff = gm["col"] - col
        p = gm["p"]
        ipsum = 1.0 / p.sum()
        irr = ((gm["irr"] + rowdiff ** 2) * p).sum() * ipsum
        irc = ((gm["irc"] + rowdiff * coldiff) * p).sum() * ipsum
        icc = ((gm["icc"] + coldiff ** 2) * p).sum() * ipsum
        T = irr + icc
        return (icc - irr) / T, 2.0 * irc / T, T

    def get_T(self):
        return self.get_e1e2T()[2]

    def get_sigma(self):
        return np.sqrt(self.get_T() / 2.0)

    def get_g1g2T(self):
        from .shape import e1e2_to_g1g2
        e1, e2, T = self.get_e1e2T()
        g1, g2 = e1e2_to_g1g2(e1, e2)
        return g1, g2, T

    def get_e1e2sigma(self):
        e1, e2, T = self.get_e1e2T()
        return e1, e2, np.sqrt(T / 2)

    def get_g1g2sigma(self):
        g1, g2, T = self.get_g1g2T()
        return g1, g2, np.sqrt(T / 2)

    def get_flux(self):
        return self._data["p"].sum()

    get_psum = get_flux

    def get_gaussap_flux(self, fwhm=None, sigma=None, T=None):
        """flux of the mixture under a round gaussian aperture weight
        (gmix/gmix.py:325-391): every component contributes p times
        sqrt(det((S^-1 + W^-1)^-1) / det S), capped at 1, with W the weight's
        covariance; components with det <= 1e-200 count in full"""
        from . import moments
        if fwhm is not None:
            sigma = moments.fwhm_to_sigma(fwhm)
        elif T is not None:
            sigma = np.sqrt(T / 2.0)
        elif sigma is not None:
            sigma = float(sigma)
        else:
            raise ValueError("send weight function sigma, fwhm, or T")
        gm = self.get_data()
        winv = 1.0 / sigma ** 2
        det = gm["det"]
        fac = np.ones(gm.size)
        ok = det > 1.0e-200
        with np.errstate(all="ignore"):
            # S^-1 + W^-1 for S = [[irr, irc], [irc, icc]]
            a = gm["icc"][ok] / det[ok] + winv
            b = -gm["irc"][ok] / det[ok]
            c = gm["irr"][ok] / det[ok] + winv
            f = np.sqrt(1.0 / ((a * c - b * b) * det[ok]))
        fac[ok] = np.where(np.isfinite(f), np.minimum(f, 1.0), 1.0)
        return float((gm["p"] * fac).sum())

    def set_flux(self, psum):
        gm = self._data
        gm["p"] *= psum / gm["p"].sum()
        gm["norm_set"] = 0

    set_psum = set_flux

    def scale_T(self, scale):
        if scale < 0.0:
            raise ValueError(f'Requested scale {scale} < 0')
        gm = self._data.copy()
        row0, col0 = self.get_cen()
        gm['row'] = (gm['row'] - row0) * np.sqrt(scale) + row0
        gm['col'] = (gm['col'] - col0) * np.sqrt(scale) + col0
        for n in ('irr', 'irc', 'icc'):
            gm[n] *= scale
        gm["norm_set"] = 0
        self._data = gm

    # ---- operations that run on the device ---------------------------
    def set_norms(self):
        """gmix_set_norms (gmix_nb.py:176-218) through ngx_gmix_set_norms"""
        gm = self._data.reshape(1, -1)
        st = _lib.gmix_set_norms(gm)
        self._data = gm[0]
        _raise_on_status(st[0], "det or T too low")

    def set_norms_if_needed(self):
        if self._data["norm_set"][0] == 0:
            self.set_norms()

    def fill(self, pars):
        npars = len(pars)
        if npars != self._npars:
            raise ValueError("model '%s' requires %s pars, got %s"
                             % (self._model_name, self._npars, npars))
        self._fill(pars)

    def _fill(self, pars):
        """gmix_fill_<model> (gmix_nb.py:307-558) through ngx_gmix_fill"""
        self._pars[:] = pars
        gm, st = _lib.gmix_fill(self._model_name, self._pars[None, :])
        _raise_on_status(st[0], "g >= 1")
        self._data = gm[0]

    def convolve(self, psf):
        """gmix_convolve_fill (gmix_nb.py:609-649) via ngx_gmix_convolve"""
        if not isinstance(psf, GMix):
            raise TypeError("Can only convolve with another GMix  got type %s"
                            % type(psf))
        out = GMix(ngauss=len(self) * len(psf))
        out._data = _lib.gmix_convolve(self._data[None], psf._data[None])[0]
        return out

    def make_image(self, dims, jacobian=None, fast_exp=False):
        """render (render_nb.py:9-36) through ngx_render"""
        dims = np.array(dims, ndmin=1, dtype="i8")
        if dims.size != 2:
            raise ValueError("images must have two dimensions, got %s"
                             % str(dims))
        image = np.zeros(dims, dtype="f8")
        self._fill_image(image, jacobian=jacobian, fast_exp=fast_exp)
        return image

    def _fill_image(self, image, jacobian=None, fast_exp=False):
        if jacobian is None:
            cen = (np.array(image.shape) - 1.0) / 2.0
            jacobian = UnitJacobian(row=cen[0], col=cen[1])
        else:
            assert isinstance(jacobian, Jacobian)
        out, st = _lib.render(self._data[None], jacobian._data, image.shape,
                              fast_exp=fast_exp, images=image[None].copy())
        _raise_on_status(st[0], "det or T too low")
        # render sets the norms when they were not set
        if self._data["norm_set"][0] == 0:
            self.set_norms()
        image[:, :] = out[0]

    def fill_fdiff(self, obs, fdiff, start=0):
        """fill_fdiff (gmix_nb.py:877-900) through ngx_loglike_pixels"""
        nuse = fdiff.size - start
        if nuse < obs.image.size:
            raise ValueError("fdiff from start must have len >= %d, got %d"
                             % (obs.image.size, nuse))
        self.set_norms_if_needed()
        pix = obs.pixels
        _, fd, st = _lib.loglike_pixels(self._data[None], pix[None],
                                        want_fdiff=True)
        _raise_on_status(st[0], "det or T too low")
        fdiff[start:start + pix.size] = fd[0]

    def get_loglike(self, obs, more=False):
        """get_loglike (gmix_nb.py:824-874) through ngx_loglike_pixels"""
        self.set_norms_if_needed()
        out4, _, st = _lib.loglike_pixels(self._data[None], obs.pixels[None])
        _raise_on_status(st[0], "det or T too low")
        if more:
            return {"loglike": out4[0, 0], "s2n_numer": out4[0, 1],
                    "s2n_denom": out4[0, 2], "npix": int(out4[0, 3])}
        return out4[0, 0]

    def get_model_s2n_sum(self, obs):
        """get_model_s2n_sum (gmix_nb.py:903-937)"""
        self.set_norms_if_needed()
        out4, _, st = _lib.loglike_pixels(self._data[None], obs.pixels[None])
        _raise_on_status(st[0], "det or T too low")
        return out4[0, 2]

    def get_model_s2n(self, obs):
        return np.sqrt(self.get_model_s2n_sum(obs))

    def get_weighted_sums(self, obs, maxrad=None, with_higher_order=False,
                          res=None):
        """
        Weighted moment sums with this mixture as the weight
        (gmix/gmix.py:706-754 -> gmix_nb.py:681-821) through
        ngx_weighted_sums.  Returns a record with flags, npix, wsum, sums,
        sums_cov, pars, F; when `res` is sent the sums are added to it.
        """
        self.set_norms_if_needed()
        if maxrad is None:
            maxrad = 100 * np.sqrt(self.get_T() / 2)
        nmom = 17 if with_higher_order else 6
        rec = _lib.weighted_sums(
            self._data[None], obs.image, obs.weight, obs.jacobian._data,
            maxrad, nmom,
            ignore_zero_weight=getattr(obs, "_ignore_zero_weight", True))[0]
        if rec["flags"] & 128:      # NGX_GMIX_RANGE_ERROR
            raise GMixRangeError("det or T too low")
        new = moments_record(rec, nmom)
        if res is None:
            return new
        for name in ("npix", "wsum", "sums", "sums_cov"):
            res[name] += new[name]
        return res

    def get_weighted_moments(self, obs, maxrad=None, with_higher_order=False):
        """weighted moments with e1, e2, T, s2n etc.
        (gmix/gmix.py:674-704)"""
        res = self.get_weighted_sums(obs, maxrad=maxrad,
                                     with_higher_order=with_higher_order)
        return get_weighted_moments_stats(res)

    # ---- pure host helpers -------------------------------------------
    def get_sheared(self, s1, s2=None):
        from .shape import Shape, get_sheared_moments
        if isinstance(s1, Shape):
            shear1, shear2 = s1.g1, s1.g2
        elif s2 is not None:
            shear1, shear2 = s1, s2
        else:
            raise ValueError("send a Shape or s1,s2")
        new_gmix = self.copy()
        nd = new_gmix._data
        nd["norm_set"] = 0
        for i in range(len(self)):
            irr, irc, icc = get_sheared_moments(
                nd["irr"][i], nd["irc"][i], nd["icc"][i], shear1, shear2)
            nd["irr"][i], nd["irc"][i], nd["icc"][i] = irr, irc, icc
            nd["det"][i] = irr * icc - irc * irc
        return new_gmix

    def make_round(self, preserve_size=False):
        from .shape import get_round_factor
        gm = self.copy()
        if preserve_size:
            e1, e2, T = gm.get_e1e2T()
            mat = 0.5 * T * np.array([[1 - e1, e2], [e2, 1 + e1]])
            factor = np.linalg.eigvals(mat).max() / (T / 2.0)
        else:
            g1, g2, T = gm.get_g1g2T()
            factor = get_round_factor(g1, g2)
        d = gm._data
        d["norm_set"] = 0
        Ti = d["irr"] + d["icc"]
        d["irc"] = 0.0
        d["irr"] = 0.5 * Ti * factor
        d["icc"] = 0.5 * Ti * factor
        d["det"] = d["irr"] * d["icc"]
        return gm

    def reset(self):
        self._pars = np.zeros(self._npars)
        self._data = np.zeros(self._ngauss, dtype=_gauss2d_dtype)

    def copy(self):
        gmix = GMix(ngauss=self._ngauss)
        gmix._data[:] = self._data[:]
        return gmix

    def __copy__(self):
        return self.copy()

    def __deepcopy__(self, memo):
        result = self.copy()
        memo[id(self)] = result
        return result

    def __eq__(self, gm):
        if not isinstance(gm, GMix):
            raise ValueError(f'expected GMix, got {type(gm)}')
        a, b = self._data, gm._data
        return all(np.all(a[n] == b[n])
                   for n in ('p', 'row', 'col', 'irr', 'irc', 'icc', 'det'))

    def __len__(self):
        return self._ngauss

    def __repr__(self):
        fmt = "p: %.4g row: %.4g col: %.4g irr: %.4g irc: %.4g icc: %.4g"
        return "\n".join(
            fmt % (t["p"], t["row"], t["col"], t["irr"], t["irc"], t["icc"])
            for t in self._data)


class GMixModel(GMix):
    """a mixture for one of the named models: gauss, turb, exp, dev, bdf, bd"""

    def __init__(self, pars, model):
        self._model = get_model_num(model)
        self._model_name = get_model_name(self._model)
        self._ngauss = get_model_ngauss(self._model)
        self._npars = get_model_npars(self._model)
        self.reset()
        self.fill(pars)

    def copy(self):
        return GMixModel(self._pars, self._model_name)

    def set_cen(self, row, col):
        super().set_cen(row, col)
        self._pars[0] = row
        self._pars[1] = col


class GMixCM(GMixModel):
    """composite model with fixed fracdev and TdByTe (gmix.py:985-1039)"""

    def __init__(self, fracdev, TdByTe, pars):
        self._fracdev = fracdev
        self._TdByTe = TdByTe
        self._Tfactor = _cm_Tfactor(fracdev, TdByTe)
        super().__init__(pars, "cm")

    def copy(self):
        return GMixCM(self._fracdev, self._TdByTe, self._pars)

    def _fill(self, pars):
        self._pars[:] = pars
        cm = np.array([[self._fracdev, self._TdByTe, self._Tfactor]])
        gm, st = _lib.gmix_fill_cm(cm, self._pars[None, :])
        _raise_on_status(st[0], "g >= 1")
        self._data = gm[0]


def _cm_Tfactor(fracdev, TdByTe):
    """get_cm_Tfactor (gmix_nb.py:561-593); a 16-term host scalar sum"""
    from .sims import _TABLES
    fe, pe = _TABLES["exp"]
    fd, pd = _TABLES["dev"]
    ifracdev = 1.0 - fracdev
    Tfactor = 0.0
    for i in range(6):
        Tfactor += (pe[i] * ifracdev) * fe[i]
    for i in range(10):
        Tfactor += (pd[i] * fracdev) * (fd[i] * TdByTe)
    return 1.0 / Tfactor


class GMixCoellip(GMixModel):
    """coelliptical gaussians, pars [cen1,cen2,g1,g2,T1..Tn,F1..Fn]"""

    def __init__(self, pars):
        self._model = GMIX_COELLIP
        self._model_name = "coellip"
        npars = len(pars)
        if ((npars - 4) % 2) != 0:
            raise ValueError("coellip must have len(pars)==4+2*ngauss, got %s"
                             % npars)
        self._ngauss = (npars - 4) // 2
        self._npars = npars
        self.reset()
        self._fill(pars)

    def copy(self):
        return GMixCoellip(self._pars)


def get_moments_result_dtype(with_higher_order=False):
    """gmix/gmix.py:1314-1330"""
    nmom = 17 if with_higher_order else 6
    return [("flags", "i4"), ("npix", "i4"), ("wsum", "f8"),
            ("sums", "f8", nmom), ("sums_cov", "f8", (nmom, nmom)),
            ("pars", "f8", nmom), ("F", "f8", nmom)]


def moments_record(rec, nmom):
    """one ngx_mom_result -> a record of the reference's moments dtype"""
    out = np.zeros(1, dtype=np.dtype(
        get_moments_result_dtype(with_higher_order=(nmom == 17)),
        align=True))[0]
    out["flags"] = 0
    out["npix"] = rec["npix"]
    out["wsum"] = rec["wsum"]
    out["sums"] = rec["sums"][:nmom]
    out["sums_cov"] = rec["sums_cov"][:nmom * nmom].reshape(nmom, nmom)
    return out


def get_weighted_moments_stats(ares):
    """summary statistics of a moments record (gmix/gmix.py:1295-1311)"""
    from . import moments
    res = {}
    for name in ares.dtype.names:
        if name in ("sums", "sums_cov"):
            res[name] = ares[name].copy()
        else:
            res[name] = ares[name]
    res.update(moments.make_mom_result(res["sums"].copy(),
                                       res["sums_cov"].copy(), res["wsum"]))
    return res


def gmix_concat(gmixes):
    if len(gmixes) == 0:
        raise ValueError('send at least one gmix')
    pars = []
    for g in gmixes:
        pars += list(g.get_full_pars())
    return GMix(pars=pars)
