"""
Moment bookkeeping on the host: the name map of the 6/17 weighted moments,
size conversions, and the summary statistics derived from raw moment sums.
Mirrors ngmix/moments.py (MOMENTS_NAME_MAP :8-61, conversions :64-131,
make_mom_result :398-540, regularize_mom_shapes :578-650).  The sums
themselves come from the device (ngx_weighted_sums).
"""
import numpy as np

from . import flags as _flags

_FWHM_PER_SIGMA = 2.3548200450309493

# index of every named moment in sums / sums_cov; the second-order ones also
# go by the piff-style Mpq names
MOMENTS_NAME_MAP = {
    "Mv": 0, "Mu": 1, "M1": 2, "M2": 3, "MT": 4, "MF": 5,
    "M00": 5, "M10": 1, "M01": 0, "M11": 4, "M20": 2, "M02": 3,
    "M21": 6, "M12": 7, "M30": 8, "M03": 9,
    "M22": 10, "M31": 11, "M13": 12, "M40": 13, "M14": 14,
    "M33": 15, "M44": 16,
}


def sigma_to_fwhm(sigma):
    return sigma * _FWHM_PER_SIGMA


def fwhm_to_sigma(fwhm):
    return fwhm / _FWHM_PER_SIGMA


def T_to_fwhm(T):
    return sigma_to_fwhm(np.sqrt(T / 2.0))


def fwhm_to_T(fwhm):
    sigma = fwhm_to_sigma(fwhm)
    return 2 * sigma ** 2


def get_ratio_var(a, b, var_a, var_b, cov_ab):
    """variance of a/b (ngmix/util.py:57-68)"""
    if np.any(b == 0):
        raise ValueError("zero in denominator")
    rsq = (a / b) ** 2
    with np.errstate(divide="ignore", invalid="ignore"):
        return rsq * (var_a / a ** 2 + var_b / b ** 2 - 2 * cov_ab / (a * b))


def get_ratio_error(a, b, var_a, var_b, cov_ab):
    """error on a/b, clipped at zero variance (ngmix/util.py:71-80)"""
    var = np.clip(get_ratio_var(a, b, var_a, var_b, cov_ab), 0.0, np.inf)
    return np.sqrt(var)


def make_mom_result(sums, sums_cov, sums_norm=None):
    """
    Result dict from unnormalised moment sums [Mv, Mu, M1, M2, MT, MF, ...]
    (6 or 17 of them) and their covariance: flux, T, e1/e2 with errors and the
    flag words, exactly as ngmix.moments.make_mom_result builds them.
    """
    if len(sums) not in (6, 17):
        raise ValueError("You must pass exactly 6 or 17 unnormalized moments "
                         "in the order [Mv, Mu, M1, M2, MT, MF, ...]")
    if sums_cov.shape not in ((6, 6), (17, 17)):
        raise ValueError("You must pass a 6x6 or 17x17 matrix")
    iv, iu = MOMENTS_NAME_MAP["Mv"], MOMENTS_NAME_MAP["Mu"]
    i1, i2 = MOMENTS_NAME_MAP["M1"], MOMENTS_NAME_MAP["M2"]
    it, if_ = MOMENTS_NAME_MAP["MT"], MOMENTS_NAME_MAP["MF"]
    nan = np.nan
    res = {
        "flags": 0, "flagstr": "", "flux": sums[if_], "sums": sums,
        "sums_cov": sums_cov,
        "sums_norm": sums_norm if sums_norm is not None else nan,
        "flux_flags": 0, "flux_flagstr": "", "T_flags": 0, "T_flagstr": "",
        "flux_err": nan, "T": nan, "T_err": nan, "s2n": nan, "e1": nan,
        "e2": nan, "e": np.array([nan, nan]), "e_err": np.array([nan, nan]),
        "e_cov": np.diag([nan, nan]), "sums_err": np.full(6, nan),
    }
    fvar, tvar = sums_cov[if_, if_], sums_cov[it, it]
    if fvar > 0:
        res["flux_err"] = np.sqrt(fvar)
        res["s2n"] = res["flux"] / res["flux_err"]
    else:
        res["flux_flags"] |= _flags.NONPOS_VAR

    if fvar > 0 and tvar > 0:
        if sums[if_] > 0:
            res["T"] = sums[it] / sums[if_]
            res["T_err"] = get_ratio_error(sums[it], sums[if_], tvar, fvar,
                                           sums_cov[it, if_])
        else:
            res["T_flags"] |= _flags.NONPOS_FLUX
    else:
        res["T_flags"] |= _flags.NONPOS_VAR

    diag = np.diagonal(sums_cov)
    if np.all(diag > 0):
        res["sums_err"] = np.sqrt(diag)
    else:
        res["flags"] |= _flags.NONPOS_VAR

    if res["flags"] == 0:
        if res["flux"] > 0:
            if res["T"] > 0:
                res["e1"] = sums[i1] / sums[it]
                res["e2"] = sums[i2] / sums[it]
                res["e"] = np.array([res["e1"], res["e2"]])
                res["pars"] = np.array([sums[iv], sums[iu], res["e1"],
                                        res["e2"], res["T"], res["flux"]])
                e_err = np.array([
                    get_ratio_error(sums[i1], sums[it], sums_cov[i1, i1],
                                    tvar, sums_cov[i1, it]),
                    get_ratio_error(sums[i2], sums[it], sums_cov[i2, i2],
                                    tvar, sums_cov[i2, it]),
                ])
                if np.all(np.isfinite(e_err)):
                    res["e_err"] = e_err
                    res["e_cov"] = np.diag(e_err ** 2)
                else:
                    res["flags"] |= _flags.NONPOS_SHAPE_VAR
            else:
                res["flags"] |= _flags.NONPOS_SIZE
        else:
            res["flags"] |= _flags.NONPOS_FLUX

    res["flagstr"] = _flags.get_flags_str(res["flags"])
    res["flux_flagstr"] = _flags.get_flags_str(res["flux_flags"])
    res["T_flagstr"] = _flags.get_flags_str(res["T_flags"])
    _add_moments_by_name(res)
    return res


def _add_moments_by_name(res):
    """named moments normalised by the flux sum, with ratio errors"""
    sums, cov = res["sums"], res["sums_cov"]
    if_ = MOMENTS_NAME_MAP["MF"]
    fsum = sums[if_]
    with np.errstate(invalid="ignore"):
        fsum_err = np.sqrt(cov[if_, if_])
    for name, ind in MOMENTS_NAME_MAP.items():
        if ind > sums.size - 1:
            continue
        if name in ("MF", "M00"):
            res[name], res[name + "_err"] = fsum, fsum_err
        elif fsum > 0:
            res[name] = sums[ind] / fsum
            res[name + "_err"] = get_ratio_error(
                sums[ind], fsum, cov[ind, ind], cov[if_, if_], cov[ind, if_])
        else:
            res[name] = res[name + "_err"] = np.nan


def regularize_mom_shapes(res, fwhm_reg):
    """e_{1,2} = M_{1,2} / (T + T_reg) with T_reg from fwhm_reg; size and flux
    are kept (ngmix/moments.py:578-650)"""
    if not fwhm_reg > 0:
        return res
    raw, raw_cov = res["sums"], res["sums_cov"]
    amat = np.eye(6)
    amat[4, 5] = fwhm_to_T(fwhm_reg)
    nan_cen = [np.isnan(raw[0]), np.isnan(raw[1])]
    work = raw.copy()
    for i in (0, 1):
        if nan_cen[i]:
            work[i] = 0
    reg = np.dot(amat, work)
    for i in (0, 1):
        if nan_cen[i]:
            reg[i] = np.nan
    out = make_mom_result(reg, np.dot(amat, np.dot(raw_cov, amat.T)))
    for col in ("T", "T_err", "T_flags", "T_flagstr"):
        out[col] = res[col]
    out["flags"] |= res["flags"]
    out["flagstr"] = _flags.get_flags_str(out["flags"])
    return out


# ---- gaussian size conversions and second-moment algebra
# (ngmix/moments.py:94-395); irr = <y^2>, irc = <xy>, icc = <x^2>
def r50_to_sigma(r50):
    """half-light radius of a gaussian = fwhm / 2"""
    return fwhm_to_sigma(2.0 * r50)


def sigma_to_r50(sigma):
    return sigma_to_fwhm(sigma) / 2.0


def r50_to_T(r50):
    return 2 * r50_to_sigma(r50) ** 2


def T_to_r50(T):
    return sigma_to_r50(np.sqrt(T / 2.0))


def moms_to_e1e2(M1, M2, T):
    """(M1, M2) / T; T must be positive"""
    from .gexceptions import GMixRangeError
    if np.any(np.asarray(T) <= 0.0):
        raise GMixRangeError("T <= 0.0")
    Tinv = 1.0 / T
    return M1 * Tinv, M2 * Tinv


def get_Tround(T, g1, g2):
    """size of the round object the shape (g1, g2) was sheared from"""
    gsq = g1 ** 2 + g2 ** 2
    return T * (1 - gsq) / (1 + gsq)


def get_T(Tround, g1, g2):
    gsq = g1 ** 2 + g2 ** 2
    return Tround * (1 + gsq) / (1 - gsq)


def mom2e(Irr, Irc, Icc):
    T = Irr + Icc
    return (Icc - Irr) / T, 2.0 * Irc / T, T


def mom2g(Irr, Irc, Icc):
    from . import shape
    e1, e2, T = mom2e(Irr, Irc, Icc)
    g1, g2 = shape.e1e2_to_g1g2(e1, e2)
    return g1, g2, T


def e2mom(e1, e2, T):
    return (1 - e1) * T / 2.0, e2 * T / 2.0, (1 + e1) * T / 2.0


def g2mom(g1, g2, T):
    from . import shape
    return e2mom(*shape.g1g2_to_e1e2(g1, g2), T)


def get_sheared_g1g2T(g1, g2, T, s1, s2):
    """shape and size after the shear (s1, s2): the round size is kept"""
    from . import shape
    g1s, g2s = shape.shear_reduced(g1, g2, s1, s2)
    return g1s, g2s, get_T(get_Tround(T, g1, g2), g1s, g2s)


def get_sheared_M1M2T(M1, M2, T, s1, s2):
    from . import shape
    g1, g2 = shape.e1e2_to_g1g2(*moms_to_e1e2(M1, M2, T))
    g1s, g2s, Ts = get_sheared_g1g2T(g1, g2, T, s1, s2)
    e1s, e2s = shape.g1g2_to_e1e2(g1s, g2s)
    return Ts * e1s, Ts * e2s, Ts


def get_sheared_moments(irr, irc, icc, s1, s2):
    g1, g2, T = mom2g(irr, irc, icc)
    return g2mom(*get_sheared_g1g2T(g1, g2, T, s1, s2))
