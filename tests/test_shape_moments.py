"""
Host-side shape and second-moment algebra (ngmix/shape.py, ngmix/moments.py)
against known answers from the reference (oracle/gen_golden.py --only
shape_moments).
"""
import numpy as np
import pytest

from conftest import load_golden


def _close(a, b):
    return np.allclose(np.asarray(a, dtype="f8"), b, rtol=1e-13, atol=1e-15)


def test_shape_and_moment_helpers():
    from ngmix_b200 import shape, moments
    from ngmix_b200.gexceptions import GMixRangeError
    g = load_golden("shape_moments")
    g1, g2, s1, s2, T, th = (g[k] for k in ("g1", "g2", "s1", "s2", "T",
                                            "theta"))
    assert _close(shape.g1g2_to_eta1eta2(g1, g2), g["eta_from_g"])
    e1, e2 = np.array([shape.g1g2_to_e1e2(a, b) for a, b in zip(g1, g2)]).T
    assert _close(shape.e1e2_to_eta1eta2(e1, e2), g["eta_from_e"])
    assert _close(shape.eta1eta2_to_g1g2(*g["eta_from_g"]), g["g_from_eta"])
    assert _close(g["g_from_eta"], [g1, g2])
    assert _close(shape.dgs_by_dgo_jacob(g1, g2, s1, s2), g["jacob"])
    assert _close(shape.rotate_shape(g1, g2, th), g["rot"])
    n = g1.size
    assert _close([moments.g2mom(g1[i], g2[i], T[i]) for i in range(n)],
                  g["g2mom"])
    assert _close([moments.mom2g(*m) for m in g["g2mom"]], g["mom2g"])
    assert _close([moments.mom2e(*m) for m in g["g2mom"]], g["mom2e"])
    assert _close([moments.e2mom(e1[i], e2[i], T[i]) for i in range(n)],
                  g["e2mom"])
    assert _close([moments.get_sheared_g1g2T(g1[i], g2[i], T[i], s1[i], s2[i])
                   for i in range(n)], g["sheared_g1g2T"])
    assert _close([moments.get_sheared_M1M2T(T[i] * e1[i], T[i] * e2[i], T[i],
                                             s1[i], s2[i])
                   for i in range(n)], g["sheared_M1M2T"])
    assert _close([moments.get_sheared_moments(*g["g2mom"][i], s1[i], s2[i])
                   for i in range(n)], g["sheared_moments"])
    assert _close(moments.get_Tround(T, g1, g2), g["Tround"])
    assert _close(moments.get_T(g["Tround"], g1, g2), g["T_from_round"])
    assert _close([moments.T_to_r50(T), moments.r50_to_T(T),
                   moments.r50_to_sigma(T), moments.sigma_to_r50(T)],
                  g["r50"])
    # scalars stay scalars; out-of-range shapes raise
    a, b = shape.g1g2_to_eta1eta2(0.1, 0.2)
    assert np.ndim(a) == 0 and np.ndim(b) == 0
    assert shape.g1g2_to_eta1eta2(0.0, 0.0) == (0.0, 0.0)
    with pytest.raises(GMixRangeError):
        shape.g1g2_to_eta1eta2(0.8, 0.7)
    with pytest.raises(GMixRangeError):
        moments.moms_to_e1e2(0.1, 0.1, 0.0)
    sh = shape.Shape(0.2, -0.1)
    r = sh.get_rotated(0.3)
    sh.rotate(0.3)
    assert (r.g1, r.g2) == (sh.g1, sh.g2)
    assert (-sh).g1 == -sh.g1
