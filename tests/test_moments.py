"""
Weighted moments: the oracle's restatement of
get_weighted_sums / get_higher_order_weighted_sums against golden vectors
from the reference (CPU), the CUDA pass against the same vectors and the
oracle (GPU), and the host-side statistics of make_mom_result.
"""
import os
import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _golden():
    return np.load(os.path.join(GOLDEN, "moments.npz"))


def _gaussmom_weight(fwhm):
    T = 2 * (fwhm / 2.3548200450309493) ** 2
    # gauss, T/2 on the diagonal, unit peak: flux = 2 pi sqrt(det)
    return np.array([2 * np.pi * (T / 2), 0, 0, T / 2, 0, T / 2]), T


def _rel(a, b):
    scale = np.abs(b).max()
    return np.abs(a - b).max() / (scale if scale > 0 else 1.0)


@pytest.mark.parametrize("nmom,tag", [(6, "m6"), (17, "m17")])
def test_oracle_gaussmom_sums(ora, nmom, tag):
    g = _golden()
    wt, T = _gaussmom_weight(float(g["fwhm"]))
    for i in range(g["images"].shape[0]):
        r = ora.weighted_sums(g["images"][i], g["weights"][i], g["jac"][i],
                              wt, 100 * np.sqrt(T / 2), nmom)
        area = g["jac"][i][7] ** 2
        assert r["npix"] == g[tag + "_npix"][i]
        assert _rel(r["sums"] / area, g[tag + "_sums"][i]) < 1e-14
        assert _rel(r["sums_cov"] / area ** 2, g[tag + "_sums_cov"][i]) < 1e-14
        assert abs(r["wsum"] / area / g[tag + "_wsum"][i] - 1) < 1e-14


@pytest.mark.parametrize("nmom,tag", [(6, "w6"), (17, "w17")])
def test_oracle_general_weight_sums(ora, nmom, tag):
    g = _golden()
    for i in range(g["images"].shape[0]):
        r = ora.weighted_sums(g["images"][i], g["weights"][i], g["jac"][i],
                              g["wt_pars"], float(g["maxrad"]), nmom)
        assert r["npix"] == g[tag + "_npix"][i]
        assert np.array_equal(r["sums"], g[tag + "_sums"][i])
        assert np.array_equal(r["sums_cov"], g[tag + "_sums_cov"][i])


def test_make_mom_result_matches_reference_stats():
    """host statistics from the golden sums reproduce the reference dict"""
    from ngmix_b200 import moments
    g = _golden()
    for tag, nmom in (("m6", 6), ("m17", 17)):
        for i in range(g["images"].shape[0]):
            area = g["jac"][i][7] ** 2
            # the reference applies 1/area after make_mom_result
            sums = g[tag + "_sums"][i] * area
            cov = g[tag + "_sums_cov"][i] * area ** 2
            r = moments.make_mom_result(sums.copy(), cov.copy(),
                                        g[tag + "_wsum"][i] * area)
            assert r["flags"] == g[tag + "_flags"][i]
            assert r["flux_flags"] == g[tag + "_flux_flags"][i]
            assert r["T_flags"] == g[tag + "_T_flags"][i]
            for k in ("T", "T_err", "s2n", "e1", "e2"):
                assert abs(r[k] - g[tag + "_" + k][i]) \
                    <= 1e-12 * abs(g[tag + "_" + k][i]), k
            assert abs(r["flux"] / area / g[tag + "_flux"][i] - 1) < 1e-12
            assert np.allclose(r["e_err"], g[tag + "_e_err"][i], rtol=1e-12)


def test_make_mom_result_flags():
    from ngmix_b200 import moments, flags
    sums = np.array([0.0, 0.0, 0.1, 0.05, 0.5, 1.0])
    cov = np.diag([1e-4] * 6)
    assert moments.make_mom_result(sums.copy(), cov.copy())["flags"] == 0
    bad = cov.copy()
    bad[2, 2] = -1.0
    assert moments.make_mom_result(sums.copy(), bad)["flags"] \
        == flags.NONPOS_VAR
    neg = sums.copy()
    neg[5] = -1.0
    r = moments.make_mom_result(neg, cov.copy())
    assert r["flags"] == flags.NONPOS_FLUX
    assert r["T_flags"] == flags.NONPOS_FLUX
    small = sums.copy()
    small[4] = -0.5
    assert moments.make_mom_result(small, cov.copy())["flags"] \
        == flags.NONPOS_SIZE
    with pytest.raises(ValueError):
        moments.make_mom_result(np.zeros(5), np.zeros((5, 5)))
    reg = moments.regularize_mom_shapes(
        moments.make_mom_result(sums.copy(), cov.copy()), 1.2)
    assert abs(reg["e1"] - 0.1 / (0.5 + moments.fwhm_to_T(1.2))) < 1e-15


# ------------------------------------------------------------------ GPU
@pytest.fixture(scope="module")
def ngx():
    from ngmix_b200 import _lib
    _lib.init(0)
    return _lib


@pytest.mark.gpu
@pytest.mark.parametrize("nmom,tag", [(6, "w6"), (17, "w17")])
def test_gpu_weighted_sums_golden(ngx, nmom, tag):
    import ngmix_b200 as ngmix
    g = _golden()
    n = g["images"].shape[0]
    wt = ngmix.GMix(pars=g["wt_pars"])
    gm = np.repeat(wt.get_data()[None, :], n, axis=0)
    recs = ngx.weighted_sums(gm, g["images"], g["weights"], g["jac"],
                             float(g["maxrad"]), nmom)
    assert np.array_equal(recs["npix"], g[tag + "_npix"])
    assert np.all(recs["flags"] == 0)
    for i in range(n):
        # tolerance: sums over ~300 pixels in a different (tree) order
        assert _rel(recs["sums"][i][:nmom], g[tag + "_sums"][i]) < 1e-12
        cov = recs["sums_cov"][i][:nmom * nmom].reshape(nmom, nmom)
        assert _rel(cov, g[tag + "_sums_cov"][i]) < 1e-12
        assert abs(recs["wsum"][i] / g[tag + "_wsum"][i] - 1) < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("higher,tag", [(False, "m6"), (True, "m17")])
def test_gpu_gaussmom_matches_reference(ngx, higher, tag):
    import ngmix_b200 as ngmix
    g = _golden()
    n = g["images"].shape[0]
    fitter = ngmix.GaussMom(fwhm=float(g["fwhm"]), with_higher_order=higher)
    batch = ngmix.gaussmom_batch(g["images"], g["weights"], g["jac"],
                                 float(g["fwhm"]), with_higher_order=higher)
    for i in range(n):
        jac = ngmix.Jacobian(row=g["jac"][i, 0], col=g["jac"][i, 1],
                             dvdrow=g["jac"][i, 2], dvdcol=g["jac"][i, 3],
                             dudrow=g["jac"][i, 4], dudcol=g["jac"][i, 5])
        obs = ngmix.Observation(g["images"][i], weight=g["weights"][i],
                                jacobian=jac)
        for r in (fitter.go(obs), batch[i]):
            assert r["flags"] == g[tag + "_flags"][i]
            assert r["npix"] == g[tag + "_npix"][i]
            assert _rel(r["sums"], g[tag + "_sums"][i]) < 1e-12
            assert _rel(r["sums_cov"], g[tag + "_sums_cov"][i]) < 1e-12
            for k in ("flux", "flux_err", "T", "T_err", "s2n", "e1", "e2",
                      "wsum", "sums_norm"):
                ref = g[tag + "_" + k][i]
                assert abs(r[k] - ref) <= 1e-9 * abs(ref), k
            if r["flags"] == 0:
                assert np.allclose(r["pars"][:6], g[tag + "_pars"][i],
                                   rtol=1e-9, atol=1e-12)


@pytest.mark.gpu
def test_gpu_weighted_sums_edge_cases(ngx, ora):
    import ngmix_b200 as ngmix
    from ngmix_b200 import sims
    b = sims.make_lm_batch(6, seed0=9300, dims=(21, 33), noise=0.05)
    b["weights"][1][:, :] = 0.0               # nothing to sum
    b["weights"][2][::2, :] = 0.0
    wt = ngmix.GMixModel([0.1, -0.2, 0.1, 0.05, 0.5, 1.0], "gauss")
    gm = np.repeat(wt.get_data()[None, :], 6, axis=0)
    for nmom in (6, 17):
        recs = ngx.weighted_sums(gm, b["images"], b["weights"], b["jac"],
                                 3.0, nmom)
        assert recs["npix"][1] == 0 and recs["wsum"][1] == 0.0
        for i in range(6):
            r = ora.weighted_sums(b["images"][i], b["weights"][i],
                                  b["jac"][i], wt.get_full_pars(), 3.0, nmom)
            assert recs["npix"][i] == r["npix"]
            if r["npix"] > 0:
                assert _rel(recs["sums"][i][:nmom], r["sums"]) < 1e-12
    # a degenerate weight raises like the reference's set_norms
    bad = ngmix.GMix(pars=[1.0, 0.0, 0.0, 1.0, 1.0, 1.0])
    jac = ngmix.DiagonalJacobian(row=10.0, col=16.0, scale=0.263)
    obs = ngmix.Observation(b["images"][0], weight=b["weights"][0],
                            jacobian=jac)
    with pytest.raises(ngmix.GMixRangeError):
        bad.get_weighted_sums(obs)
    # accumulation into a caller's record
    res = wt.get_weighted_sums(obs, maxrad=3.0)
    res2 = wt.get_weighted_sums(obs, maxrad=3.0, res=res.copy())
    assert np.allclose(res2["sums"], 2 * res["sums"])
    assert res2["npix"] == 2 * res["npix"]
