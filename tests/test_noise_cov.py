"""
Noise-power sandwich covariance of LM fits, Fitter(use_noise_image=True)
(ngmix/fitting/noise_cov.py).  Golden vectors: oracle/gen_golden.py --only
noise_cov (the reference's Fitter on stamps with correlated noise).  The
oracle restates the reference's Fourier-space sum with numpy; the CUDA path
evaluates the same quantity in real space (Parseval), csrc/ngx_noise.cu.
"""
import numpy as np
import pytest

from conftest import load_golden


def _rel(a, b):
    return np.max(np.abs(a - b)) / np.max(np.abs(b))


@pytest.mark.parametrize("tag,model", [("exp", "exp"), ("bdf", "bdf")])
def test_oracle_noise_cov_matches_reference(ora, tag, model):
    g = load_golden("noise_cov")
    n = g[tag + "_images"].shape[0]
    for i in range(n):
        cov, B = ora.noise_cov(model, g[tag + "_res_pars"][i],
                               g[tag + "_res_pars_cov0"][i],
                               g[tag + "_weights"][i], g[tag + "_noise"][i],
                               g[tag + "_jac"][i], psf_pars=g[tag + "_psf"][i])
        assert _rel(cov, g[tag + "_res_pars_cov"][i]) < 1e-8
        assert np.allclose(B, B.T)


@pytest.fixture(scope="module")
def ng():
    import ngmix_b200
    from ngmix_b200 import _lib
    _lib.init(0)
    return ngmix_b200


@pytest.mark.gpu
@pytest.mark.parametrize("tag,model", [("exp", "exp"), ("bdf", "bdf")])
def test_gpu_noise_b_vs_oracle(ng, ora, tag, model):
    """B from the device (real-space correlation) against the oracle's
    Fourier-space sum at the reference's solutions"""
    from ngmix_b200 import noise_cov
    g = load_golden("noise_cov")
    pars = g[tag + "_res_pars"]
    B = noise_cov.calc_noise_b(model, pars, g[tag + "_psf"], g[tag + "_jac"],
                               g[tag + "_weights"], g[tag + "_noise"])
    for i in range(pars.shape[0]):
        _, Bo = ora.noise_cov(model, pars[i], g[tag + "_res_pars_cov0"][i],
                              g[tag + "_weights"][i], g[tag + "_noise"][i],
                              g[tag + "_jac"][i], psf_pars=g[tag + "_psf"][i])
        assert _rel(B[i], Bo) < 1e-9, i


@pytest.mark.gpu
@pytest.mark.parametrize("tag,model", [("exp", "exp"), ("bdf", "bdf")])
def test_gpu_fitter_use_noise_image(ng, tag, model):
    g = load_golden("noise_cov")
    prior = None
    if model == "bdf":
        pr = ng.priors
        prior = ng.joint_prior.PriorBDFSep(
            pr.CenPrior(0.0, 0.0, 0.263, 0.263), pr.GPriorBA(0.3),
            pr.FlatPrior(-1.0, 1.0e3), pr.Normal(0.5, 0.1),
            pr.FlatPrior(-100.0, 1.0e9))
    fitter = ng.Fitter(model=model, prior=prior, use_noise_image=True)
    n = g[tag + "_images"].shape[0]
    for i in range(n):
        rec = g[tag + "_jac"][i]
        jac = ng.Jacobian(row=rec[0], col=rec[1], dvdrow=rec[2],
                          dvdcol=rec[3], dudrow=rec[4], dudcol=rec[5])
        psf_obs = ng.Observation(np.ones((5, 5)),
                                 gmix=ng.GMix(pars=g[tag + "_psf"][i].ravel()))
        obs = ng.Observation(g[tag + "_images"][i],
                             weight=g[tag + "_weights"][i],
                             noise=g[tag + "_noise"][i], jacobian=jac,
                             psf=psf_obs)
        r = fitter.go(obs, g[tag + "_guess"][i])
        assert r["flags"] == g[tag + "_res_flags"][i]
        assert r["nfev"] == g[tag + "_res_nfev"][i]
        want = g[tag + "_res_pars_err"][i]
        # the fitted point agrees to 1e-5 of the error scale, the sandwich
        # covariance follows it
        assert np.all(np.abs(r["pars_err"] / want - 1) < 1e-3), \
            (r["pars_err"], want)
        assert _rel(r["pars_cov"], g[tag + "_res_pars_cov"][i]) < 1e-3
        # the reference applies the sandwich BEFORE set_fit_result
        # (fitters.py:105-110): every derived key comes from it
        cov = g[tag + "_res_pars_cov"][i]
        assert abs(r["T_err"] / np.sqrt(cov[4, 4]) - 1) < 1e-3
        assert abs(r["flux_err"] / np.sqrt(cov[-1, -1]) - 1) < 1e-3
        assert _rel(r["g_cov"], cov[2:4, 2:4]) < 1e-3
        assert np.all(np.abs(r["g_err"] / want[2:4] - 1) < 1e-3)
        # and it is not the plain covariance
        assert np.any(np.abs(r["pars_err"] / g[tag + "_res_pars_err_plain"][i]
                             - 1) > 0.05)
    # without a noise image the reference raises ValueError
    obs2 = ng.Observation(g[tag + "_images"][0],
                          weight=g[tag + "_weights"][0], jacobian=jac,
                          psf=psf_obs)
    with pytest.raises(ValueError):
        fitter.go(obs2, g[tag + "_guess"][0])


@pytest.mark.gpu
def test_gpu_fit_batch_noise_images(ng):
    """fit_batch(noise_images=...) equals the per-object Fitter path; white
    noise at the weight level gives back (chi2-unscaled) pars_cov0"""
    g = load_golden("noise_cov")
    recs = ng.fit_batch("exp", g["exp_images"], g["exp_weights"], g["exp_jac"],
                        g["exp_psf"], g["exp_guess"],
                        noise_images=g["exp_noise"])
    for i in range(recs.size):
        assert _rel(recs["pars_cov"][i][:36].reshape(6, 6),
                    g["exp_res_pars_cov"][i]) < 1e-3
        cov = g["exp_res_pars_cov"][i]
        assert abs(recs["T_err"][i] / np.sqrt(cov[4, 4]) - 1) < 1e-3
        assert abs(recs["flux_err"][i] / np.sqrt(cov[5, 5]) - 1) < 1e-3
    # an object whose stepped parameters leave the model's domain flags only
    # itself (the reference: NaN covariance -> covariance flags)
    from ngmix_b200 import noise_cov
    B = noise_cov.calc_noise_b(
        "bdf", np.array([[0.0, 0.0, 0.1, 0.1, 0.3, 0.5, 100.0],
                         [0.0, 0.0, 0.7, 0.71409284, 0.3, 0.5, 100.0]]),
        g["bdf_psf"][:2], g["bdf_jac"][:2], g["bdf_weights"][:2],
        g["bdf_noise"][:2], per_object=True)
    assert np.all(np.isfinite(B[0])) and np.all(np.isnan(B[1]))
    # odd sizes / column padding (5 x 7) and a stamp larger than shared
    # memory (110 x 106: staged planes in global scratch) against numpy's fft
    from ngmix_b200 import _lib
    rng = np.random.RandomState(3)
    for (H, W), n, nk in (((5, 7), 2, 3), ((110, 106), 1, 2)):
        D = rng.normal(size=(n, nk, H * W))
        w = rng.uniform(0.5, 2.0, size=(n, H, W))
        nz = rng.normal(size=(n, H, W))
        B = _lib.noise_b(D, w, nz)
        for i in range(n):
            K = [np.fft.fft2(w[i] * D[i, a].reshape(H, W)) for a in range(nk)]
            P = np.abs(np.fft.fft2(nz[i])) ** 2
            for a in range(nk):
                for b in range(nk):
                    want = np.sum(np.conj(K[a]) * K[b] * P).real / (H * W) ** 2
                    assert abs(B[i, a, b] - want) \
                        < 1e-9 * abs(B[i, a, a] * B[i, b, b]) ** 0.5 + 1e-12
