"""
GPU parity on the remaining BASELINE.json configs (at sizes the CPU oracle
finishes in seconds) and on edge cases: config 4 (bdf, 64x64, 48 composed
gaussians, PriorBDFSep, forward-difference Jacobian), config 5 (metacal-style
5 sheared stamps per object), empty batches, tiny and non-square stamps,
no-psf fits, masked stamps, size-independent properties at full batch size.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

PAR_TOL = 1e-5


@pytest.fixture(scope="module")
def ngx():
    from ngmix_b200 import _lib
    _lib.init(0)
    return _lib


def _bdf_prior(ngx):
    pr = ngx.Prior()
    pr.present = 1
    pr.g_kind = 1
    pr.cen1, pr.cen2, pr.sigma1, pr.sigma2 = 0.0, 0.0, 0.263, 0.263
    pr.g_sigma = 0.3
    pr.T = ngx.Prior1D(1, 0, -1.0, 1.0e3, 0, 0)
    pr.fracdev = ngx.Prior1D(3, 0, 0.5, 0.1, 0, 0)
    pr.F = ngx.Prior1D(1, 0, -100.0, 1.0e9, 0, 0)
    return pr


def _close(got, ref, err, np_, frac=1e-3):
    tol = np.maximum(PAR_TOL * np.abs(ref[:np_]), frac * err[:np_])
    return np.all(np.abs(got[:np_] - ref[:np_]) <= tol)


def test_config4_bdf_64(ngx, ora):
    from ngmix_b200 import sims
    b = sims.make_bdf_batch(6, seed0=4100, dims=(64, 64), noise=0.02)
    res = ngx.lm_fit("bdf", b["images"], b["weights"], b["jac"], b["psf"],
                     b["guess"], prior=_bdf_prior(ngx))
    ref = ora.lm_fit_batch("bdf", b["images"], b["weights"], b["jac"],
                           b["psf"], b["guess"], prior=ora.make_prior("bdf"),
                           nthreads=6)
    assert np.array_equal(res["flags"], ref["flags"])
    assert np.array_equal(res["npix"], ref["npix"])
    for i in range(6):
        if ref["flags"][i] == 0:
            # bdf is degenerate in (T, fracdev): compare on the sigma scale
            assert _close(res["pars"][i], ref["pars"][i], ref["pars_err"][i],
                          7, frac=1e-2), i


def test_config5_metacal_style(ngx, ora):
    from ngmix_b200 import sims
    b = sims.make_metacal_batch(12, seed0=5100)
    res = ngx.lm_fit("exp", b["images"], b["weights"], b["jac"], b["psf"],
                     b["guess"])
    ref = ora.lm_fit_batch("exp", b["images"], b["weights"], b["jac"],
                           b["psf"], b["guess"], nthreads=8)
    assert np.array_equal(res["flags"], ref["flags"])
    assert np.all(res["flags"] == 0)
    assert np.mean((res["nfev"] == ref["nfev"])
                   & (res["ier"] == ref["ier"])) >= 0.95
    for i in range(res.size):
        assert _close(res["pars"][i], ref["pars"][i], ref["pars_err"][i], 6)
    # response sanity: R11 = (<g1>_1p - <g1>_1m) / 0.02 ~ 1, same on both
    g1 = res["pars"][:, 2].reshape(-1, 5)
    g1r = ref["pars"][:, 2].reshape(-1, 5)
    R11 = (g1[:, 1] - g1[:, 2]).mean() / 0.02
    R11r = (g1r[:, 1] - g1r[:, 2]).mean() / 0.02
    assert abs(R11 - R11r) < 1e-6
    assert 0.5 < R11 < 1.5


def test_empty_batch(ngx):
    res = ngx.lm_fit("exp", np.zeros((0, 8, 8)), np.zeros((0, 8, 8)),
                     np.zeros((0, 8)), np.zeros((0, 1, 6)), np.zeros((0, 6)))
    assert res.shape == (0,)
    r, p = ngx.em(np.zeros((0, 8, 8)), np.zeros((0, 8, 8)), np.zeros((0, 8)),
                  np.zeros((0, 2, 6)))
    assert r.shape == (0,) and p.shape == (0, 2, 6)
    a = ngx.admom(np.zeros((0, 8, 8)), np.zeros((0, 8, 8)), np.zeros((0, 8)),
                  np.zeros((0, 6)))
    assert a.shape == (0,)
    # the float32, coelliptical, multi-band (narrow and wide records) and
    # noise-covariance entry points
    z32 = np.zeros((0, 8, 8), dtype=np.float32)
    assert ngx.lm_fit("exp", z32, z32, np.zeros((0, 8)), np.zeros((0, 1, 6)),
                      np.zeros((0, 6))).shape == (0,)
    assert ngx.lm_fit_coellip(2, np.zeros((0, 8, 8)), np.zeros((0, 8, 8)),
                              np.zeros((0, 8)), None,
                              np.zeros((0, 8))).shape == (0,)
    for nband in (2, 5):
        r = ngx.lm_fit_multi("exp", [0], np.zeros(0, "i4"), nband,
                             np.zeros((0, 8, 8)), np.zeros((0, 8, 8)),
                             np.zeros((0, 8)), None, np.zeros((0, 5 + nband)))
        assert r.shape == (0,)
    assert ngx.noise_b(np.zeros((0, 3, 64)), np.zeros((0, 8, 8)),
                       np.zeros((0, 8, 8))).shape == (0, 3, 3)


def test_tiny_stamps(ngx, ora):
    """the smallest stamps MINPACK accepts for six parameters (it refuses
    fewer residuals than parameters): 3 x 3 and 4 x 2 go through every pass
    and report what the oracle reports; a 1 x 1 stamp runs without faults"""
    one = ngx.lm_fit("gauss", np.ones((2, 1, 1)), np.ones((2, 1, 1)),
                     np.array([[0.0, 0.0, 0.263, 0.0, 0.0, 0.263,
                                0.263 ** 2, 0.263]] * 2), None,
                     np.array([[0.0, 0.0, 0.0, 0.0, 0.3, 1.0]] * 2))
    assert np.all(one["flags"] != 0)
    from ngmix_b200 import sims
    for dims in ((5, 5), (6, 4)):
        b = sims.make_lm_batch(6, seed0=990 + dims[0], dims=dims, noise=0.05,
                               model="gauss")
        res = ngx.lm_fit("gauss", b["images"], b["weights"], b["jac"],
                         b["psf"], b["guess"])
        ref = ora.lm_fit_batch("gauss", b["images"], b["weights"], b["jac"],
                               b["psf"], b["guess"])
        # the fitted point and the evaluation counts; the covariance of a
        # fit this poorly constrained is singular to rounding, so its flag
        # bits (512, 1024, 2048) are not compared
        mask = ~np.int32(512 | 1024 | 2048)
        assert np.array_equal(res["flags"] & mask, ref["flags"] & mask)
        assert np.array_equal(res["npix"], ref["npix"])
        assert (res["nfev"] == ref["nfev"]).sum() >= 5
        gm, _ = ngx.gmix_fill("gauss", b["guess"])
        ngx.gmix_set_norms(gm)
        out4, fd, st = ngx.loglike(gm, b["images"], b["weights"], b["jac"],
                                   want_fdiff=True)
        assert not st.any() and np.all(out4[:, 3] == dims[0] * dims[1])
        ag = np.array([[1.0, 0, 0, 0.2, 0, 0.2]] * 6)
        a = ngx.admom(b["images"], b["weights"], b["jac"], ag)
        oa = ora.admom_go_batch(b["images"], b["weights"], b["jac"], ag)
        assert np.array_equal(a["flags"], oa["flags"])
        assert np.array_equal(a["numiter"], oa["numiter"])


@pytest.mark.parametrize("dims", [(31, 47), (16, 16), (9, 11)])
def test_ragged_dims_no_psf_masked(ngx, ora, dims):
    """non-square / tiny stamps, gauss model without psf, masked pixels"""
    from ngmix_b200 import sims
    n = 10
    b = sims.make_lm_batch(n, seed0=6100 + dims[0], dims=dims, model="gauss",
                           psf_model="gauss", noise=0.05)
    rng = np.random.RandomState(dims[1])
    b["weights"][rng.uniform(size=b["weights"].shape) < 0.1] = 0.0
    b["weights"][0, :, :] = 0.0           # nothing to fit: NO_ATTEMPT
    res = ngx.lm_fit("gauss", b["images"], b["weights"], b["jac"], None,
                     b["guess"])
    assert res["flags"][0] == 1 and res["nfev"][0] == 0
    ref = ora.lm_fit_batch("gauss", b["images"][1:], b["weights"][1:],
                           b["jac"][1:], None, b["guess"][1:], nthreads=4)
    r = res[1:]
    assert np.array_equal(r["flags"], ref["flags"])
    assert np.array_equal(r["npix"], ref["npix"])
    ok = ref["flags"] == 0
    for i in np.where(ok)[0]:
        assert _close(r["pars"][i], ref["pars"][i], ref["pars_err"][i], 6)


def test_full_size_properties(ngx):
    """size-independent properties on a full 10k-stamp batch: determinism,
    independence of the batch split (sharding), linearity of the model in
    the flux, s2n consistency"""
    from ngmix_b200 import sims
    n = 10000
    b = sims.make_lm_batch(n, seed0=10 ** 6, truth_images=False)
    gm0, _ = ngx.gmix_fill("exp", b["truth"])
    gp, _ = ngx.gmix_fill("full", b["psf"].reshape(n, -1))
    gm = ngx.gmix_convolve(gm0, gp)
    images, st = ngx.render(gm, b["jac"], (48, 48), fast_exp=False)
    assert not st.any()
    rng = np.random.RandomState(12345)
    images += rng.normal(0.0, 0.01, size=images.shape)
    res = ngx.lm_fit("exp", images, b["weights"], b["jac"], b["psf"],
                     b["guess"])
    assert np.all(res["flags"] == 0)
    # idempotence / determinism: same call, bit-identical records
    res2 = ngx.lm_fit("exp", images, b["weights"], b["jac"], b["psf"],
                      b["guess"])
    assert res.tobytes() == res2.tobytes()
    # shard independence: any contiguous split gives the same records
    lo, hi = 3333, 7777
    part = ngx.lm_fit("exp", images[lo:hi], b["weights"][lo:hi],
                      b["jac"][lo:hi], b["psf"][lo:hi], b["guess"][lo:hi])
    assert part.tobytes() == res[lo:hi].tobytes()
    # the fits recover the truth within the errors (5 sigma, 99.9%)
    dev = np.abs(res["pars"][:, 2:6] - b["truth"][:, 2:]) \
        / res["pars_err"][:, 2:6]
    assert np.mean(dev < 5) > 0.999
    # linearity in the flux: doubling image and flux guess doubles the flux
    sub = slice(0, 256)
    r2 = ngx.lm_fit("exp", 2 * images[sub], b["weights"][sub] / 4,
                    b["jac"][sub], b["psf"][sub],
                    b["guess"][sub] * np.array([1, 1, 1, 1, 1, 2.0]))
    assert np.allclose(r2["pars"][:, 5], 2 * res["pars"][sub, 5], rtol=1e-9)
    assert np.allclose(r2["pars"][:, 2:5], res["pars"][sub, 2:5], rtol=1e-7,
                       atol=1e-10)
    assert np.array_equal(r2["nfev"], res["nfev"][sub])


@pytest.mark.parametrize("noise,prior_kind", [(1.0, None), (1.0, "flat"),
                                              (0.3, "erf")])
def test_config2_low_s2n_and_priors(ngx, ora, noise, prior_kind):
    """config 2's second pass: sigma = 1.0 (S/N 10-20, fits that take up to
    ~40 evaluations) with and without the reference tests' prior sets; flags,
    nfev and ier bit-exact, parameters to 1e-5"""
    from ngmix_b200 import sims
    n = 400
    b = sims.make_lm_batch(n, seed0=123000, noise=noise)
    pr = opr = None
    if prior_kind:
        opr = ora.make_prior(prior_kind)
        pr = ngx.Prior()
        pr.present, pr.g_kind = 1, 1
        pr.cen1 = pr.cen2 = 0.0
        pr.sigma1 = pr.sigma2 = 0.263
        pr.g_sigma = 0.3
        for name in ("T", "F"):
            src = getattr(opr, name)
            setattr(pr, name, ngx.Prior1D(src.kind, 0, src.a, src.b, src.c,
                                          src.d))
    res = ngx.lm_fit("exp", b["images"], b["weights"], b["jac"], b["psf"],
                     b["guess"], prior=pr)
    ref = ora.lm_fit_batch("exp", b["images"], b["weights"], b["jac"],
                           b["psf"], b["guess"], prior=opr, nthreads=8)
    assert np.array_equal(res["flags"], ref["flags"])
    assert np.array_equal(res["nfev"], ref["nfev"])
    assert np.array_equal(res["ier"], ref["ier"])
    assert ref["nfev"].max() > 6 or noise < 1.0     # the slow fits are there
    ok = ref["flags"] == 0
    tol = np.maximum(PAR_TOL * np.abs(ref["pars"][:, :6]),
                     1e-3 * ref["pars_err"][:, :6])
    assert np.all(np.abs(res["pars"][:, :6] - ref["pars"][:, :6])[ok]
                  <= tol[ok])


def test_stamps_larger_than_shared_memory(ngx, ora):
    """stamps that do not fit an SM's shared memory take the global-memory
    paths (pixel records for loglike, scratch buffers for EM / admom, the
    multi-epoch kernel for the analytic LM fit) and give the same answers"""
    from ngmix_b200 import sims
    # LM, 200 x 200
    b = sims.make_lm_batch(3, seed0=808, dims=(200, 200), noise=0.05)
    res = ngx.lm_fit("exp", b["images"], b["weights"], b["jac"], b["psf"],
                     b["guess"])
    ref = ora.lm_fit_batch("exp", b["images"], b["weights"], b["jac"],
                           b["psf"], b["guess"])
    assert np.array_equal(res["flags"], ref["flags"])
    assert np.array_equal(res["nfev"], ref["nfev"])
    err = np.abs(res["pars"][:, :6] - ref["pars"][:, :6])
    assert np.all(err <= np.maximum(1e-5 * np.abs(ref["pars"][:, :6]),
                                    1e-3 * ref["pars_err"][:, :6]))
    assert np.allclose(res["lnprob"], ref["lnprob"], rtol=1e-9)
    # forward-difference LM (bdf with prior), 128 x 128: value and base-model
    # planes in the per-CTA global scratch
    b = sims.make_bdf_batch(3, seed0=810, dims=(128, 128), noise=0.02)
    res = ngx.lm_fit("bdf", b["images"], b["weights"], b["jac"], b["psf"],
                     b["guess"], prior=_bdf_prior(ngx))
    ref = ora.lm_fit_batch("bdf", b["images"], b["weights"], b["jac"],
                           b["psf"], b["guess"], prior=ora.make_prior("bdf"))
    assert np.array_equal(res["flags"], ref["flags"])
    assert np.array_equal(res["nfev"], ref["nfev"])
    err = np.abs(res["pars"][:, :7] - ref["pars"][:, :7])
    assert np.all(err <= np.maximum(1e-5 * np.abs(ref["pars"][:, :7]),
                                    1e-2 * ref["pars_err"][:, :7]))
    # loglike + fdiff, 128 x 128 with masked pixels
    b = sims.make_lm_batch(2, seed0=809, dims=(128, 128), noise=0.05)
    b["weights"][:, 5:9, 100:120] = 0.0
    gm0, _ = ngx.gmix_fill("exp", b["guess"])
    gp, _ = ngx.gmix_fill("full", b["psf"].reshape(2, -1))
    gm = ngx.gmix_convolve(gm0, gp)
    ngx.gmix_set_norms(gm)
    out4, fdiff, st = ngx.loglike(gm, b["images"], b["weights"], b["jac"],
                                  want_fdiff=True)
    assert not st.any()
    for i in range(2):
        og0, _ = ora.gmix_fill("exp", b["guess"][i])
        ogp, _ = ora.gmix_fill("full", b["psf"][i].ravel())
        og = ora.gmix_convolve(og0, ogp)
        ora.gmix_set_norms(og)
        pix = ora.make_pixels(b["images"][i], b["weights"][i], b["jac"][i])
        want = np.zeros(pix.size)
        ora.fill_fdiff(og, pix, want)
        o4, _ = ora.get_loglike(og, pix)
        assert out4[i][3] == pix.size
        assert np.allclose(out4[i][:3], o4[:3], rtol=1e-11)
        assert np.abs(fdiff[i][:pix.size] - want).max() \
            <= 1e-11 * np.abs(want).max()
    # EM and admom, 100 x 100 / 128 x 128 PSF-like stamps
    for dims in ((100, 100), (128, 128)):
        n = 2
        rng = np.random.RandomState(dims[0])
        jac = np.array([sims.diag_jacobian((dims[0] - 1) / 2 + 0.2,
                                           (dims[1] - 1) / 2 - 0.3, 0.263)] * n)
        full = sims.simple_full_pars([0.0, 0.0, 0.05, -0.03, 4.0, 1.0], "turb")
        images = np.array([sims.render_full_pars(full, dims, jac[i])
                           + rng.normal(0, 1e-5, dims) for i in range(n)])
        weights = np.full(images.shape, 1e10)
        guess = np.array([[[0.6, 0.0, 0.0, 1.5, 0.0, 1.5],
                           [0.4, 0.0, 0.0, 3.0, 0.0, 3.0]]] * n)
        res, pars = ngx.em(images, weights, jac, guess, maxiter=200)
        ores, opars = ora.em_go_batch(images, weights, jac, guess,
                                      maxiter=200)
        assert np.array_equal(res["flags"], ores["flags"])
        assert np.array_equal(res["numiter"], ores["numiter"])
        assert np.allclose(pars.reshape(n, -1), opars.reshape(n, -1),
                           rtol=1e-9, atol=1e-12)
        ag = np.array([[1.0, 0.0, 0.0, 2.0, 0.0, 2.0]] * n)
        ares = ngx.admom(images, weights, jac, ag)
        oam = ora.admom_go_batch(images, weights, jac, ag)
        assert np.array_equal(ares["flags"], oam["flags"])
        assert np.array_equal(ares["numiter"], oam["numiter"])
        assert np.allclose(ares["pars"], oam["pars"], rtol=1e-9, atol=1e-12)


def test_more_objects_than_a_grid_dimension(ngx):
    """batches beyond 65 535 objects (the y / z limit of a CUDA grid): the
    kernels that use a two-dimensional grid keep the objects on x"""
    n = 70000
    rng = np.random.RandomState(2)
    pars = np.zeros((n, 6))
    pars[:, 4] = rng.uniform(0.2, 0.5, n)
    pars[:, 5] = rng.uniform(1.0, 2.0, n)
    jac = np.array([[1.5, 1.5, 0.263, 0.0, 0.0, 0.263, 0.263 ** 2,
                     0.263]] * n)
    gm, st = ngx.gmix_fill("gauss", pars)
    assert not st.any()
    im, st = ngx.render(gm, jac, (4, 4), fast_exp=True)
    assert not st.any()
    tail, _ = ngx.render(gm[-3:], jac[-3:], (4, 4), fast_exp=True)
    assert np.array_equal(im[-3:], tail) and im[-1].sum() > 0
    coords = ngx.make_coords((4, 4), jac)
    assert np.array_equal(coords[-1], coords[0])
    gpars = np.stack([gm[k] for k in ("p", "row", "col", "irr", "irc", "icc")],
                     axis=-1)
    dcov = np.zeros((n, 1, 3, 3))
    dcov[:, 0, 2, 0] = dcov[:, 0, 2, 2] = 0.5
    out = ngx.deriv_images(gpars, dcov, jac, (4, 4))
    assert np.allclose(out[:, 0].reshape(n, 4, 4), im, rtol=1e-12)
    recs = ngx.lm_fit("gauss", im, np.full(im.shape, 1e4), jac, None,
                      pars * 1.01)
    assert (recs["flags"] == 0).sum() > 0.99 * n


def test_config4_bdf_parity_2000_stamps(ora):
    """
    BASELINE configs[3] at test scale: 2 000 seeded bdf fits on 64x64 stamps
    (48 gaussians, PriorBDFSep, MINPACK's forward-difference Jacobian) against
    the oracle on identical pixels.  flags must agree on every stamp.  The
    forward-difference Jacobian divides 1e-16 rounding differences of the
    model by steps of 1.5e-8 |x|, so evaluation counts and parameters follow
    the reference up to that noise: nfev and ier agree on >= 98 % of the
    stamps (the oracle compiled with fused multiply-adds agrees with itself
    on the same fraction, profiles/parity_sweep_r2c.json), parameters within
    max(1e-5 |p|, 1e-3 sigma) on >= 97 %, and T and flux within 1e-5 relative
    on >= 99 %.
    """
    import os
    import torch
    import bench_configs
    from ngmix_b200 import _lib, sims_device
    _lib.init(0)
    n = 2000
    dev = torch.device("cuda", 0)
    t = sims_device.make_bdf_batch_device(n, 4100000, dev)
    d_res = torch.zeros(n * _lib.LM_RESULT_DTYPE.itemsize, dtype=torch.uint8,
                        device=dev)
    _lib.lm_fit_device("bdf", n, t["images"].data_ptr(),
                       t["weights"].data_ptr(), t["jac"].data_ptr(), 64, 64,
                       t["psf"].data_ptr(), 3, t["guess"].data_ptr(),
                       d_res.data_ptr(), prior=bench_configs.bdf_prior(_lib),
                       analytic_jacobian=False)
    torch.cuda.synchronize()
    res = np.frombuffer(d_res.cpu().numpy().tobytes(),
                        dtype=_lib.LM_RESULT_DTYPE)
    ref = ora.lm_fit_batch("bdf", t["images"].cpu().numpy(),
                           t["weights"].cpu().numpy(), t["h_jac"], t["h_psf"],
                           t["h_guess"], prior=ora.make_prior("bdf"),
                           nthreads=os.cpu_count())
    p = bench_configs.lm_parity(res, ref, 7, Tcol=4, Fcol=6)
    assert p["flags_equal"] == n
    assert p["n_good"] >= 0.99 * n
    assert p["nfev_ier_equal"] >= 0.98 * n
    assert p["pars_within_tol"] >= 0.97 * p["n_good"]
    assert p["n_rel_dev_T_or_flux_above_1e-5"] <= 0.01 * p["n_good"]
    assert p["max_abs_dev_g"] < 1e-4
