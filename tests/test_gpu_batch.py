"""
ObservationBatch: the resident-batch entry points give
exactly what the host-buffer calls give, and a PSF EM -> galaxy LM chain runs
without the PSF mixtures leaving the device.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ngx():
    from ngmix_b200 import _lib
    _lib.init(0)
    return _lib


def _same_records(a, b):
    assert a.dtype == b.dtype and a.shape == b.shape
    for name in a.dtype.names:
        assert np.array_equal(a[name], b[name], equal_nan=True), name


def test_batch_fit_equals_host_call(ngx):
    import ngmix_b200 as ngmix
    from ngmix_b200 import sims
    b = sims.make_lm_batch(40, seed0=8100, noise=0.02)
    ob = ngmix.ObservationBatch.from_arrays(b["images"], b["weights"],
                                            b["jac"], psf_pars=b["psf"])
    assert len(ob) == 40 and ob.dims == (48, 48) and ob.npsf == 3
    got = ob.fit("exp", b["guess"])
    ref = ngmix.fit_batch("exp", b["images"], b["weights"], b["jac"],
                          b["psf"], b["guess"])
    _same_records(got, ref)
    # a second fitter on the same resident stamps, with a prior
    prior = ngmix.joint_prior.PriorSimpleSep(
        ngmix.priors.CenPrior(0.0, 0.0, 0.263, 0.263),
        ngmix.priors.GPriorBA(0.3),
        ngmix.priors.FlatPrior(-1.0, 1.0e3),
        ngmix.priors.FlatPrior(-100.0, 1.0e9))
    got = ob.fit("gauss", b["guess"], prior=prior)
    ref = ngmix.fit_batch("gauss", b["images"], b["weights"], b["jac"],
                          b["psf"], b["guess"], prior=prior)
    _same_records(got, ref)
    # shards are views of the same resident arrays
    s1 = ob.shard(1, 2).fit("exp", b["guess"][20:])
    _same_records(s1, ngmix.fit_batch("exp", b["images"][20:],
                                      b["weights"][20:], b["jac"][20:],
                                      b["psf"][20:], b["guess"][20:]))


def test_batch_em_admom_equal_host_call(ngx):
    import ngmix_b200 as ngmix
    from ngmix_b200 import sims
    b = sims.make_psf_batch(32, seed0=8200)
    ob = ngmix.ObservationBatch.from_arrays(b["images"], b["weights"],
                                            b["jac"])
    res, pars = ob.em(b["em_guess"])
    rres, rpars = ngx.em(b["images"], b["weights"], b["jac"], b["em_guess"])
    _same_records(res, rres)
    assert np.array_equal(pars, rpars)
    am = ob.admom(b["am_guess"])
    _same_records(am, ngx.admom(b["images"], b["weights"], b["jac"],
                                b["am_guess"]))
    pix, npix = ob.make_pixels()
    assert np.all(npix == 625) and pix.shape == (32, 625)


def test_em_admom_host_calls_pipelined(ngx):
    """the host calls move the planes in chunks on two streams (pageable
    planes through the copy pool): several chunks, a short last one, pageable
    and pinned planes give the records of the resident batch"""
    import ngmix_b200 as ngmix
    from ngmix_b200 import sims
    n = 15000          # chunks of 6710 stamps at 25x25
    b = sims.make_psf_batch(64, seed0=8300)
    rep = (n + 63) // 64
    tile = lambda a: np.ascontiguousarray(
        np.concatenate([a] * rep, axis=0)[:n])
    im, wt, jac = tile(b["images"]), tile(b["weights"]), tile(b["jac"])
    im += np.random.RandomState(1).normal(0.0, 1e-5, size=im.shape)
    ag = tile(b["am_guess"])
    eg = tile(b["em_guess"])
    ob = ngmix.ObservationBatch.from_arrays(im, wt, jac)
    want = ob.admom(ag)
    _same_records(ngx.admom(im, wt, jac, ag), want)
    pim, pwt = ngmix.pinned_copy(im), ngmix.pinned_copy(wt)
    _same_records(ngx.admom(pim, pwt, jac, ag), want)
    conf_kw = dict(maxiter=40)
    wres, wpars = ob.em(eg, **conf_kw)
    res, pars = ngx.em(im, wt, jac, eg, **conf_kw)
    _same_records(res, wres)
    assert np.array_equal(pars, wpars)
    res, pars = ngx.em(pim, pwt, jac, eg, **conf_kw)
    _same_records(res, wres)
    assert np.array_equal(pars, wpars)


def test_psf_em_to_galaxy_fit_chain_on_device(ngx):
    """PSF stamps -> EM mixtures (device) -> psf of the galaxy batch -> LM"""
    import ngmix_b200 as ngmix
    from ngmix_b200 import sims
    n = 24
    g = sims.make_lm_batch(n, seed0=8300, noise=0.02)
    p = sims.make_psf_batch(n, seed0=8300)
    psf_batch = ngmix.ObservationBatch.from_arrays(p["images"], p["weights"],
                                                   p["jac"])
    res, d_pars = psf_batch.em(p["em_guess"], return_device=True)
    gal = ngmix.ObservationBatch.from_arrays(g["images"], g["weights"],
                                             g["jac"])
    gal.set_psf(d_pars)              # no host round trip
    got = gal.fit("exp", g["guess"])
    ref = ngmix.fit_batch("exp", g["images"], g["weights"], g["jac"],
                          d_pars.cpu().numpy(), g["guess"])
    _same_records(got, ref)
    assert np.all(got["flags"][res["flags"] == 0] >= 0)


def test_batch_gaussmom_equals_host_call(ngx):
    import ngmix_b200 as ngmix
    from ngmix_b200 import sims
    b = sims.make_lm_batch(16, seed0=8500, noise=0.05)
    ob = ngmix.ObservationBatch.from_arrays(b["images"], b["weights"],
                                            b["jac"])
    for higher in (False, True):
        got = ob.gaussmom(1.2, with_higher_order=higher)
        ref = ngmix.gaussmom_batch(b["images"], b["weights"], b["jac"], 1.2,
                                   with_higher_order=higher)
        for a, r in zip(got, ref):
            assert a["flags"] == r["flags"]
            assert np.array_equal(a["sums"], r["sums"])
            assert np.array_equal(a["sums_cov"], r["sums_cov"])
            assert a["T"] == r["T"] and a["e1"] == r["e1"]


def test_from_observations(ngx):
    import ngmix_b200 as ngmix
    from ngmix_b200 import sims
    b = sims.make_lm_batch(5, seed0=8400, noise=0.02)
    obslist = []
    for i in range(5):
        jac = ngmix.Jacobian(row=b["jac"][i, 0], col=b["jac"][i, 1],
                             dvdrow=b["jac"][i, 2], dvdcol=b["jac"][i, 3],
                             dudrow=b["jac"][i, 4], dudcol=b["jac"][i, 5])
        psf_gm = ngmix.GMix(pars=b["psf"][i].ravel())
        psf_obs = ngmix.Observation(np.zeros((5, 5)) + 1.0, gmix=psf_gm)
        obslist.append(ngmix.Observation(b["images"][i], weight=b["weights"][i],
                                         jacobian=jac, psf=psf_obs))
    ob = ngmix.ObservationBatch.from_observations(obslist)
    got = ob.fit("exp", b["guess"])
    ref = ngmix.fit_batch("exp", b["images"], b["weights"], b["jac"],
                          b["psf"], b["guess"])
    _same_records(got, ref)


def test_run_fit_batch_retries_only_failures(ngx):
    """batched Runner: the second round refits exactly the objects that
    failed the first (here: a guess outside the model's domain), and what
    comes back equals fitting everything from the good guesses"""
    import ngmix_b200 as ng
    from ngmix_b200 import sims
    from ngmix_b200.runners import run_fit_batch
    b = sims.make_lm_batch(12, seed0=5150, noise=0.05, dims=(32, 32))
    bad = np.array([1, 4, 7])
    calls = []

    def guesser(idx, itry):
        calls.append(idx.copy())
        g = b["guess"][idx].copy()
        if itry == 0:
            g[np.isin(idx, bad), 4] = -1.0          # T < 0: range error
        return g

    recs, tries = run_fit_batch("exp", b["images"], b["weights"], b["jac"],
                                b["psf"], guesser, ntry=3)
    assert len(calls) == 2 and np.array_equal(calls[1], bad)
    assert np.array_equal(tries, np.where(np.isin(np.arange(12), bad), 2, 1))
    want = ng.fit_batch("exp", b["images"], b["weights"], b["jac"], b["psf"],
                        b["guess"])
    assert not recs["flags"].any()
    assert np.array_equal(recs["pars"], want["pars"])


def test_lm_fit_float32_host_stamps(ngx):
    """float32 host stamps (ngx_lm_fit_host_f32) are widened exactly: the
    records equal those of the float64 call on the same values, over several
    pipeline chunks and an odd pixel count"""
    from ngmix_b200 import sims
    b = sims.make_lm_batch(1500, seed0=6200, noise=0.05, dims=(21, 23),
                           truth_images=False)
    rng = np.random.RandomState(4)
    gm0, _ = ngx.gmix_fill("exp", b["truth"])
    gp, _ = ngx.gmix_fill("full", b["psf"].reshape(1500, -1))
    images, _ = ngx.render(ngx.gmix_convolve(gm0, gp), b["jac"], (21, 23))
    images += rng.normal(0.0, 0.05, images.shape)
    im32 = images.astype(np.float32)
    wt32 = b["weights"].astype(np.float32)
    wt32[:, 3, 4:7] = 0.0
    r64 = ngx.lm_fit("exp", im32.astype(np.float64), wt32.astype(np.float64),
                     b["jac"], b["psf"], b["guess"])
    r32 = ngx.lm_fit("exp", im32, wt32, b["jac"], b["psf"], b["guess"])
    _same_records(r32, r64)
    assert (r32["flags"] == 0).sum() > 1400


def test_lm_fit_compact_survey_format(ngx):
    """float32 pixels + one weight per stamp + a packed good-pixel mask
    (ngx_lm_fit_host_compact) give the records of the expanded float64
    arrays bit for bit, over several pipeline chunks, through fit_batch too"""
    import ngmix_b200 as ng
    from ngmix_b200 import sims
    n = 1400
    b = sims.make_lm_batch(n, seed0=6300, noise=0.05, dims=(21, 23),
                           truth_images=False)
    rng = np.random.RandomState(5)
    gm0, _ = ngx.gmix_fill("exp", b["truth"])
    gp, _ = ngx.gmix_fill("full", b["psf"].reshape(n, -1))
    images, _ = ngx.render(ngx.gmix_convolve(gm0, gp), b["jac"], (21, 23))
    images += rng.normal(0.0, 0.05, images.shape)
    im32 = images.astype(np.float32)
    wsc = rng.uniform(300.0, 500.0, size=n)
    mask = rng.uniform(size=im32.shape) > 0.03
    mask[7] = True
    wt = np.where(mask, wsc[:, None, None], 0.0)
    want = ngx.lm_fit("exp", im32.astype(np.float64), wt, b["jac"], b["psf"],
                      b["guess"])
    got = ngx.lm_fit_compact("exp", im32, wsc, b["jac"], b["psf"],
                             b["guess"], mask=mask)
    _same_records(got, want)
    got2 = ng.fit_batch("exp", im32, wsc, b["jac"], b["psf"], b["guess"],
                        mask=mask)
    _same_records(got2, want)
    # the same format into a device-resident batch
    ob = ng.ObservationBatch.from_arrays(im32, wsc, b["jac"],
                                         psf_pars=b["psf"], mask=mask)
    _same_records(ob.fit("exp", b["guess"]), want)
    # no mask, scalar weight for the whole batch
    want3 = ngx.lm_fit("exp", im32.astype(np.float64),
                       np.full(im32.shape, 400.0), b["jac"], b["psf"],
                       b["guess"])
    got3 = ng.fit_batch("exp", im32, 400.0, b["jac"], b["psf"], b["guess"])
    _same_records(got3, want3)
    # float64 pixels with a per-stamp weight go through the expanded path
    got4 = ng.fit_batch("exp", im32.astype(np.float64), wsc, b["jac"],
                        b["psf"], b["guess"], mask=mask)
    _same_records(got4, want)
    assert (want["flags"] == 0).sum() > 1300


def test_pinned_host_memory_helpers(ngx):
    import ngmix_b200 as ng
    from ngmix_b200 import sims
    b = sims.make_lm_batch(64, seed0=6400, noise=0.05, dims=(24, 24))
    want = ngx.lm_fit("exp", b["images"], b["weights"], b["jac"], b["psf"],
                      b["guess"])
    pim = ng.pinned_copy(b["images"])
    pwt = ng.pinned_copy(b["weights"])
    assert pim.shape == b["images"].shape and pim.dtype == np.float64
    got = ngx.lm_fit("exp", pim, pwt, b["jac"], b["psf"], b["guess"])
    _same_records(got, want)
    im, wt = b["images"].copy(), b["weights"].copy()
    with ng.pinned(im, wt):
        got = ngx.lm_fit("exp", im, wt, b["jac"], b["psf"], b["guess"])
    _same_records(got, want)
    e = ng.pinned_empty((3, 5), "f4")
    e[...] = 1.5
    assert e.sum() == 22.5
    del e, pim, pwt


def test_streamed_host_pipeline(ngx, monkeypatch):
    """pinned stamps take the streamed pipeline (one persistent launch fed by
    the copy engine, csrc/ngx_lm.cu:lm_fit_host_stream): the same records, to
    the bit, as the chunked pipeline, for every host format, ragged chunk
    boundaries, several batches per call, stamps without a usable pixel and a
    forward-difference model"""
    import ngmix_b200 as ng
    from ngmix_b200 import sims
    from ngmix_b200.flags import NO_ATTEMPT
    n = 1501
    b = sims.make_lm_batch(n, seed0=8800, noise=0.05)
    im32 = b["images"].astype(np.float32)
    wt32 = b["weights"].astype(np.float32)
    wt32[7] = 0.0                      # GMixFatalError("no weights > 0")
    wt32[1500] = 0.0
    wt32[33, 10:20, 5:9] = 0.0
    mask = np.random.RandomState(5).uniform(size=im32.shape) > 0.03
    wsc = wt32[:, 0, 0].astype(np.float64)
    wsc[7] = 1.0
    pin = ng.pinned_copy
    cases = {
        "f64": (ngx.lm_fit, (pin(im32.astype("f8")), pin(wt32.astype("f8"))), {}),
        "f32": (ngx.lm_fit, (pin(im32), pin(wt32)), {}),
        "compact": (ngx.lm_fit_compact, (pin(im32), pin(wsc)), {"mask": mask}),
    }
    small = (pin(b["jac"]), pin(b["psf"]), pin(b["guess"]))
    # small calls keep the chunked pipeline unless told otherwise
    monkeypatch.setenv("NGX_STREAM_ALWAYS", "1")
    for name, (fn, planes, kw) in cases.items():
        monkeypatch.setenv("NGX_NO_STREAM", "1")
        want = fn("exp", *planes, *small, **kw)
        monkeypatch.setenv("NGX_NO_STREAM", "0")
        _same_records(fn("exp", *planes, *small, **kw), want)
        # pageable small arrays and a pinned result array
        out = ng.pinned_empty((n,), ngx.LM_RESULT_DTYPE)
        got = fn("exp", *planes, b["jac"], b["psf"], b["guess"], out=out, **kw)
        _same_records(got, want)
        # ordinary (pageable) numpy stamps: staged through pinned slots by
        # the copy pool (csrc/ngx_copy_pool.h)
        pageable = [np.array(p) for p in planes]
        _same_records(fn("exp", *pageable, b["jac"], b["psf"], b["guess"],
                         **kw), want)
        # several batches per call, the last one short
        monkeypatch.setenv("NGX_STREAM_BATCH", "400")
        _same_records(fn("exp", *pageable, b["jac"], b["psf"], b["guess"],
                         **kw), want)
        _same_records(fn("exp", *planes, *small, **kw), want)
        monkeypatch.delenv("NGX_STREAM_BATCH")
        if name != "compact":
            assert want["flags"][7] == NO_ATTEMPT
            assert want["flags"][1500] == NO_ATTEMPT
        assert (want["flags"] == 0).sum() > 1400
    # forward differences (lmdif) from streamed stamps
    bb = sims.make_bdf_batch(300, seed0=8900, dims=(40, 40))
    args = [pin(bb[k]) for k in ("images", "weights", "jac", "psf", "guess")]
    monkeypatch.setenv("NGX_NO_STREAM", "1")
    want = ngx.lm_fit("bdf", *args)
    monkeypatch.setenv("NGX_NO_STREAM", "0")
    _same_records(ngx.lm_fit("bdf", *args), want)


def test_streamed_pipeline_failure_releases_the_device(ngx, monkeypatch):
    """a feed that stops half way (test hook) must fail the call promptly -
    the resident kernel's waiters are released through the error flag, not
    by the 4 s timeout per stamp - and the next call must work"""
    import time
    import ngmix_b200 as ng
    from ngmix_b200 import sims
    n = 3000
    b = sims.make_lm_batch(n, seed0=8850, noise=0.05)
    pin = ng.pinned_copy
    args = [pin(b[k].astype("f4") if k in ("images", "weights") else b[k])
            for k in ("images", "weights", "jac", "psf", "guess")]
    want = ngx.lm_fit("exp", *args)
    # pageable stamps: the kernel is already resident and waiting when the
    # feed stops; pinned stamps: the copies are queued before the launch
    import os
    pageable = [np.array(a) for a in args]
    cases = [args] if os.environ.get("CUDA_LAUNCH_BLOCKING", "0") != "0" \
        else [pageable, args]      # blocking launches: pageable is chunked
    for aa in cases:
        monkeypatch.setenv("NGX_STREAM_FAIL_AT", "3")
        t0 = time.perf_counter()
        with pytest.raises(Exception):
            ngx.lm_fit("exp", *aa)
        assert time.perf_counter() - t0 < 3.0
        monkeypatch.delenv("NGX_STREAM_FAIL_AT")
        _same_records(ngx.lm_fit("exp", *aa), want)


def test_streamed_pipeline_is_not_slower_than_chunked(ngx, monkeypatch):
    """guard against feeding mistakes (a waiter polling host memory once
    starved the copies it waited for: 50x slower, same records)"""
    import time
    import ngmix_b200 as ng
    from ngmix_b200 import sims
    n = 6000
    b = sims.make_lm_batch(n, seed0=8870, noise=0.05)
    pin = ng.pinned_copy
    args = [pin(b[k].astype("f4") if k in ("images", "weights") else b[k])
            for k in ("images", "weights", "jac", "psf", "guess")]
    out = ng.pinned_empty((n,), ngx.LM_RESULT_DTYPE)

    def best(mode):
        monkeypatch.setenv("NGX_NO_STREAM", mode)
        ngx.lm_fit("exp", *args, out=out)
        ts = []
        for _ in range(4):
            t0 = time.perf_counter()
            ngx.lm_fit("exp", *args, out=out)
            ts.append(time.perf_counter() - t0)
        return min(ts)

    chunked, streamed = best("1"), best("0")
    assert streamed < 1.5 * chunked, (streamed, chunked)
