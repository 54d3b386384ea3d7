"""
The device-resident production chain (ngx_bootstrap / ObservationBatch.
bootstrap): psf EM -> psf-template flux -> guess -> LM with retries, replacing
bootstrap + PSFRunner + Runner of the reference (bootstrap.py:43-102,
runners.py:116-149, guessers.py:78-144, 205-261, 767-1051) for a batch.

Parity definition (a counter-based device generator
cannot reproduce a numpy RandomState stream, so the chain has its own):
 1. same draws given the same counter - the device's Philox4x32-10 stream of
    (seed, object, try, stage) equals the numpy restatement below bit for bit,
    and the guesses equal the reference guessers' recipes evaluated on those
    draws;
 2. given those guesses every stage equals the oracle: EM numiter / flags,
    template flux, LM flags / nfev / parameters - including which objects
    needed a second try.
"""
import numpy as np
import pytest

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c, k):
    """c [..., 4] uint64 holding 32-bit words, k (k0, k1) ints"""
    c0, c1, c2, c3 = (c[..., i].astype(np.uint64) for i in range(4))
    k0, k1 = int(k[0]), int(k[1])
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        n0 = hi1 ^ c1 ^ np.uint64(k0)
        n2 = hi0 ^ c3 ^ np.uint64(k1)
        c0, c1, c2, c3 = n0, lo1, n2, lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return np.stack([c0, c1, c2, c3], axis=-1)


def draws(seed, n, tr, stage, ndraw):
    """u[i, k]: draw k of stream (object i, try, stage)"""
    nb = (ndraw + 1) // 2
    c = np.zeros((n, nb, 4), dtype=np.uint64)
    c[:, :, 0] = np.arange(n, dtype=np.uint64)[:, None]
    c[:, :, 1] = tr
    c[:, :, 2] = np.arange(nb, dtype=np.uint64)[None, :]
    c[:, :, 3] = stage
    w = philox4x32_10(c, (seed & 0xFFFFFFFF, seed >> 32))
    a = (w[..., 0::2] >> np.uint64(5)).astype(np.float64)
    b = (w[..., 1::2] >> np.uint64(6)).astype(np.float64)
    u = (a * 67108864.0 + b) / 9007199254740992.0
    return u.reshape(n, 2 * nb)[:, :ndraw]


class Stream(object):
    def __init__(self, u):
        self.u, self.k = u, 0

    def uniform(self, lo, hi):
        v = lo + (hi - lo) * self.u[self.k]
        self.k += 1
        return v


def psf_guess_host(seed, i, tr, ngauss, scale, flux):
    """GMixPSFGuesser's recipe (guessers.py:832-842, 879-1030) on stream
    (seed, i, tr, 1)"""
    from ngmix_b200.guessers import _PSF_P, _PSF_F, _PSF_ENTRY
    d = Stream(draws(seed, i + 1, tr, 1, 6 * ngauss)[i])
    sigma = scale * 3.5 / 2.3548200450309493
    sigma2 = (2.0 * sigma * sigma) / 2
    if ngauss == 1:
        return np.array([flux * d.uniform(0.9, 1.1), d.uniform(-0.1, 0.1),
                         d.uniform(-0.1, 0.1),
                         sigma2 * (1.0 + d.uniform(-0.1, 0.1)),
                         d.uniform(-0.2 * sigma2, 0.2 * sigma2),
                         sigma2 * (1.0 + d.uniform(-0.1, 0.1))])
    out = np.zeros(6 * ngauss)
    for k, e in enumerate(_PSF_ENTRY[ngauss]):
        pf, ff = _PSF_P[ngauss][e], _PSF_F[ngauss][e]
        p = pf * flux if ngauss == 2 \
            else flux * pf * (1.0 + d.uniform(-0.1, 0.1))
        row, col = d.uniform(-0.1, 0.1), d.uniform(-0.1, 0.1)
        irr = ff * sigma2 * (1.0 + d.uniform(-0.1, 0.1))
        irc = 0.0 if ngauss == 2 else d.uniform(-0.01, 0.01)
        icc = ff * sigma2 * (1.0 + d.uniform(-0.1, 0.1))
        out[6 * k:6 * k + 6] = (p, row, col, irr, irc, icc)
    return out


def device_stamp_sum(im):
    """obs.image.sum() as k_stamp_sums associates it: 32 lane sums over
    pixels lane, lane + 32, ... in order, then an xor butterfly (the value
    of lane 0).  3-gaussian EM fits are chaotic in their iteration count, so
    the last bit of the flux guess matters for a bit-exact comparison."""
    flat = np.asarray(im, dtype="f8").ravel()
    parts = np.array([np.cumsum(flat[l::32])[-1] if flat[l::32].size else 0.0
                      for l in range(32)])
    idx = np.arange(32)
    for o in (16, 8, 4, 2, 1):
        parts = parts + parts[idx ^ o]
    return parts[0]


def obj_guess_host(seed, i, tr, T, flux):
    """TPSFFluxGuesser's recipe (guessers.py:116-134) on stream
    (seed, i, tr, 2)"""
    d = Stream(draws(seed, i + 1, tr, 2, 6)[i])
    return np.array([d.uniform(-0.01, 0.01), d.uniform(-0.01, 0.01),
                     d.uniform(-0.02, 0.02), d.uniform(-0.02, 0.02),
                     T * d.uniform(0.9, 1.1), flux * d.uniform(0.9, 1.1)])


def test_philox_known_answer():
    """Random123's known-answer vectors for philox4x32-10"""
    z = philox4x32_10(np.zeros((1, 4), dtype=np.uint64), (0, 0))[0]
    assert [int(x) for x in z] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c,
                                   0x9b00dbd8]
    f = np.full((1, 4), 0xFFFFFFFF, dtype=np.uint64)
    z = philox4x32_10(f, (0xFFFFFFFF, 0xFFFFFFFF))[0]
    assert [int(x) for x in z] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6,
                                   0x6d5451fd]


@pytest.fixture(scope="module")
def chain_inputs():
    from ngmix_b200 import sims
    n = 96
    b = sims.make_lm_batch(n, seed0=7700, noise=0.02)
    p = sims.make_psf_batch(n, seed0=7800)
    # make a few fits hard so that retries happen: dead psf stamps
    p["images"][5] = 0.0
    p["weights"][9] = 0.0
    return b, p


@pytest.mark.gpu
def test_device_draws_match_numpy():
    import torch
    from ngmix_b200 import _lib
    _lib.init(0)
    out = torch.zeros((50, 11), dtype=torch.float64, device="cuda")
    for seed, tr, stage in ((0, 0, 1), (123456789012345, 3, 2)):
        _lib.chain_draws_device(50, seed, tr, stage, 11, out.data_ptr())
        torch.cuda.synchronize()
        assert np.array_equal(out.cpu().numpy(),
                              draws(seed, 50, tr, stage, 11))


@pytest.mark.gpu
def test_bootstrap_chain_matches_oracle_stage_by_stage(ora, chain_inputs):
    import ngmix_b200 as ng
    from ngmix_b200 import _lib
    _lib.init(0)
    b, p = chain_inputs
    n = b["images"].shape[0]
    seed = 20261017
    objs = ng.ObservationBatch.from_arrays(b["images"], b["weights"], b["jac"])
    psfs = ng.ObservationBatch.from_arrays(p["images"], p["weights"], p["jac"])
    out = objs.bootstrap(psfs, model="exp", psf_ngauss=3, psf_ntry=2, ntry=2,
                         seed=seed)
    res = out["res"]
    # ---- stage 1: psf guesses and EM, try by try
    todo = np.ones(n, dtype=bool)
    psf_pars = np.zeros((n, 3, 6))
    psf_flags = np.zeros(n, dtype="i4")
    psf_ntry = np.zeros(n, dtype="i4")
    for tr in range(2):
        for i in np.flatnonzero(todo):
            g = psf_guess_host(seed, i, tr, 3, p["jac"][i][7],
                               device_stamp_sum(p["images"][i]))
            try:
                r, pars = ora.em_go(p["images"][i], p["weights"][i],
                                    p["jac"][i], g.reshape(3, 6))
                flags = int(r["flags"])
            except Exception:
                flags, pars = 1, np.zeros((3, 6))
            psf_ntry[i] = tr + 1
            psf_flags[i] = flags
            if flags == 0:
                assert int(out["psf_res"]["numiter"][i]) == int(r["numiter"])
                psf_pars[i] = np.asarray(pars).reshape(3, 6)
                todo[i] = False
    assert np.array_equal(out["psf_ntry"], psf_ntry)
    ok = psf_flags == 0
    # 3-gaussian EM fits of a turb psf run into maxiter on a third of the
    # guesses (flags 32), which is what the retries are for
    assert ok.sum() >= n // 2 and not ok[5] and (psf_ntry == 2).sum() > 5
    assert np.array_equal(out["psf_res"]["flags"][ok], psf_flags[ok])
    assert np.allclose(out["psf_pars"][ok], psf_pars[ok], rtol=1e-9, atol=0)
    # ---- stage 2: psf fluxes
    assert np.all(res["flags"][~ok] == 1)
    # ---- stage 3: guesses and LM fits from the device's own psf mixtures
    for i in np.flatnonzero(ok):
        tr = int(out["ntry"][i]) - 1
        T = (out["psf_pars"][i][:, 0]
             * (out["psf_pars"][i][:, 3] + out["psf_pars"][i][:, 5])).sum() \
            / out["psf_pars"][i][:, 0].sum()
        g = obj_guess_host(seed, i, tr, T, out["psf_flux"]["flux"][i])
        assert np.allclose(out["guess"][i], g, rtol=1e-13, atol=0)
        r = ora.lm_fit("exp", b["images"][i], b["weights"][i], b["jac"][i],
                       out["psf_pars"][i], out["guess"][i])
        assert int(r["flags"]) == int(res["flags"][i])
        assert int(r["nfev"]) == int(res["nfev"][i])
        if r["flags"] == 0:
            assert np.all(np.abs(res["pars"][i][:6] - r["pars"][:6])
                          <= np.maximum(1e-5 * np.abs(r["pars"][:6]),
                                        1e-3 * r["pars_err"][:6]))
    assert (res["flags"][ok] == 0).sum() >= ok.sum() - 2


@pytest.mark.gpu
def test_bootstrap_chain_is_independent_of_batching(chain_inputs):
    """the guesses depend on (seed, index, try) only: a batch split in two
    with the generator's index offset... is not offered; the SAME batch run
    twice, and with a different number of stamps behind it, gives the same
    records for the shared objects"""
    import ngmix_b200 as ng
    b, p = chain_inputs
    objs = ng.ObservationBatch.from_arrays(b["images"], b["weights"], b["jac"])
    psfs = ng.ObservationBatch.from_arrays(p["images"], p["weights"], p["jac"])
    a = objs.bootstrap(psfs, seed=5, psf_ntry=2, ntry=2)
    c = objs.bootstrap(psfs, seed=5, psf_ntry=2, ntry=2)
    for k in ("res", "psf_res", "psf_flux"):
        for name in a[k].dtype.names:
            assert np.array_equal(a[k][name], c[k][name], equal_nan=True)
    m = 40
    o2 = ng.ObservationBatch.from_arrays(b["images"][:m], b["weights"][:m],
                                         b["jac"][:m])
    p2 = ng.ObservationBatch.from_arrays(p["images"][:m], p["weights"][:m],
                                         p["jac"][:m])
    d = o2.bootstrap(p2, seed=5, psf_ntry=2, ntry=2)
    for name in d["res"].dtype.names:
        assert np.array_equal(d["res"][name], a["res"][name][:m],
                              equal_nan=True)
    e = objs.bootstrap(psfs, seed=6, psf_ntry=2, ntry=2)
    assert not np.array_equal(e["guess"], a["guess"])
