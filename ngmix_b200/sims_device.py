"""
Full-size synthetic batches for the BASELINE.json configs, built ON the
device: parameters are drawn on the host (vectorised numpy, one Generator per
batch), the truth mixtures are filled, convolved and rendered by the engine's
own device entry points (ngx_gmix_fill / ngx_gmix_convolve / ngx_render, exact
exp as GMix.make_image does) and gaussian noise is added with torch's device
generator.  Nothing here is on the fitting path; it exists because 10^5-10^6
stamps cannot be rendered by the per-object numpy loops of sims.py in bench
time.  Parity samples are copied back to the host and refit by the oracle on
identical pixel values.

The per-object recipes are those of sims.py / SURVEY.md section 8d.
"""
import numpy as np

from . import _lib
from .sims import _TABLES


def _torch():
    import torch
    return torch


def _diag_jac(n, H, W, scale, rng):
    jac = np.zeros((n, 8))
    jac[:, 0] = (H - 1) / 2 + rng.uniform(-0.5, 0.5, size=n)
    jac[:, 1] = (W - 1) / 2 + rng.uniform(-0.5, 0.5, size=n)
    jac[:, 2] = scale
    jac[:, 5] = scale
    jac[:, 6] = scale * scale
    jac[:, 7] = scale
    return jac


def _draw_g(n, rng, sigma=0.15, gmax=0.6):
    g = rng.normal(0.0, sigma, size=(n, 2))
    for _ in range(100):
        bad = np.hypot(g[:, 0], g[:, 1]) >= gmax
        if not bad.any():
            break
        g[bad] = rng.normal(0.0, sigma, size=(int(bad.sum()), 2))
    return g


def _turb_full(n, rng):
    """'turb' psf mixtures as full parameters [n, 3, 6] (vectorised
    simple_full_pars)"""
    g1 = rng.normal(0.0, 0.01, size=n)
    g2 = rng.normal(0.0, 0.01, size=n)
    T = rng.uniform(0.22, 0.32, size=n)
    return simple_full(np.stack([np.zeros(n), np.zeros(n), g1, g2, T,
                                 np.ones(n)], axis=1), "turb")


def simple_full(pars, model):
    """vectorised sims.simple_full_pars: [n, 6] -> [n, ngauss, 6]"""
    fv, pv = (np.array(x) for x in _TABLES[model])
    n = pars.shape[0]
    g = np.hypot(pars[:, 2], pars[:, 3])
    e = np.tanh(2 * np.arctanh(np.minimum(g, 0.999999)))
    fac = np.where(g > 0, e / np.maximum(g, 1e-300), 0.0)
    e1, e2 = fac * pars[:, 2], fac * pars[:, 3]
    out = np.zeros((n, fv.size, 6))
    Th = 0.5 * pars[:, 4, None] * fv[None, :]
    out[:, :, 0] = pars[:, 5, None] * pv[None, :]
    out[:, :, 1] = pars[:, 0, None]
    out[:, :, 2] = pars[:, 1, None]
    out[:, :, 3] = Th * (1 - e1[:, None])
    out[:, :, 4] = Th * e2[:, None]
    out[:, :, 5] = Th * (1 + e1[:, None])
    return out


def _sheared_moments(irr, irc, icc, s1, s2):
    """vectorised shape.get_sheared_moments (GMix.get_sheared,
    ngmix/gmix/gmix.py:478-516, moments.py:268-291)"""
    T = irr + icc
    e1, e2 = (icc - irr) / T, 2 * irc / T
    e = np.hypot(e1, e2)
    gg = np.tanh(0.5 * np.arctanh(np.minimum(e, 0.999999999)))
    fac = np.where(e > 0, gg / np.maximum(e, 1e-300), 0.0)
    g = (e1 + 1j * e2) * fac
    sh = s1 + 1j * s2
    gs = (g + sh) / (1.0 + np.conj(sh) * g)
    gsq0 = np.abs(g) ** 2
    Tround = T * (1 - gsq0) / (1 + gsq0)
    gsq = np.abs(gs) ** 2
    Ts = Tround * (1 + gsq) / (1 - gsq)
    ga = np.abs(gs)
    ee = np.tanh(2 * np.arctanh(ga))
    f2 = np.where(ga > 0, ee / np.maximum(ga, 1e-300), 0.0)
    e1s, e2s = gs.real * f2, gs.imag * f2
    return (1 - e1s) * Ts / 2.0, e2s * Ts / 2.0, (1 + e1s) * Ts / 2.0


def render_noisy(model, truth, psf_full, jac, dims, noise, device, seed,
                 chunk=65536, shared_noise_groups=1, dtype_check=True):
    """
    images [n, H, W] float64 on `device`: truth model convolved with the psf
    mixtures, rendered with the exact exp, plus N(0, noise).  With
    shared_noise_groups = k every k consecutive stamps share one noise field
    (the metacal-style batch).  Pixel values are rounded to float32 precision
    (what a survey stores), like bench.py's config-2 stamps.
    """
    torch = _torch()
    n = truth.shape[0]
    H, W = dims
    ng0 = _lib.MODEL_NGAUSS[model]
    npsf = psf_full.shape[1]
    images = torch.zeros((n, H, W), dtype=torch.float64, device=device)
    gen = torch.Generator(device=device)
    gen.manual_seed(int(seed))
    s = torch.cuda.current_stream().cuda_stream
    gsz = _lib.GAUSS_DTYPE.itemsize
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        m = hi - lo
        d_pars = torch.from_numpy(np.ascontiguousarray(truth[lo:hi])).to(device)
        d_psfp = torch.from_numpy(np.ascontiguousarray(
            psf_full[lo:hi].reshape(m, -1))).to(device)
        d_jac = torch.from_numpy(np.ascontiguousarray(jac[lo:hi])).to(device)
        d_g0 = torch.zeros(m * ng0 * gsz, dtype=torch.uint8, device=device)
        d_gp = torch.zeros(m * npsf * gsz, dtype=torch.uint8, device=device)
        d_gc = torch.zeros(m * ng0 * npsf * gsz, dtype=torch.uint8,
                           device=device)
        d_st = torch.zeros(m, dtype=torch.int32, device=device)
        _lib.gmix_fill_device(model, m, d_pars.data_ptr(), truth.shape[1],
                              ng0, d_g0.data_ptr(), d_st.data_ptr(), stream=s)
        _lib.gmix_fill_device("full", m, d_psfp.data_ptr(), 6 * npsf, npsf,
                              d_gp.data_ptr(), d_st.data_ptr(), stream=s)
        _lib.gmix_convolve_device(m, ng0, npsf, d_g0.data_ptr(),
                                  d_gp.data_ptr(), d_gc.data_ptr(), stream=s)
        _lib.render_device(m, ng0 * npsf, d_gc.data_ptr(), d_jac.data_ptr(),
                           H, W, 0, images[lo:hi].data_ptr(),
                           d_st.data_ptr(), stream=s)
        if dtype_check:
            assert int(d_st.abs().sum().item()) == 0
        k = shared_noise_groups
        if k == 1:
            nz = torch.randn((m, H, W), dtype=torch.float64, device=device,
                             generator=gen)
            images[lo:hi].add_(nz, alpha=float(noise))
        else:
            assert lo % k == 0 and m % k == 0
            nz = torch.randn((m // k, 1, H, W), dtype=torch.float64,
                             device=device, generator=gen)
            images[lo:hi].view(m // k, k, H, W).add_(nz, alpha=float(noise))
        del nz, d_g0, d_gp, d_gc
        images[lo:hi].copy_(images[lo:hi].to(torch.float32))
    return images


def make_lm_batch_device(n, seed, device, dims=(48, 48), noise=0.01,
                         scale=0.263, guess_frac=0.05):
    """config 2 recipe (sims.make_lm_batch), vectorised; returns a dict of
    device tensors + the host parameter arrays"""
    torch = _torch()
    rng = np.random.default_rng(seed)
    H, W = dims
    jac = _diag_jac(n, H, W, scale, rng)
    g = _draw_g(n, rng)
    truth = np.stack([np.zeros(n), np.zeros(n), g[:, 0], g[:, 1],
                      rng.uniform(0.15, 1.0, size=n),
                      rng.uniform(50.0, 200.0, size=n)], axis=1)
    psf = _turb_full(n, rng)
    guess = np.zeros((n, 6))
    guess[:, 0:2] = rng.uniform(-0.01, 0.01, size=(n, 2))
    guess[:, 2:] = truth[:, 2:] * (1.0 + rng.uniform(-guess_frac, guess_frac,
                                                     size=(n, 4)))
    images = render_noisy("exp", truth, psf, jac, dims, noise, device, seed)
    w32 = float(np.float32(1.0 / noise ** 2))
    weights = torch.full((n, H, W), w32, dtype=torch.float64, device=device)
    return dict(images=images, weights=weights,
                jac=torch.from_numpy(jac).to(device),
                psf=torch.from_numpy(psf).to(device),
                guess=torch.from_numpy(guess).to(device),
                h_jac=jac, h_psf=psf, h_guess=guess, truth=truth)


def make_bdf_batch_device(n, seed, device, dims=(64, 64), noise=0.02,
                          scale=0.263, guess_frac=0.05):
    """config 4 recipe (sims.make_bdf_batch), vectorised"""
    torch = _torch()
    rng = np.random.default_rng(seed)
    H, W = dims
    jac = _diag_jac(n, H, W, scale, rng)
    g = _draw_g(n, rng)
    truth = np.stack([np.zeros(n), np.zeros(n), g[:, 0], g[:, 1],
                      rng.uniform(0.15, 1.0, size=n),
                      rng.uniform(0.1, 0.9, size=n),
                      rng.uniform(50.0, 200.0, size=n)], axis=1)
    psf = _turb_full(n, rng)
    guess = np.zeros((n, 7))
    guess[:, 0:2] = rng.uniform(-0.01, 0.01, size=(n, 2))
    guess[:, 2:] = truth[:, 2:] * (1.0 + rng.uniform(-guess_frac, guess_frac,
                                                     size=(n, 5)))
    images = render_noisy("bdf", truth, psf, jac, dims, noise, device, seed,
                          chunk=32768)
    w32 = float(np.float32(1.0 / noise ** 2))
    weights = torch.full((n, H, W), w32, dtype=torch.float64, device=device)
    return dict(images=images, weights=weights,
                jac=torch.from_numpy(jac).to(device),
                psf=torch.from_numpy(psf).to(device),
                guess=torch.from_numpy(guess).to(device),
                h_jac=jac, h_psf=psf, h_guess=guess, truth=truth)


def make_metacal_batch_device(nobj, seed, device, dims=(48, 48), noise=0.01,
                              scale=0.263, guess_frac=0.05):
    """config 5 recipe (sims.make_metacal_batch), vectorised: 5 stamps per
    object (noshear, 1p, 1m, 2p, 2m), object-major, one noise field per
    object, round 1-gaussian reconvolution psf T = 0.35"""
    torch = _torch()
    from .sims import METACAL_TYPES
    rng = np.random.default_rng(seed)
    H, W = dims
    n = 5 * nobj
    j0 = _diag_jac(nobj, H, W, scale, rng)
    g = _draw_g(nobj, rng)
    truth = np.stack([np.zeros(nobj), np.zeros(nobj), g[:, 0], g[:, 1],
                      rng.uniform(0.15, 1.0, size=nobj),
                      rng.uniform(50.0, 200.0, size=nobj)], axis=1)
    g0 = np.zeros((nobj, 6))
    g0[:, 0:2] = rng.uniform(-0.01, 0.01, size=(nobj, 2))
    g0[:, 2:] = truth[:, 2:] * (1.0 + rng.uniform(-guess_frac, guess_frac,
                                                  size=(nobj, 4)))
    full0 = simple_full(truth, "exp")                  # [nobj, 6, 6]
    full = np.zeros((nobj, 5, 6, 6))
    for k, (_, s1, s2) in enumerate(METACAL_TYPES):
        full[:, k] = full0
        irr, irc, icc = _sheared_moments(full0[:, :, 3], full0[:, :, 4],
                                         full0[:, :, 5], s1, s2)
        full[:, k, :, 3], full[:, k, :, 4], full[:, k, :, 5] = irr, irc, icc
    full = full.reshape(n, 36)
    ppars = simple_full(np.array([[0.0, 0.0, 0.0, 0.0, 0.35, 1.0]]),
                        "gauss")[0]                    # [1, 6]
    psf = np.repeat(ppars[None], n, axis=0)            # [n, 1, 6]
    jac = np.repeat(j0, 5, axis=0)
    guess = np.repeat(g0, 5, axis=0)
    # the sheared truth is a general 6-gaussian mixture: render it as 'full'
    images = _render_full_noisy(full, psf, jac, dims, noise, device, seed)
    w32 = float(np.float32(1.0 / noise ** 2))
    weights = torch.full((n, H, W), w32, dtype=torch.float64, device=device)
    return dict(images=images, weights=weights,
                jac=torch.from_numpy(jac).to(device),
                psf=torch.from_numpy(psf).to(device),
                guess=torch.from_numpy(guess).to(device),
                h_jac=jac, h_psf=psf, h_guess=guess, truth=truth)


def _render_full_noisy(full, psf, jac, dims, noise, device, seed,
                       chunk=65535 * 5):
    """render_noisy for mixtures given as full parameters [n, 6 ngauss]"""
    torch = _torch()
    n = full.shape[0]
    H, W = dims
    ng0 = full.shape[1] // 6
    npsf = psf.shape[1]
    images = torch.zeros((n, H, W), dtype=torch.float64, device=device)
    gen = torch.Generator(device=device)
    gen.manual_seed(int(seed))
    s = torch.cuda.current_stream().cuda_stream
    gsz = _lib.GAUSS_DTYPE.itemsize
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        m = hi - lo
        d_pars = torch.from_numpy(np.ascontiguousarray(full[lo:hi])).to(device)
        d_psfp = torch.from_numpy(np.ascontiguousarray(
            psf[lo:hi].reshape(m, -1))).to(device)
        d_jac = torch.from_numpy(np.ascontiguousarray(jac[lo:hi])).to(device)
        d_g0 = torch.zeros(m * ng0 * gsz, dtype=torch.uint8, device=device)
        d_gp = torch.zeros(m * npsf * gsz, dtype=torch.uint8, device=device)
        d_gc = torch.zeros(m * ng0 * npsf * gsz, dtype=torch.uint8,
                           device=device)
        d_st = torch.zeros(m, dtype=torch.int32, device=device)
        _lib.gmix_fill_device("full", m, d_pars.data_ptr(), 6 * ng0, ng0,
                              d_g0.data_ptr(), d_st.data_ptr(), stream=s)
        _lib.gmix_fill_device("full", m, d_psfp.data_ptr(), 6 * npsf, npsf,
                              d_gp.data_ptr(), d_st.data_ptr(), stream=s)
        _lib.gmix_convolve_device(m, ng0, npsf, d_g0.data_ptr(),
                                  d_gp.data_ptr(), d_gc.data_ptr(), stream=s)
        _lib.render_device(m, ng0 * npsf, d_gc.data_ptr(), d_jac.data_ptr(),
                           H, W, 0, images[lo:hi].data_ptr(),
                           d_st.data_ptr(), stream=s)
        nz = torch.randn((m // 5, 1, H, W), dtype=torch.float64,
                         device=device, generator=gen)
        images[lo:hi].view(m // 5, 5, H, W).add_(nz, alpha=float(noise))
        del nz
        images[lo:hi].copy_(images[lo:hi].to(torch.float32))
    return images


def make_psf_batch_device(n, seed, device, dims=(25, 25), noise=1.0e-4,
                          scale=0.263):
    """config 3 recipe (sims.make_psf_batch), vectorised: 'turb' psf stamps,
    EM 3-gaussian guesses and 1-gaussian admom guesses"""
    torch = _torch()
    from .sims import _EM3_P, _EM3_F, fwhm_to_T
    rng = np.random.default_rng(seed)
    H, W = dims
    jac = _diag_jac(n, H, W, scale, rng)
    truth = np.stack([np.zeros(n), np.zeros(n),
                      rng.normal(0.0, 0.02, size=n),
                      rng.normal(0.0, 0.02, size=n),
                      rng.uniform(0.22, 0.32, size=n), np.ones(n)], axis=1)
    unit_psf = np.zeros((n, 1, 6))
    # no further convolution: a delta-like 'psf' would change the truth, so
    # the turb mixture is rendered directly
    full = simple_full(truth, "turb").reshape(n, 18)
    images = _render_plain_noisy(full, jac, dims, noise, device, seed)
    del unit_psf
    flux = images.sum(dim=(1, 2)).cpu().numpy()
    sigma2 = fwhm_to_T(scale * 3.5) / 2
    em_guess = np.zeros((n, 3, 6))
    for k in range(3):
        em_guess[:, k, 0] = flux * _EM3_P[k] * (1 + rng.uniform(-0.1, 0.1, n))
        em_guess[:, k, 1] = rng.uniform(-0.1, 0.1, n)
        em_guess[:, k, 2] = rng.uniform(-0.1, 0.1, n)
        em_guess[:, k, 3] = _EM3_F[k] * sigma2 * (1 + rng.uniform(-0.1, 0.1, n))
        em_guess[:, k, 4] = rng.uniform(-0.01, 0.01, n)
        em_guess[:, k, 5] = _EM3_F[k] * sigma2 * (1 + rng.uniform(-0.1, 0.1, n))
    am = np.zeros((n, 6))
    am[:, 0:2] = rng.uniform(-0.5 * scale, 0.5 * scale, size=(n, 2))
    am[:, 2:4] = rng.uniform(-0.3, 0.3, size=(n, 2))
    am[:, 4] = fwhm_to_T(scale * 3.5) * (1 + rng.uniform(-0.1, 0.1, n))
    am[:, 5] = 1.0
    am_guess = simple_full(am, "gauss")[:, 0, :]
    w = float(1.0 / noise ** 2)
    weights = torch.full((n, H, W), w, dtype=torch.float64, device=device)
    return dict(images=images, weights=weights,
                jac=torch.from_numpy(jac).to(device),
                em_guess=torch.from_numpy(em_guess).to(device),
                am_guess=torch.from_numpy(np.ascontiguousarray(am_guess)).to(device),
                h_jac=jac, h_em_guess=em_guess, h_am_guess=am_guess,
                truth=truth)


def _render_plain_noisy(full, jac, dims, noise, device, seed):
    torch = _torch()
    n = full.shape[0]
    H, W = dims
    ng = full.shape[1] // 6
    images = torch.zeros((n, H, W), dtype=torch.float64, device=device)
    gen = torch.Generator(device=device)
    gen.manual_seed(int(seed))
    s = torch.cuda.current_stream().cuda_stream
    gsz = _lib.GAUSS_DTYPE.itemsize
    d_pars = torch.from_numpy(np.ascontiguousarray(full)).to(device)
    d_jac = torch.from_numpy(np.ascontiguousarray(jac)).to(device)
    d_g = torch.zeros(n * ng * gsz, dtype=torch.uint8, device=device)
    d_st = torch.zeros(n, dtype=torch.int32, device=device)
    _lib.gmix_fill_device("full", n, d_pars.data_ptr(), 6 * ng, ng,
                          d_g.data_ptr(), d_st.data_ptr(), stream=s)
    _lib.render_device(n, ng, d_g.data_ptr(), d_jac.data_ptr(), H, W, 0,
                       images.data_ptr(), d_st.data_ptr(), stream=s)
    images.add_(torch.randn((n, H, W), dtype=torch.float64, device=device,
                            generator=gen), alpha=float(noise))
    return images
