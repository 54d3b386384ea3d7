"""
Synthetic stamp batches for the BASELINE.json configs.

Everything here is INPUT synthesis (numpy, host): it is used by the golden
fixture generator (oracle/gen_golden.py), by the tests and by bench.py to
build seeded stamps.  It is not on the fitting path.  Per-object seeds
(RandomState(seed0 + index)) make every object independent of how the batch
is sharded over ranks.

Layouts produced (all float64, C order) are the on-wire batch format of the
engine ("Data layout"):

  images  [N, H, W]    weights [N, H, W]
  jac     [N, 8]       row0, col0, dvdrow, dvdcol, dudrow, dudcol, det, scale
                       (the reference's _jacobian_dtype record,
                        ngmix/jacobian/jacobian.py:406-414)
  psf     [N, npsf, 6] full gaussian pars p,row,col,irr,irc,icc
  guess   [N, npars]
"""
import numpy as np

# mixture tables of the simple models: (fvals, pvals); values are the data
# constants of ngmix/gmix/gmix_nb.py:243-304
_TABLES = {
    "exp": (
        [0.002467115141477932, 0.018147435573256168, 0.07944063151366336,
         0.27137669897479122, 0.79782256866993773, 2.1623306025075739],
        [0.00061601229677880041, 0.0079461395724623237, 0.053280454055540001,
         0.21797364640726541, 0.45496740582554868, 0.26521634184240478],
    ),
    "dev": (
        [2.9934935706271918e-07, 3.4651596338231207e-06,
         2.4807910570562753e-05, 1.4307404300535354e-04,
         7.2753169298239500e-04, 3.4582464394427260e-03,
         1.6086645440719100e-02, 7.7006776775654429e-02,
         4.1012562102501476e-01, 2.9812509778548648e00],
        [6.5288960012625658e-05, 0.00044199216814302695,
         0.0020859587871659754, 0.0075913681418996841, 0.02260266219257237,
         0.056532254390212859, 0.11939049233042602, 0.20969545753234975,
         0.29254151133139222, 0.28905301416582552],
    ),
    "turb": (
        [0.5793612389470884, 1.621860687127999, 7.019347162356363],
        [0.596510042804182, 0.4034898268889178, 1.303069003078001e-07],
    ),
    "gauss": ([1.0], [1.0]),
}


def g1g2_to_e1e2(g1, g2):
    """shear-style g to distortion-style e, e = tanh(2 atanh g)"""
    g = np.sqrt(g1 * g1 + g2 * g2)
    if g >= 1.0:
        raise ValueError("g >= 1")
    if g == 0.0:
        return 0.0, 0.0
    e = np.tanh(2 * np.arctanh(g))
    if e >= 1.0:
        e = 0.99999999
    fac = e / g
    return fac * g1, fac * g2


def simple_full_pars(pars, model):
    """[cen1,cen2,g1,g2,T,F] of a simple model -> [ngauss,6] full pars"""
    fvals, pvals = _TABLES[model]
    row, col, g1, g2, T, flux = pars
    e1, e2 = g1g2_to_e1e2(g1, g2)
    out = np.zeros((len(fvals), 6))
    for i, (f, p) in enumerate(zip(fvals, pvals)):
        T_i_2 = 0.5 * T * f
        out[i] = (flux * p, row, col,
                  T_i_2 * (1 - e1), T_i_2 * e2, T_i_2 * (1 + e1))
    return out


def bdf_full_pars(pars):
    """[cen1,cen2,g1,g2,T,fracdev,F] -> [16,6] (ngmix/gmix/gmix_nb.py:515)"""
    row, col, g1, g2, T, fracdev, flux = pars
    fe, pe = _TABLES["exp"]
    fd, pd = _TABLES["dev"]
    ifracdev = 1.0 - fracdev
    Tfactor = 0.0
    for i in range(6):
        Tfactor += (pe[i] * ifracdev) * fe[i]
    for i in range(10):
        Tfactor += (pd[i] * fracdev) * (fd[i] * 1.0)
    Tfactor = 1.0 / Tfactor
    T = T * Tfactor
    e1, e2 = g1g2_to_e1e2(g1, g2)
    out = np.zeros((16, 6))
    for i in range(16):
        if i < 6:
            p = pe[i] * ifracdev
            f = fe[i]
        else:
            p = pd[i - 6] * fracdev
            f = fd[i - 6] * 1.0
        T_i_2 = 0.5 * T * f
        out[i] = (flux * p, row, col,
                  T_i_2 * (1 - e1), T_i_2 * e2, T_i_2 * (1 + e1))
    return out


def convolve_full_pars(obj, psf):
    """object-major convolution of two [n,6] full-par sets"""
    psum = psf[:, 0].sum()
    prow = (psf[:, 0] * psf[:, 1]).sum() / psum
    pcol = (psf[:, 0] * psf[:, 2]).sum() / psum
    out = np.zeros((obj.shape[0] * psf.shape[0], 6))
    k = 0
    for o in obj:
        for p in psf:
            out[k] = (o[0] * p[0] / psum,
                      o[1] + (p[1] - prow), o[2] + (p[2] - pcol),
                      o[3] + p[3], o[4] + p[4], o[5] + p[5])
            k += 1
    return out


def render_full_pars(full, dims, jac):
    """
    exact-exp render of [ngauss,6] full pars on a (H,W) grid with an 8-element
    jacobian record; host numpy, for input synthesis only
    """
    H, W = dims
    row0, col0, dvdrow, dvdcol, dudrow, dudcol, det, scale = jac
    rows, cols = np.mgrid[0:H, 0:W]
    rd = rows - row0
    cd = cols - col0
    v = dvdrow * rd + dvdcol * cd
    u = dudrow * rd + dudcol * cd
    area = scale * scale
    im = np.zeros((H, W))
    for p, r, c, irr, irc, icc in full:
        d = irr * icc - irc * irc
        dv = v - r
        du = u - c
        chi2 = (icc / d) * dv * dv + (irr / d) * du * du \
            - 2.0 * (irc / d) * dv * du
        im += (p / (2 * np.pi * np.sqrt(d))) * np.exp(-0.5 * chi2) * area
    return im


def diag_jacobian(row, col, scale):
    """8-element record of a DiagonalJacobian (jacobian.py:332-377)"""
    det = scale * scale
    return np.array([row, col, scale, 0.0, 0.0, scale, det,
                     np.sqrt(abs(det))])


def make_lm_batch(n, seed0=0, dims=(48, 48), model="exp", psf_model="turb",
                  noise=0.01, scale=0.263, guess_frac=0.05,
                  truth_images=True):
    """
    Config 1/2/5 style batch: exp (or gauss/dev) galaxies
    convolved with a 1- or 3-gaussian PSF, gaussian noise, guesses within
    guess_frac of the truth.

    truth_images=False skips the (slow, host) rendering and returns zero
    images; bench.py then renders on the device through the engine itself.
    """
    H, W = dims
    npsf = len(_TABLES[psf_model][0])
    images = np.zeros((n, H, W))
    weights = np.zeros((n, H, W)) + 1.0 / noise ** 2
    jac = np.zeros((n, 8))
    psf = np.zeros((n, npsf, 6))
    truth = np.zeros((n, 6))
    guess = np.zeros((n, 6))
    noise_seeds = np.zeros(n, dtype="i8")
    for i in range(n):
        rng = np.random.RandomState(seed0 + i)
        dy, dx = rng.uniform(-0.5, 0.5, size=2)
        jac[i] = diag_jacobian((H - 1) / 2 + dy, (W - 1) / 2 + dx, scale)
        while True:
            g1, g2 = rng.normal(0.0, 0.15, size=2)
            if np.hypot(g1, g2) < 0.6:
                break
        T = rng.uniform(0.15, 1.0)
        flux = rng.uniform(50.0, 200.0)
        truth[i] = (0.0, 0.0, g1, g2, T, flux)
        if psf_model == "gauss":
            ppars = [0.0, 0.0, 0.01, -0.01, 0.27, 1.0]
        else:
            ppars = [0.0, 0.0, rng.normal(0, 0.01), rng.normal(0, 0.01),
                     rng.uniform(0.22, 0.32), 1.0]
        psf[i] = simple_full_pars(ppars, psf_model)
        guess[i, 0:2] = rng.uniform(-0.01, 0.01, size=2)
        guess[i, 2:] = truth[i, 2:] * (
            1.0 + rng.uniform(-guess_frac, guess_frac, size=4))
        noise_seeds[i] = rng.randint(0, 2 ** 31 - 1)
        if truth_images:
            full = convolve_full_pars(
                simple_full_pars(truth[i], model), psf[i])
            im = render_full_pars(full, dims, jac[i])
            nrng = np.random.RandomState(noise_seeds[i])
            images[i] = im + nrng.normal(0.0, noise, size=im.shape)
    return dict(images=images, weights=weights, jac=jac, psf=psf,
                truth=truth, guess=guess, noise_seeds=noise_seeds,
                model=model, noise=noise)


# EM guess tables for 3 gaussians, the 'turb' decomposition also used by
# GMixPSFGuesser (ngmix/guessers.py:1027-1030)
_EM3_P = np.array([0.596510042804182, 0.4034898268889178,
                   1.303069003078001e-07])
_EM3_F = np.array([0.5793612389470884, 1.621860687127999, 7.019347162356363])


def fwhm_to_T(fwhm):
    sigma = fwhm / 2.3548200450309493
    return 2 * sigma ** 2


def make_psf_batch(n, seed0=0, dims=(25, 25), noise=1.0e-4, scale=0.263,
                   truth_images=True):
    """
    Config 3 batch: 'turb' PSF stamps with EM 3-gaussian
    guesses (the GMixPSFGuesser ngauss=3 recipe, guessers.py:921-949) and
    1-gaussian admom guesses (the admom.py:392-403 recipe).
    """
    H, W = dims
    images = np.zeros((n, H, W))
    weights = np.zeros((n, H, W)) + 1.0 / noise ** 2
    jac = np.zeros((n, 8))
    truth = np.zeros((n, 6))
    em_guess = np.zeros((n, 3, 6))
    am_guess = np.zeros((n, 6))
    for i in range(n):
        rng = np.random.RandomState(seed0 + i)
        dy, dx = rng.uniform(-0.5, 0.5, size=2)
        jac[i] = diag_jacobian((H - 1) / 2 + dy, (W - 1) / 2 + dx, scale)
        g1, g2 = rng.normal(0.0, 0.02, size=2)
        T = rng.uniform(0.22, 0.32)
        truth[i] = (0.0, 0.0, g1, g2, T, 1.0)
        if truth_images:
            im = render_full_pars(
                simple_full_pars(truth[i], "turb"), dims, jac[i])
            images[i] = im + rng.normal(0.0, noise, size=im.shape)
        flux = images[i].sum() if truth_images else 1.0
        sigma2 = fwhm_to_T(scale * 3.5) / 2
        for k in range(3):
            em_guess[i, k] = (
                flux * _EM3_P[k] * (1.0 + rng.uniform(-0.1, 0.1)),
                rng.uniform(-0.1, 0.1),
                rng.uniform(-0.1, 0.1),
                _EM3_F[k] * sigma2 * (1.0 + rng.uniform(-0.1, 0.1)),
                rng.uniform(-0.01, 0.01),
                _EM3_F[k] * sigma2 * (1.0 + rng.uniform(-0.1, 0.1)),
            )
        am_pars = np.zeros(6)
        am_pars[0:2] = rng.uniform(-0.5 * scale, 0.5 * scale, size=2)
        am_pars[2:4] = rng.uniform(-0.3, 0.3, size=2)
        am_pars[4] = fwhm_to_T(scale * 3.5) * (1.0 + rng.uniform(-0.1, 0.1))
        am_pars[5] = 1.0
        am_guess[i] = simple_full_pars(am_pars, "gauss")[0]
    return dict(images=images, weights=weights, jac=jac, truth=truth,
                em_guess=em_guess, am_guess=am_guess, noise=noise)


def make_bdf_batch(n, seed0=0, dims=(64, 64), noise=0.01, scale=0.263,
                   guess_frac=0.05, truth_images=True):
    """Config 4 batch: bdf truth, 3-gaussian PSF (48 composed gaussians)"""
    H, W = dims
    images = np.zeros((n, H, W))
    weights = np.zeros((n, H, W)) + 1.0 / noise ** 2
    jac = np.zeros((n, 8))
    psf = np.zeros((n, 3, 6))
    truth = np.zeros((n, 7))
    guess = np.zeros((n, 7))
    for i in range(n):
        rng = np.random.RandomState(seed0 + i)
        dy, dx = rng.uniform(-0.5, 0.5, size=2)
        jac[i] = diag_jacobian((H - 1) / 2 + dy, (W - 1) / 2 + dx, scale)
        while True:
            g1, g2 = rng.normal(0.0, 0.15, size=2)
            if np.hypot(g1, g2) < 0.6:
                break
        T = rng.uniform(0.15, 1.0)
        fracdev = rng.uniform(0.1, 0.9)
        flux = rng.uniform(50.0, 200.0)
        truth[i] = (0.0, 0.0, g1, g2, T, fracdev, flux)
        ppars = [0.0, 0.0, rng.normal(0, 0.01), rng.normal(0, 0.01),
                 rng.uniform(0.22, 0.32), 1.0]
        psf[i] = simple_full_pars(ppars, "turb")
        guess[i, 0:2] = rng.uniform(-0.01, 0.01, size=2)
        guess[i, 2:] = truth[i, 2:] * (
            1.0 + rng.uniform(-guess_frac, guess_frac, size=5))
        if truth_images:
            full = convolve_full_pars(bdf_full_pars(truth[i]), psf[i])
            im = render_full_pars(full, dims, jac[i])
            images[i] = im + rng.normal(0.0, noise, size=im.shape)
    return dict(images=images, weights=weights, jac=jac, psf=psf,
                truth=truth, guess=guess, model="bdf", noise=noise)


def _shear_full_pars(full, s1, s2):
    """apply reduced shear (s1, s2) to every gaussian of [n,6] full pars, as
    GMix.get_sheared does (ngmix/gmix/gmix.py:478-516)"""
    from .shape import get_sheared_moments
    out = full.copy()
    for k in range(full.shape[0]):
        out[k, 3], out[k, 4], out[k, 5] = get_sheared_moments(
            full[k, 3], full[k, 4], full[k, 5], s1, s2)
    return out


METACAL_TYPES = (("noshear", 0.0, 0.0), ("1p", 0.01, 0.0), ("1m", -0.01, 0.0),
                 ("2p", 0.0, 0.01), ("2m", 0.0, -0.01))


def make_metacal_batch(nobj, seed0=0, dims=(48, 48), noise=0.01, scale=0.263,
                       guess_frac=0.05):
    """
    Config 5 batch: per object 5 stamps (noshear, 1p, 1m, 2p,
    2m) - the exp truth sheared by +-0.01, convolved with a round 1-gaussian
    reconvolution PSF (T = 0.35), the SAME noise field on all five - to be
    fit as 5*nobj independent exp LM fits.  Stamp order is object-major.
    """
    H, W = dims
    n = 5 * nobj
    images = np.zeros((n, H, W))
    weights = np.zeros((n, H, W)) + 1.0 / noise ** 2
    jac = np.zeros((n, 8))
    psf = np.zeros((n, 1, 6))
    guess = np.zeros((n, 6))
    truth = np.zeros((nobj, 6))
    for i in range(nobj):
        rng = np.random.RandomState(seed0 + i)
        dy, dx = rng.uniform(-0.5, 0.5, size=2)
        j = diag_jacobian((H - 1) / 2 + dy, (W - 1) / 2 + dx, scale)
        while True:
            g1, g2 = rng.normal(0.0, 0.15, size=2)
            if np.hypot(g1, g2) < 0.6:
                break
        truth[i] = (0.0, 0.0, g1, g2, rng.uniform(0.15, 1.0),
                    rng.uniform(50.0, 200.0))
        ppars = simple_full_pars([0.0, 0.0, 0.0, 0.0, 0.35, 1.0], "gauss")
        g0 = np.zeros(6)
        g0[0:2] = rng.uniform(-0.01, 0.01, size=2)
        g0[2:] = truth[i, 2:] * (1.0 + rng.uniform(-guess_frac, guess_frac,
                                                   size=4))
        noise_im = rng.normal(0.0, noise, size=(H, W))
        full0 = simple_full_pars(truth[i], "exp")
        for k, (_, s1, s2) in enumerate(METACAL_TYPES):
            full = convolve_full_pars(_shear_full_pars(full0, s1, s2), ppars)
            m = 5 * i + k
            images[m] = render_full_pars(full, dims, j) + noise_im
            jac[m] = j
            psf[m] = ppars
            guess[m] = g0
    return dict(images=images, weights=weights, jac=jac, psf=psf,
                guess=guess, truth=truth, model="exp", noise=noise)
