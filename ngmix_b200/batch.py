"""
ObservationBatch: N same-size stamps resident on the GPU.

The reference builds one Observation per stamp on the host and recomputes its
pixel array at every mutation (ngmix/observation.py:814-830,
ngmix/pixels/pixels.py:6-52); at the kernel rates of this engine that host
work, and the host->device copy of the stamps, is what bounds a pipeline
(SURVEY.md 8f rank 1).  An ObservationBatch uploads the stacked arrays once
and every fitter then runs on the resident copy: PSF EM -> PSF mixture ->
galaxy LM fit chain without the stamps or the PSF mixtures leaving the
device.  Pixel construction (v, u, ierr, zero-weight handling; pixels_nb.py)
happens inside the kernels from (image, weight, jacobian).

Device memory is held in torch tensors (plumbing only); all arithmetic goes
through the C ABI's device-pointer entry points (include/ngx.h).
"""
import numpy as np

from . import _lib
from .defaults import DEFAULT_LM_PARS
from .gmix import get_model_name, get_model_num, get_model_npars


def _torch():
    import torch
    return torch


class ObservationBatch(object):
    """
    N stamps [n, H, W] with weights, jacobians and (optionally) psf mixtures
    on one device.

    Use from_arrays or from_observations to construct.
    """

    def __init__(self, images, weights, jac, psf_pars=None, device=0,
                 ignore_zero_weight=True):
        torch = _torch()
        _lib.init(device)
        self.device = torch.device("cuda", device)
        self.ignore_zero_weight = bool(ignore_zero_weight)
        self.images = self._up(images)
        self.weights = self._up(weights)
        if self.images.dim() != 3 or self.images.shape != self.weights.shape:
            raise ValueError("images and weights must both be [n, H, W]")
        self.n, self.nrow, self.ncol = (int(x) for x in self.images.shape)
        self.jac = self._up(_lib.as_jac(jac, self.n) if not self._is_tensor(jac)
                            else jac)
        self.psf_pars = None
        self.npsf = 0
        if psf_pars is not None:
            self.set_psf(psf_pars)

    # ------------------------------------------------------------ building
    @classmethod
    def from_arrays(cls, images, weights, jacobians, psf_pars=None, device=0,
                    ignore_zero_weight=True, mask=None):
        """
        images, weights: [n, H, W] (numpy, pinned or pageable, or torch)
        jacobians: [n, 8] rows (row0, col0, dvdrow, dvdcol, dudrow, dudcol,
            det, scale) or Jacobian records
        psf_pars: [n, npsf, 6] full gaussian parameters, optional
        mask: [n, H, W] bool, optional (True = use the pixel)

        Compact survey format: float32 `images` with ONE weight per stamp
        (`weights` a scalar or an [n] array) and an optional mask are uploaded
        as they are (4.1 bytes per pixel) and expanded on the device to the
        planes the reference's Observation would hold: image.astype(f8) and
        where(mask, weight, 0) (observation.py:114-126, pixels/pixels.py:6-52).
        """
        w_arr = None if _torch().is_tensor(weights) else np.asarray(weights)
        if w_arr is not None and w_arr.ndim <= 1:
            return cls._from_compact(images, w_arr, jacobians, psf_pars,
                                     device, ignore_zero_weight, mask)
        if mask is not None:
            weights = np.where(np.asarray(mask, dtype=bool), w_arr, 0.0)
        return cls(images, weights, jacobians, psf_pars=psf_pars,
                   device=device, ignore_zero_weight=ignore_zero_weight)

    @classmethod
    def _from_compact(cls, images, w_arr, jacobians, psf_pars, device,
                      ignore_zero_weight, mask):
        torch = _torch()
        _lib.init(device)
        dev = torch.device("cuda", device)
        images = np.ascontiguousarray(images, dtype=np.float32)
        if images.ndim != 3:
            raise ValueError("images must be [n, H, W]")
        n, H, W = images.shape
        ws = np.array(np.broadcast_to(w_arr.astype("f8"), (n,)))
        d_im32 = torch.from_numpy(images).to(dev, non_blocking=True)
        d_ws = torch.from_numpy(ws).to(dev, non_blocking=True)
        d_mask = None
        if mask is not None:
            mask = np.asarray(mask, dtype=bool)
            if mask.shape != images.shape:
                raise ValueError("mask and images differ in shape")
            d_mask = torch.from_numpy(
                np.packbits(mask.reshape(n, -1), axis=1)).to(dev)
        d_im = torch.empty((n, H, W), dtype=torch.float64, device=dev)
        d_wt = torch.empty((n, H, W), dtype=torch.float64, device=dev)
        if n > 0:
            _lib.expand_compact_device(
                n, H, W, d_im32.data_ptr(), d_ws.data_ptr(),
                d_mask.data_ptr() if d_mask is not None else 0,
                d_im.data_ptr(), d_wt.data_ptr(),
                stream=torch.cuda.current_stream(dev).cuda_stream)
        return cls(d_im, d_wt, jacobians, psf_pars=psf_pars, device=device,
                   ignore_zero_weight=ignore_zero_weight)

    @classmethod
    def from_observations(cls, obslist, device=0):
        """stack same-size Observations; psf mixtures are taken from
        obs.psf.gmix when every observation has one"""
        if len(obslist) == 0:
            raise ValueError("empty observation list")
        images = np.stack([o.image for o in obslist])
        weights = np.stack([o.weight for o in obslist])
        jac = np.stack([np.asarray(o.jacobian._data.view("f8")).reshape(8)
                        for o in obslist])
        psf = None
        if all(o.has_psf_gmix() for o in obslist):
            psf = np.stack([o.psf.gmix.get_full_pars().reshape(-1, 6)
                            for o in obslist])
        izw = all(getattr(o, "_ignore_zero_weight", True) for o in obslist)
        return cls(images, weights, jac, psf_pars=psf, device=device,
                   ignore_zero_weight=izw)

    def _is_tensor(self, a):
        return _torch().is_tensor(a)

    def _up(self, a):
        torch = _torch()
        if torch.is_tensor(a):
            return a.to(device=self.device, dtype=torch.float64).contiguous()
        a = np.ascontiguousarray(a, dtype="f8")
        return torch.from_numpy(a).to(self.device, non_blocking=True)

    def set_psf(self, psf_pars):
        """psf mixtures as full parameters [n, npsf, 6]; a device tensor (for
        example the output of em()) is adopted without a copy"""
        t = self._up(psf_pars).reshape(self.n, -1, 6)
        self.psf_pars = t
        self.npsf = int(t.shape[1])

    def __len__(self):
        return self.n

    @property
    def dims(self):
        return self.nrow, self.ncol

    def _stream(self):
        return _torch().cuda.current_stream(self.device).cuda_stream

    def _empty_records(self, dtype):
        torch = _torch()
        return torch.zeros(max(self.n, 1) * dtype.itemsize, dtype=torch.uint8,
                           device=self.device)

    def _records(self, t, dtype):
        return np.frombuffer(t.cpu().numpy().tobytes(),
                             dtype=dtype)[:self.n].copy()

    # ------------------------------------------------------------- fitters
    def fit(self, model, guesses, prior=None, fit_pars=None,
            analytic_jacobian=True, return_device=False):
        """
        Levenberg-Marquardt fit of every stamp (Fitter.go,
        fitting/fitters.py:64-112) on the resident data.  guesses [n, npars].
        Returns the LM_RESULT_DTYPE records (numpy), or the raw device buffer
        with return_device=True.
        """
        from .fitting import _prior_record, _LM_MODELS
        model = get_model_name(get_model_num(model))
        if model not in _LM_MODELS:
            raise ValueError("model '%s' cannot be fit with ngx_lm_fit" % model)
        np_ = get_model_npars(model)
        g = self._up(guesses).reshape(self.n, np_)
        fp = dict(DEFAULT_LM_PARS)
        if fit_pars is not None:
            fp.update(fit_pars)
        out = self._empty_records(_lib.LM_RESULT_DTYPE)
        if self.n > 0:
            _lib.lm_fit_device(
                model, self.n, self.images.data_ptr(), self.weights.data_ptr(),
                self.jac.data_ptr(), self.nrow, self.ncol,
                self.psf_pars.data_ptr() if self.npsf else 0, self.npsf,
                g.data_ptr(), out.data_ptr(), prior=_prior_record(prior),
                maxfev=fp["maxfev"], ftol=fp["ftol"], xtol=fp["xtol"],
                analytic_jacobian=analytic_jacobian,
                ignore_zero_weight=self.ignore_zero_weight,
                stream=self._stream())
        if return_device:
            return out
        return self._records(out, _lib.LM_RESULT_DTYPE)

    def em(self, guesses, use_psf=False, variant="full", tol=1.0e-5,
           miniter=None, maxiter=500, vary_sky=False, sky=None,
           return_device=False):
        """
        EM fit of every stamp (EMFitter.go, em/em.py:228-317); guesses
        [n, ngauss, 6] full parameters.  Returns (records, pars [n,ngauss,6]);
        with return_device=True pars stays a device tensor, ready for
        set_psf() of another batch.
        """
        torch = _torch()
        g = self._up(guesses).reshape(self.n, -1, 6)
        ngauss = int(g.shape[1])
        if miniter is None:
            miniter = 20 if variant == "fluxonly" else 40
        pars = torch.zeros((self.n, ngauss, 6), dtype=torch.float64,
                           device=self.device)
        out = self._empty_records(_lib.EM_RESULT_DTYPE)
        npsf = self.npsf if use_psf else 0
        if self.n > 0:
            _lib.em_device(
                self.n, self.images.data_ptr(), self.weights.data_ptr(),
                self.jac.data_ptr(), self.nrow, self.ncol, g.data_ptr(),
                ngauss, pars.data_ptr(), out.data_ptr(),
                d_psf=self.psf_pars.data_ptr() if npsf else 0, npsf=npsf,
                variant=variant, tol=tol, miniter=miniter, maxiter=maxiter,
                vary_sky=vary_sky, sky=sky,
                ignore_zero_weight=self.ignore_zero_weight,
                stream=self._stream())
        res = self._records(out, _lib.EM_RESULT_DTYPE)
        if return_device:
            return res, pars
        return res, pars.cpu().numpy()

    def admom(self, guesses, maxiter=200, shiftmax=5.0, etol=1.0e-5,
              Ttol=1.0e-3, cenonly=False):
        """adaptive moments of every stamp (admom/admom.py:325-360); guesses
        [n, 6] full parameters of the weight gaussian"""
        g = self._up(guesses).reshape(self.n, 6)
        out = self._empty_records(_lib.ADMOM_RESULT_DTYPE)
        if self.n > 0:
            _lib.admom_device(
                self.n, self.images.data_ptr(), self.weights.data_ptr(),
                self.jac.data_ptr(), self.nrow, self.ncol, g.data_ptr(),
                out.data_ptr(), maxiter=maxiter, shiftmax=shiftmax, etol=etol,
                Ttol=Ttol, cenonly=cenonly, stream=self._stream())
        return self._records(out, _lib.ADMOM_RESULT_DTYPE)

    def gaussmom(self, fwhm, with_higher_order=False, return_records=False):
        """
        Gaussian-weighted moments of every stamp (GaussMom.go,
        gaussmom.py:49-73) on the resident data: a list of result dicts, or
        the raw MOM_RESULT_DTYPE records with return_records=True.
        """
        torch = _torch()
        from .gaussmom import GaussMom, _to_flux_units
        from .gmix import get_weighted_moments_stats, moments_record
        gm = GaussMom(fwhm, with_higher_order).weight
        nmom = 17 if with_higher_order else 6
        wt = np.repeat(gm.get_data()[None, :], max(self.n, 1), axis=0)
        d_wt = torch.from_numpy(
            np.frombuffer(wt.tobytes(), dtype=np.uint8).copy()).to(self.device)
        d_maxrad = torch.full((max(self.n, 1),),
                              100.0 * float(np.sqrt(gm.get_T() / 2.0)),
                              dtype=torch.float64, device=self.device)
        out = self._empty_records(_lib.MOM_RESULT_DTYPE)
        if self.n > 0:
            _lib.weighted_sums_device(
                self.n, 1, d_wt.data_ptr(), self.images.data_ptr(),
                self.weights.data_ptr(), self.jac.data_ptr(), self.nrow,
                self.ncol, d_maxrad.data_ptr(), nmom, out.data_ptr(),
                ignore_zero_weight=self.ignore_zero_weight,
                stream=self._stream())
        recs = self._records(out, _lib.MOM_RESULT_DTYPE)
        if return_records:
            return recs
        scale = self.jac[:, 7].cpu().numpy()
        res = []
        for i in range(self.n):
            r = get_weighted_moments_stats(moments_record(recs[i], nmom))
            if r["flags"] == 0:
                _to_flux_units(r, 1.0 / scale[i] ** 2)
            res.append(r)
        return res

    # ------------------------------------------------------ fitting chain
    def bootstrap(self, psf_batch, model="exp", psf_ngauss=3, psf_ntry=4,
                  ntry=4, T_guess=None, prior=None, fit_pars=None, seed=0,
                  em_pars=None, analytic_jacobian=True):
        """
        The production chain in one device pass (ngx_bootstrap): for every
        object an EM fit of its psf stamp (`psf_batch`, same length) from
        GMixPSFGuesser-style guesses, the psf-template flux of its image, a
        TPSFFluxGuesser-style guess and the LM fit, each stage retried with a
        fresh guess while flags != 0 (bootstrap.py:43-102, runners.py:116-149
        of the reference).  Guesses come from a counter-based generator keyed
        by (seed, object index, try): results do not depend on batch size or
        sharding when `index0`-shifted seeds are not used.

        Returns a dict: res (LM records), psf_pars [n, psf_ngauss, 6],
        psf_res (EM records), psf_flux (flux records), guess [n, npars] (the
        last guess used), psf_ntry / ntry [n] (tries used).  The psf mixtures
        also become this batch's psf (set_psf) without leaving the device.
        """
        torch = _torch()
        from .fitting import _prior_record, _LM_MODELS
        model = get_model_name(get_model_num(model))
        if model not in _LM_MODELS:
            raise ValueError("model '%s' cannot be fit with ngx_lm_fit" % model)
        if len(psf_batch) != self.n:
            raise ValueError("one psf stamp per object is needed")
        np_ = get_model_npars(model)
        fp = dict(DEFAULT_LM_PARS)
        if fit_pars is not None:
            fp.update(fit_pars)
        ep = dict(tol=1.0e-5, miniter=40, maxiter=500)
        if em_pars is not None:
            ep.update(em_pars)
        n = self.n
        dev = self.device
        psf_pars = torch.zeros((max(n, 1), psf_ngauss, 6),
                               dtype=torch.float64, device=dev)
        guess = torch.zeros((max(n, 1), np_), dtype=torch.float64, device=dev)
        ntry_used = torch.zeros((2, max(n, 1)), dtype=torch.int32, device=dev)
        d_res = self._empty_records(_lib.LM_RESULT_DTYPE)
        d_pres = self._empty_records(_lib.EM_RESULT_DTYPE)
        d_pflux = self._empty_records(_lib.FLUX_RESULT_DTYPE)
        wbytes = _lib.bootstrap_workspace_bytes(max(n, 1), psf_ngauss)
        work = torch.zeros(wbytes, dtype=torch.uint8, device=dev)
        d_T = 0
        if T_guess is not None:
            tg = self._up(np.broadcast_to(np.asarray(T_guess, dtype="f8"),
                                          (n,)).copy())
            d_T = tg.data_ptr()
        if n > 0:
            _lib.bootstrap_device(
                n, psf_batch.images.data_ptr(), psf_batch.weights.data_ptr(),
                psf_batch.jac.data_ptr(), psf_batch.nrow, psf_batch.ncol,
                psf_ngauss, psf_ntry, self.images.data_ptr(),
                self.weights.data_ptr(), self.jac.data_ptr(), self.nrow,
                self.ncol, model, d_T, ntry, int(seed), psf_pars.data_ptr(),
                d_pres.data_ptr(), d_pflux.data_ptr(), guess.data_ptr(),
                d_res.data_ptr(), ntry_used.data_ptr(), work.data_ptr(),
                wbytes, prior=_prior_record(prior), maxfev=fp["maxfev"],
                ftol=fp["ftol"], xtol=fp["xtol"],
                analytic_jacobian=analytic_jacobian, em_tol=ep["tol"],
                em_miniter=ep["miniter"], em_maxiter=ep["maxiter"],
                bounds=getattr(prior, "bounds", None),
                stream=self._stream())
        self.set_psf(psf_pars[:n])
        nt = ntry_used.cpu().numpy()
        return dict(res=self._records(d_res, _lib.LM_RESULT_DTYPE),
                    psf_pars=psf_pars[:n].cpu().numpy(),
                    psf_res=self._records(d_pres, _lib.EM_RESULT_DTYPE),
                    psf_flux=self._records(d_pflux, _lib.FLUX_RESULT_DTYPE),
                    guess=guess[:n].cpu().numpy(),
                    psf_ntry=nt[0, :n].copy(), ntry=nt[1, :n].copy())

    # -------------------------------------------------------------- pixels
    def make_pixels(self):
        """the reference's pixel records (pixels/pixels.py:6-52) for every
        stamp: (pixels [n, H*W], npix [n]); pixels[i, :npix[i]] are valid"""
        return _lib.make_pixels(self.images.cpu().numpy(),
                                self.weights.cpu().numpy(),
                                self.jac.cpu().numpy(),
                                ignore_zero_weight=self.ignore_zero_weight)

    def shard(self, rank, world_size):
        """contiguous index range of this batch owned by `rank` (no copy)"""
        from .shard import shard_range
        lo, hi = shard_range(self.n, rank, world_size)
        sub = object.__new__(ObservationBatch)
        sub.device = self.device
        sub.ignore_zero_weight = self.ignore_zero_weight
        sub.images = self.images[lo:hi]
        sub.weights = self.weights[lo:hi]
        sub.jac = self.jac[lo:hi]
        sub.n, sub.nrow, sub.ncol = hi - lo, self.nrow, self.ncol
        sub.psf_pars = None if self.psf_pars is None else self.psf_pars[lo:hi]
        sub.npsf = self.npsf
        return sub
