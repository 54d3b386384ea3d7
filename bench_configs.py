"""
The other BASELINE.json configs, measured beside the headline line of
bench.py (which stays configs[1] at noise 0.01):

  config1   one exp LM fit through the public Fitter.go (latency)
  config2b  configs[1] at noise 1.0 (S/N 10-20: many-iteration fits)
  config3   100 000 PSF stamps 25x25: EM 3-gaussian fits + adaptive moments
  config4   1 000 000 bdf fits on 64x64 stamps, sharded over the ranks
  config5   200 000 objects x 5 sheared images, exp LM fits, sharded

Each returns a dict with throughput, fit statistics, a roofline fraction
(algorithmic fp64 flops of SURVEY.md 8d / measured DFMA peak) and, on a
single-GPU run, the agreement with the CPU oracle on a bounded sample of the
same stamps.  configs 4 and 5 are fixed-size jobs split over the ranks
(strong scaling); bench.py reports them under "configs".
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))

FLOPS_VALUE, FLOPS_DERIV = 28.0, 74.0
FLOPS_EM = 35.0            # per gaussian-pixel E+M step
FLOPS_ADMOM = 240.0 + 34   # per pixel and iteration: momsums + censums


def _oracle():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle
    oracle.build()
    return oracle


def _time_steps(torch, fn, steps, warmup, dist=None):
    """CUDA-event time of `steps` calls of fn on the current stream, ms per
    step, max over ranks"""
    for _ in range(warmup):
        fn()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def _sum_ranks(torch, dist, vals):
    if dist is None:
        return [float(v) for v in vals]
    t = torch.tensor([float(v) for v in vals], dtype=torch.float64,
                     device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [float(x) for x in t.tolist()]


def rel_dev(got, ref, cols):
    """max relative deviation |got - ref| / |ref| per named column"""
    out = {}
    for name, c in cols.items():
        d = np.abs(got[:, c] - ref[:, c]) / np.maximum(np.abs(ref[:, c]),
                                                      1e-300)
        out[name] = float(d.max()) if d.size else 0.0
    return out


def lm_parity(res, ores, npars, Tcol=4, Fcol=None):
    """agreement of device records `res` with oracle records `ores` on the
    same stamps: flags / nfev / ier counts, parameters within
    max(1e-5 |p|, 1e-3 sigma), and the true max relative deviation of T and
    flux (and max |dg|) over the fits both sides flag as good"""
    n = ores["flags"].size
    Fcol = npars - 1 if Fcol is None else Fcol
    same = ores["flags"] == res["flags"][:n]
    good = (ores["flags"] == 0) & (res["flags"][:n] == 0)
    perr = np.abs(res["pars"][:n, :npars] - ores["pars"][:, :npars])
    tol = np.maximum(1e-5 * np.abs(ores["pars"][:, :npars]),
                     1e-3 * ores["pars_err"][:, :npars])
    out = {"n": int(n), "flags_equal": int(same.sum()),
           "nfev_ier_equal": int(((ores["nfev"] == res["nfev"][:n])
                                  & (ores["ier"] == res["ier"][:n])).sum()),
           "pars_within_tol": int(np.all(perr <= tol, axis=1)[good].sum()),
           "n_good": int(good.sum())}
    if good.any():
        g = res["pars"][:n][good]
        o = ores["pars"][good]
        rd = rel_dev(g, o, {"T": Tcol, "flux": Fcol})
        out["max_rel_dev_T"] = rd["T"]
        out["max_rel_dev_flux"] = rd["flux"]
        out["max_abs_dev_g"] = float(np.abs(g[:, 2:4] - o[:, 2:4]).max())
        out["n_rel_dev_T_or_flux_above_1e-5"] = int(
            ((np.abs(g[:, Tcol] - o[:, Tcol])
              > 1e-5 * np.abs(o[:, Tcol]))
             | (np.abs(g[:, Fcol] - o[:, Fcol])
                > 1e-5 * np.abs(o[:, Fcol]))).sum())
    return out


def bdf_prior(_lib):
    """PriorBDFSep of examples/fitting/fitting_bd_empsf.py:101-142"""
    pr = _lib.Prior()
    pr.present = 1
    pr.g_kind = 1
    pr.cen1, pr.cen2, pr.sigma1, pr.sigma2 = 0.0, 0.0, 0.263, 0.263
    pr.g_sigma = 0.3
    pr.T = _lib.Prior1D(1, 0, -1.0, 1.0e3, 0, 0)
    pr.fracdev = _lib.Prior1D(3, 0, 0.5, 0.1, 0, 0)
    pr.F = _lib.Prior1D(1, 0, -100.0, 1.0e9, 0, 0)
    return pr


# ---------------------------------------------------------------------------
def config1_latency(args, _lib):
    """configs[0]: one exp fit of one 48x48 stamp with a 1-gaussian psf
    through the public Fitter.go (host Observation in, result dict out)"""
    import ngmix_b200 as ng
    rng = np.random.RandomState(42)
    dy, dx = rng.uniform(-0.5, 0.5, size=2)
    jac = ng.DiagonalJacobian(scale=0.263, row=23.5 + dy, col=23.5 + dx)
    truth = np.array([0.0, 0.0, 0.1, 0.05, 0.4, 100.0])
    psf_gm = ng.GMixModel([0.0, 0.0, 0.01, -0.01, 0.27, 1.0], "gauss")
    gm = ng.GMixModel(truth, "exp").convolve(psf_gm)
    im = gm.make_image((48, 48), jacobian=jac)
    im = im + rng.normal(0.0, 0.01, size=im.shape)
    wt = np.zeros_like(im) + 1.0 / 0.01 ** 2
    pjac = ng.DiagonalJacobian(scale=0.263, row=12.0, col=12.0)
    psf_obs = ng.Observation(psf_gm.make_image((25, 25), jacobian=pjac),
                             jacobian=pjac, gmix=psf_gm)
    obs = ng.Observation(im, weight=wt, jacobian=jac, psf=psf_obs)
    guess = truth * (1.0 + rng.uniform(-0.05, 0.05, size=6))
    guess[0:2] = rng.uniform(-0.01, 0.01, size=2)
    fitter = ng.Fitter(model="exp")
    res = fitter.go(obs, guess)
    ts = []
    for _ in range(50):
        t0 = time.perf_counter()
        res = fitter.go(obs, guess)
        ts.append(time.perf_counter() - t0)
    out = {"workload": "configs[0]: one exp LM fit, 48x48 stamp, 1-gaussian "
                       "psf, through Fitter.go (host in, dict out)",
           "latency_ms_median": 1e3 * float(np.median(ts)),
           "latency_ms_min": 1e3 * float(np.min(ts)),
           "nfev": int(res["nfev"]), "flags": int(res["flags"])}
    if not args.no_cpu:
        oracle = _oracle()
        jrec = np.asarray(jac._data.view("f8")).reshape(1, 8)
        ppars = psf_gm.get_full_pars().reshape(1, 1, 6)
        t0 = time.perf_counter()
        for _ in range(20):
            o = oracle.lm_fit_batch("exp", im[None], wt[None], jrec, ppars,
                                    guess[None], nthreads=1)
        out["cpu_port_latency_ms"] = 1e3 * (time.perf_counter() - t0) / 20
        out["parity"] = {
            "flags_equal": bool(o["flags"][0] == res["flags"]),
            "nfev_equal": bool(o["nfev"][0] == res["nfev"]),
            "max_rel_dev_pars": float(np.max(
                np.abs(res["pars"] - o["pars"][0, :6])
                / np.maximum(np.abs(o["pars"][0, :6]), 1e-300))[()]),
            "max_rel_dev_T_flux": float(np.max(
                np.abs(res["pars"][4:6] - o["pars"][0, 4:6])
                / np.abs(o["pars"][0, 4:6])))}
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import ref_arm
        if ref_arm.available()[0]:
            b = dict(images=im[None], weights=wt[None], jac=jrec, psf=ppars,
                     guess=guess[None])
            lat, rr = ref_arm.single_fit_latency(b)
            out["reference_numba_latency_ms"] = 1e3 * lat
            out["parity"]["reference_nfev_equal"] = bool(
                rr["nfev"] == res["nfev"])
            out["parity"]["max_rel_dev_vs_reference_T_flux"] = float(np.max(
                np.abs(res["pars"][4:6] - rr["pars"][4:6])
                / np.abs(rr["pars"][4:6])))
    return out


# ---------------------------------------------------------------------------
def _lm_run(args, torch, _lib, dist, t, model, n, dims, npsf, prior=None,
            analytic=True):
    H, W = dims
    d_res = torch.zeros(max(n, 1) * _lib.LM_RESULT_DTYPE.itemsize,
                        dtype=torch.uint8, device=t["images"].device)
    s = torch.cuda.current_stream().cuda_stream

    def fit():
        _lib.lm_fit_device(model, n, t["images"].data_ptr(),
                           t["weights"].data_ptr(), t["jac"].data_ptr(), H,
                           W, t["psf"].data_ptr(), npsf,
                           t["guess"].data_ptr(), d_res.data_ptr(),
                           prior=prior, analytic_jacobian=analytic, stream=s)

    ms = _time_steps(torch, fit, args.cfg_steps, 1, dist)
    res = np.frombuffer(d_res.cpu().numpy().tobytes(),
                        dtype=_lib.LM_RESULT_DTYPE)[:n]
    return ms, res


def config2_noisy(args, torch, _lib, dev, rank, world, dist, dfma_per_s):
    """configs[1] at noise 1.0: S/N 10-20, up to tens of evaluations per
    fit, uneven work per stamp (the work queue's case)"""
    from ngmix_b200 import sims_device
    n = args.nobj
    t = sims_device.make_lm_batch_device(n, 7000 + rank, dev, noise=1.0)
    ms, res = _lm_run(args, torch, _lib, dist, t, "exp", n, (48, 48), 3)
    nfev, njev = res["nfev"].astype("f8"), res["njev"].astype("f8")
    flops = 18 * 2304 * float((nfev * FLOPS_VALUE + njev * FLOPS_DERIV).sum())
    tot_n, tot_flops = _sum_ranks(torch, dist, [n, flops])
    out = {"workload": "configs[1] at noise 1.0 (S/N 10-20): %d exp LM fits "
                       "per GPU, 48x48, 3-gaussian psf" % n,
           "fits_per_s": tot_n / (ms * 1e-3), "ms_per_step": ms,
           "median_nfev": float(np.median(nfev)),
           "max_nfev": float(nfev.max()),
           "nonzero_flags": int((res["flags"] != 0).sum()),
           "roofline_frac": tot_flops / (ms * 1e-3)
           / (2.0 * dfma_per_s * world)}
    if not args.no_cpu and world == 1:
        oracle = _oracle()
        ns = min(n, 2048)
        t0 = time.perf_counter()
        o = oracle.lm_fit_batch(
            "exp", t["images"][:ns].cpu().numpy(),
            t["weights"][:ns].cpu().numpy(), t["h_jac"][:ns],
            t["h_psf"][:ns], t["h_guess"][:ns], nthreads=os.cpu_count())
        out["cpu_port_fits_per_s"] = ns / (time.perf_counter() - t0)
        out["parity_on_sample"] = lm_parity(res, o, 6)
    return out


def config3_psf(args, torch, _lib, dev, rank, world, dist, dfma_per_s):
    """configs[2]: 100 000 PSF stamps 25x25 per GPU, EM 3-gaussian fits
    (tol 1e-5, miniter 40, maxiter 500) and adaptive moments"""
    from ngmix_b200 import sims_device
    n = args.npsf
    t = sims_device.make_psf_batch_device(n, 4242 + rank, dev)
    d_pars = torch.zeros(n * 18, dtype=torch.float64, device=dev)
    d_em = torch.zeros(n * _lib.EM_RESULT_DTYPE.itemsize, dtype=torch.uint8,
                       device=dev)
    d_am = torch.zeros(n * _lib.ADMOM_RESULT_DTYPE.itemsize,
                       dtype=torch.uint8, device=dev)
    s = torch.cuda.current_stream().cuda_stream

    def em():
        _lib.em_device(n, t["images"].data_ptr(), t["weights"].data_ptr(),
                       t["jac"].data_ptr(), 25, 25, t["em_guess"].data_ptr(),
                       3, d_pars.data_ptr(), d_em.data_ptr(), stream=s)

    def am():
        _lib.admom_device(n, t["images"].data_ptr(), t["weights"].data_ptr(),
                          t["jac"].data_ptr(), 25, 25,
                          t["am_guess"].data_ptr(), d_am.data_ptr(), stream=s)

    ms_em = _time_steps(torch, em, args.cfg_steps, 1, dist)
    ms_am = _time_steps(torch, am, max(args.cfg_steps, 5), 2, dist)
    rem = np.frombuffer(d_em.cpu().numpy().tobytes(),
                        dtype=_lib.EM_RESULT_DTYPE)
    ram = np.frombuffer(d_am.cpu().numpy().tobytes(),
                        dtype=_lib.ADMOM_RESULT_DTYPE)
    em_flops = float(rem["numiter"].sum()) * 3 * 625 * FLOPS_EM
    am_flops = float(ram["numiter"].sum()) * 625 * FLOPS_ADMOM
    tot_n, em_flops, am_flops = _sum_ranks(torch, dist,
                                           [n, em_flops, am_flops])
    out = {"workload": "configs[2]: %d PSF stamps 25x25 per GPU, EM "
                       "3-gaussian (tol 1e-5) + admom" % n,
           "em": {"fits_per_s": tot_n / (ms_em * 1e-3), "ms_per_step": ms_em,
                  "median_numiter": float(np.median(rem["numiter"])),
                  "frac_maxiter": float(np.mean(rem["flags"] == 32)),
                  "roofline_frac": em_flops / (ms_em * 1e-3)
                  / (2.0 * dfma_per_s * world)},
           "admom": {"fits_per_s": tot_n / (ms_am * 1e-3),
                     "ms_per_step": ms_am,
                     "median_numiter": float(np.median(ram["numiter"])),
                     "nonzero_flags": int((ram["flags"] != 0).sum()),
                     "roofline_frac": am_flops / (ms_am * 1e-3)
                     / (2.0 * dfma_per_s * world)}}
    if not args.no_cpu and world == 1:
        oracle = _oracle()
        cores = os.cpu_count()
        ns = min(n, 40 * cores)
        im = t["images"][:max(ns, min(n, 2000 * cores))].cpu().numpy()
        wt = t["weights"][:im.shape[0]].cpu().numpy()
        t0 = time.perf_counter()
        ores, _ = oracle.em_go_batch(im[:ns], wt[:ns], t["h_jac"][:ns],
                                     t["h_em_guess"][:ns], nthreads=cores)
        out["em"]["cpu_port_fits_per_s"] = ns / (time.perf_counter() - t0)
        out["em"]["parity_on_sample"] = {
            "n": int(ns),
            "numiter_equal": int((ores["numiter"]
                                  == rem["numiter"][:ns]).sum()),
            "flags_equal": int((ores["flags"] == rem["flags"][:ns]).sum())}
        na = im.shape[0]
        t0 = time.perf_counter()
        oam = oracle.admom_go_batch(im, wt, t["h_jac"][:na],
                                    t["h_am_guess"][:na], nthreads=cores)
        out["admom"]["cpu_port_fits_per_s"] = na / (time.perf_counter() - t0)
        out["admom"]["parity_on_sample"] = {
            "n": int(na),
            "numiter_equal": int((oam["numiter"]
                                  == ram["numiter"][:na]).sum()),
            "flags_equal": int((oam["flags"] == ram["flags"][:na]).sum())}
        out["cpu_cores"] = cores
    return out


def config4_bdf(args, torch, _lib, dev, rank, world, dist, dfma_per_s):
    """configs[3]: bdf fits (7 parameters, 16 x 3 = 48 gaussians) on 64x64
    stamps with PriorBDFSep and the forward-difference Jacobian of the
    reference's lmdif path; a fixed job of args.nbdf objects split by index
    over the ranks"""
    from ngmix_b200 import sims_device, shard
    lo, hi = shard.shard_range(args.nbdf, rank, world)
    n = hi - lo
    t = sims_device.make_bdf_batch_device(n, 777000 + rank, dev)
    ms, res = _lm_run(args, torch, _lib, dist, t, "bdf", n, (64, 64), 3,
                      prior=bdf_prior(_lib), analytic=False)
    nfev = res["nfev"].astype("f8")
    flops = 48 * 4096 * float(nfev.sum()) * FLOPS_VALUE
    tot_n, tot_flops, nbad = _sum_ranks(
        torch, dist, [n, flops, (res["flags"] != 0).sum()])
    out = {"workload": "configs[3]: %d bdf LM fits in total (%d on this "
                       "rank), 64x64 stamps, 48 gaussians, PriorBDFSep, "
                       "forward-difference Jacobian" % (args.nbdf, n),
           "scaling": "strong",
           "fits_per_s": tot_n / (ms * 1e-3), "ms_per_step": ms,
           "median_nfev": float(np.median(nfev)),
           "nonzero_flags": int(nbad),
           "gaussian_pixel_evals_per_s": tot_flops / FLOPS_VALUE
           / (ms * 1e-3),
           "roofline_frac": tot_flops / (ms * 1e-3)
           / (2.0 * dfma_per_s * world)}
    if not args.no_cpu and world == 1:
        oracle = _oracle()
        cores = os.cpu_count()
        ns = min(n, 32 * cores)
        t0 = time.perf_counter()
        o = oracle.lm_fit_batch(
            "bdf", t["images"][:ns].cpu().numpy(),
            t["weights"][:ns].cpu().numpy(), t["h_jac"][:ns],
            t["h_psf"][:ns], t["h_guess"][:ns],
            prior=oracle.make_prior("bdf"), nthreads=cores)
        out["cpu_port_fits_per_s"] = ns / (time.perf_counter() - t0)
        out["cpu_cores"] = cores
        out["parity_on_sample"] = lm_parity(res, o, 7, Tcol=4, Fcol=6)
    return out


def config5_metacal(args, torch, _lib, dev, rank, world, dist, dfma_per_s):
    """configs[4]: metacal-style batch, 5 sheared images per object fit as
    independent exp LM fits (round 1-gaussian reconvolution psf): a fixed
    job of args.nmcal objects split by index over the ranks; R11 / R22 from
    the +-0.01 pairs as the sanity number"""
    from ngmix_b200 import sims_device, shard
    lo, hi = shard.shard_range(args.nmcal, rank, world)
    nobj = hi - lo
    n = 5 * nobj
    t = sims_device.make_metacal_batch_device(nobj, 555000 + rank, dev)
    ms, res = _lm_run(args, torch, _lib, dist, t, "exp", n, (48, 48), 1)
    nfev, njev = res["nfev"].astype("f8"), res["njev"].astype("f8")
    flops = 6 * 2304 * float((nfev * FLOPS_VALUE + njev * FLOPS_DERIV).sum())
    g = res["pars"][:, 2:4].reshape(nobj, 5, 2)
    ok = (res["flags"].reshape(nobj, 5) == 0).all(axis=1)
    s11 = float(np.sum(g[ok, 1, 0] - g[ok, 2, 0]))
    s22 = float(np.sum(g[ok, 3, 1] - g[ok, 4, 1]))
    tot_n, tot_flops, s11, s22, nok, nbad = _sum_ranks(
        torch, dist, [n, flops, s11, s22, ok.sum(),
                      (res["flags"] != 0).sum()])
    out = {"workload": "configs[4]: %d objects x 5 sheared images in total "
                       "(%d objects on this rank), exp LM fits, 48x48, "
                       "1-gaussian psf (6 gaussians)" % (args.nmcal, nobj),
           "scaling": "strong",
           "fits_per_s": tot_n / (ms * 1e-3),
           "objects_per_s": tot_n / 5 / (ms * 1e-3), "ms_per_step": ms,
           "median_nfev": float(np.median(nfev)),
           "nonzero_flags": int(nbad),
           "R11": s11 / max(nok, 1) / 0.02, "R22": s22 / max(nok, 1) / 0.02,
           "roofline_frac": tot_flops / (ms * 1e-3)
           / (2.0 * dfma_per_s * world)}
    if not args.no_cpu and world == 1:
        oracle = _oracle()
        ns = min(n, 2560)
        t0 = time.perf_counter()
        o = oracle.lm_fit_batch(
            "exp", t["images"][:ns].cpu().numpy(),
            t["weights"][:ns].cpu().numpy(), t["h_jac"][:ns],
            t["h_psf"][:ns], t["h_guess"][:ns], nthreads=os.cpu_count())
        out["cpu_port_fits_per_s"] = ns / (time.perf_counter() - t0)
        out["parity_on_sample"] = lm_parity(res, o, 6)
    return out
