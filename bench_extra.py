#!/usr/bin/env python
"""
Secondary measurements (BASELINE.json configs[2]): EM 3-gaussian fits and
adaptive moments on 25x25 PSF stamps, device-resident, beside the CPU oracle
port on the host threads.  Prints one JSON line; not the headline bench.

    python bench_extra.py [--nobj 20000] [--steps 3]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


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


def bench_gaussmom(args, torch, _lib, b, t, dev):
    """Gaussian-weighted moments (GaussMom, fwhm 1.2) of the PSF stamps"""
    import ngmix_b200 as ngmix
    n = b["images"].shape[0]
    out = {}
    for nmom in (6, 17):
        gm = ngmix.GaussMom(1.2, with_higher_order=(nmom == 17)).weight
        wt = np.repeat(gm.get_data()[None, :], n, axis=0)
        d_wt = torch.from_numpy(
            np.frombuffer(wt.tobytes(), dtype=np.uint8).copy()).to(dev)
        maxrad = 100.0 * float(np.sqrt(gm.get_T() / 2.0))
        d_mr = torch.full((n,), maxrad, dtype=torch.float64, device=dev)
        d_res = torch.zeros(n * _lib.MOM_RESULT_DTYPE.itemsize,
                            dtype=torch.uint8, device=dev)
        s = torch.cuda.current_stream().cuda_stream

        def fn():
            _lib.weighted_sums_device(
                n, 1, d_wt.data_ptr(), t["images"].data_ptr(),
                t["weights"].data_ptr(), t["jac"].data_ptr(), 25, 25,
                d_mr.data_ptr(), nmom, d_res.data_ptr(), stream=s)

        fn()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        o = {"stamps_per_s": n / (ms * 1e-3), "ms_per_step": ms}
        if not args.no_cpu:
            import oracle
            ns = min(n, 2000)
            wpars = gm.get_full_pars()
            t0 = time.perf_counter()
            refs = [oracle.weighted_sums(b["images"][i], b["weights"][i],
                                         b["jac"][i], wpars, maxrad, nmom)
                    for i in range(ns)]
            dt = time.perf_counter() - t0
            rec = np.frombuffer(d_res.cpu().numpy().tobytes(),
                                dtype=_lib.MOM_RESULT_DTYPE)
            worst = max(np.abs(rec["sums"][i][:nmom] - refs[i]["sums"]).max()
                        / np.abs(refs[i]["sums"]).max() for i in range(ns))
            o["cpu_port_stamps_per_s_1core"] = ns / dt
            o["max_rel_diff_sums_on_sample"] = float(worst)
            o["sample"] = ns
        out["nmom%d" % nmom] = o
    return out


def bench_multi(args, _lib, sims, torch, dev):
    """multi-epoch / multi-band exp fits: 2 bands x 2 epochs of 48x48 stamps
    per object (Fitter.go on a MultiBandObsList), host-buffer call"""
    import ngmix_b200 as ngmix
    nobj = 2000
    b = sims.make_lm_batch(4 * nobj, seed0=31000, noise=0.02)
    starts = np.arange(0, 4 * nobj + 1, 4)
    bands = np.tile(np.array([0, 0, 1, 1], dtype="i4"), nobj)
    g = b["guess"][::4]
    guesses = np.column_stack([g, g[:, 5]])
    # the four stamps of an object share the truth of its first stamp only
    # approximately (independent sims): the fit is a throughput workload
    t0 = time.perf_counter()
    recs = ngmix.fit_batch_multi("exp", starts, bands, 2, b["images"],
                                 b["weights"], b["jac"], b["psf"], guesses)
    t1 = time.perf_counter()
    recs = ngmix.fit_batch_multi("exp", starts, bands, 2, b["images"],
                                 b["weights"], b["jac"], b["psf"], guesses)
    dt = time.perf_counter() - t1
    # the same call from page-locked host arrays (ngmix_b200.pinned_copy):
    # pageable copies run at ~10 GB/s and bound the call above
    pim, pwt = ngmix.pinned_copy(b["images"]), ngmix.pinned_copy(b["weights"])
    ngmix.fit_batch_multi("exp", starts, bands, 2, pim, pwt, b["jac"],
                          b["psf"], guesses)
    t2 = time.perf_counter()
    ngmix.fit_batch_multi("exp", starts, bands, 2, pim, pwt, b["jac"],
                          b["psf"], guesses)
    dt_pinned = time.perf_counter() - t2
    del pim, pwt
    # device-resident timing of the same workload
    tt = {k: torch.from_numpy(np.ascontiguousarray(b[k])).to(dev)
          for k in ("images", "weights", "jac", "psf")}
    d_ss = torch.from_numpy(starts.astype("i4")).to(dev)
    d_sb = torch.from_numpy(bands).to(dev)
    d_g = torch.from_numpy(np.ascontiguousarray(guesses)).to(dev)
    d_res = torch.zeros(nobj * _lib.LM_RESULT_DTYPE.itemsize,
                        dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream

    def fn():
        _lib.lm_fit_multi_device(
            "exp", nobj, d_ss.data_ptr(), d_sb.data_ptr(), 2,
            tt["images"].data_ptr(), tt["weights"].data_ptr(),
            tt["jac"].data_ptr(), 48, 48, tt["psf"].data_ptr(), 3,
            d_g.data_ptr(), d_res.data_ptr(), stream=st)

    fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    o = {"config": "%d objects x (2 bands x 2 epochs) 48x48, exp, 7 "
                   "parameters" % nobj,
         "objects_per_s": nobj / (ms * 1e-3),
         "stamps_per_s": 4 * nobj / (ms * 1e-3), "ms_per_step": ms,
         "host_call_objects_per_s": nobj / dt,
         "host_call_pinned_objects_per_s": nobj / dt_pinned,
         "median_nfev": float(np.median(recs["nfev"]))}
    if not args.no_cpu:
        import oracle
        ns = 32
        t0 = time.perf_counter()
        same = 0
        for i in range(ns):
            sl = slice(4 * i, 4 * i + 4)
            r = oracle.lm_fit_multi("exp", b["images"][sl], b["weights"][sl],
                                    b["jac"][sl], b["psf"][sl], bands[sl], 2,
                                    guesses[i])
            same += int(r["flags"] == recs["flags"][i]
                        and r["nfev"] == recs["nfev"][i])
        o["cpu_port_objects_per_s_1core"] = ns / (time.perf_counter() - t0)
        o["flags_nfev_equal_on_sample"] = same
        o["sample"] = ns
    return o


def bench_bdf(args, torch, _lib, sims, dev):
    """configs[3]: bdf fits (7 pars, 16 x 3 = 48 gaussians, 64x64 stamps,
    PriorBDFSep, forward-difference Jacobian as the reference's lmdif path)"""
    n = args.nbdf
    b = sims.make_bdf_batch(n, seed0=777000, dims=(64, 64), noise=0.02)
    t = {k: torch.from_numpy(np.ascontiguousarray(b[k])).to(dev)
         for k in ("images", "weights", "jac", "psf", "guess")}
    d_res = torch.zeros(n * _lib.LM_RESULT_DTYPE.itemsize, dtype=torch.uint8,
                        device=dev)
    s = torch.cuda.current_stream().cuda_stream
    pr = bdf_prior(_lib)

    def fit():
        _lib.lm_fit_device("bdf", n, t["images"].data_ptr(),
                           t["weights"].data_ptr(), t["jac"].data_ptr(), 64,
                           64, t["psf"].data_ptr(), 3, t["guess"].data_ptr(),
                           d_res.data_ptr(), prior=pr, stream=s)

    fit()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        fit()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    res = np.frombuffer(d_res.cpu().numpy().tobytes(),
                        dtype=_lib.LM_RESULT_DTYPE)
    nfev = res["nfev"].astype(float)
    o = {"config": "configs[3]: %d bdf fits, 64x64 stamps, 48 gaussians, "
                   "PriorBDFSep, forward-difference Jacobian" % n,
         "fits_per_s": n / (ms * 1e-3), "ms_per_step": ms,
         "median_nfev": float(np.median(nfev)),
         "nonzero_flags": int((res["flags"] != 0).sum()),
         # every nfev is one full-mixture evaluation of the stamp
         "model_pixel_evals_per_s": float(nfev.sum() * 4096 / (ms * 1e-3)),
         "gaussian_pixel_evals_per_s": float(nfev.sum() * 4096 * 48
                                             / (ms * 1e-3))}
    if not args.no_cpu:
        import oracle
        cores = os.cpu_count()
        ns = min(n, 4 * cores)
        t0 = time.perf_counter()
        ref = oracle.lm_fit_batch("bdf", b["images"][:ns], b["weights"][:ns],
                                  b["jac"][:ns], b["psf"][:ns],
                                  b["guess"][:ns],
                                  prior=oracle.make_prior("bdf"),
                                  nthreads=cores)
        dt = time.perf_counter() - t0
        o["cpu_port_fits_per_s"] = ns / dt
        o["cpu_cores"] = cores
        o["sample"] = ns
        o["flags_equal_on_sample"] = int((ref["flags"] == res["flags"][:ns]).sum())
        o["nfev_equal_on_sample"] = int((ref["nfev"] == res["nfev"][:ns]).sum())
    return o


def bench_metacal(args, torch, _lib, sims, dev):
    """configs[4]: metacal-style batch, 5 sheared images per object fit as
    independent exp LM fits (1-gaussian reconvolution psf); R11 / R22 from
    the +-0.01 pairs as a sanity number"""
    nobj = args.nmcal
    b = sims.make_metacal_batch(nobj, seed0=555000, noise=0.01)
    n = 5 * nobj
    t = {k: torch.from_numpy(np.ascontiguousarray(b[k])).to(dev)
         for k in ("images", "weights", "jac", "psf", "guess")}
    d_res = torch.zeros(n * _lib.LM_RESULT_DTYPE.itemsize, dtype=torch.uint8,
                        device=dev)
    s = torch.cuda.current_stream().cuda_stream

    def fit():
        _lib.lm_fit_device("exp", n, t["images"].data_ptr(),
                           t["weights"].data_ptr(), t["jac"].data_ptr(), 48,
                           48, t["psf"].data_ptr(), 1, t["guess"].data_ptr(),
                           d_res.data_ptr(), stream=s)

    fit()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        fit()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    res = np.frombuffer(d_res.cpu().numpy().tobytes(),
                        dtype=_lib.LM_RESULT_DTYPE)
    g = res["pars"][:, 2:4].reshape(nobj, 5, 2)
    ok = (res["flags"].reshape(nobj, 5) == 0).all(axis=1)
    R11 = float(np.mean(g[ok, 1, 0] - g[ok, 2, 0]) / 0.02)
    R22 = float(np.mean(g[ok, 3, 1] - g[ok, 4, 1]) / 0.02)
    return {"config": "configs[4]: %d objects x 5 sheared images, exp LM "
                      "fits, 48x48, 1-gaussian psf (6 gaussians)" % nobj,
            "fits_per_s": n / (ms * 1e-3),
            "objects_per_s": nobj / (ms * 1e-3), "ms_per_step": ms,
            "nonzero_flags": int((res["flags"] != 0).sum()),
            "median_nfev": float(np.median(res["nfev"])),
            "R11": R11, "R22": R22}


def bench_coellip(args, _lib, b):
    """coelliptical psf fits (CoellipFitter, forward differences) of the PSF
    stamps, through the host-buffer call"""
    import ngmix_b200 as ngmix
    n = min(b["images"].shape[0], 20000)
    out = {}
    for ngauss in (2, 3):
        # CoellipPSFGuesser's recipe (guessers.py:1106-1216), vectorised over
        # the batch instead of one Observation at a time
        guess = np.zeros((n, 4 + 2 * ngauss))
        rng = np.random.RandomState(11)
        T = ngmix.moments.fwhm_to_T(0.263 * 3.5)
        flux = b["images"][:n].sum(axis=(1, 2))
        for i in range(n):
            g = np.zeros(4 + 2 * ngauss)
            g[0:2] = rng.uniform(-0.01, 0.01, 2)
            g[2:4] = rng.uniform(-0.05, 0.05, 2)
            fr = ngmix.guessers._COELLIP_F[ngauss]
            pr = ngmix.guessers._COELLIP_P[ngauss]
            g[4:4 + ngauss] = T * np.array(fr) * rng.uniform(0.99, 1.01,
                                                               ngauss)
            g[4 + ngauss:] = flux[i] * np.array(pr) * rng.uniform(0.99, 1.01,
                                                                  ngauss)
            guess[i] = g
        _lib.lm_fit_coellip(ngauss, b["images"][:64], b["weights"][:64],
                            b["jac"][:64], None, guess[:64])
        dts = []
        for _ in range(3):
            t0 = time.perf_counter()
            r = _lib.lm_fit_coellip(ngauss, b["images"][:n], b["weights"][:n],
                                    b["jac"][:n], None, guess)
            dts.append(time.perf_counter() - t0)
        dt = float(np.median(dts))
        # the kernel alone, stamps resident (CUDA events)
        import torch
        dev = torch.device("cuda", 0)
        tt = {k: torch.from_numpy(np.ascontiguousarray(b[k][:n])).to(dev)
              for k in ("images", "weights", "jac")}
        d_g = torch.from_numpy(guess).to(dev)
        d_res = torch.zeros(n * _lib.LM_RESULT_WIDE_DTYPE.itemsize,
                            dtype=torch.uint8, device=dev)
        st = torch.cuda.current_stream().cuda_stream

        def fn():
            _lib.lm_fit_coellip_device(
                ngauss, n, tt["images"].data_ptr(), tt["weights"].data_ptr(),
                tt["jac"].data_ptr(), 25, 25, 0, 0, d_g.data_ptr(),
                d_res.data_ptr(), stream=st)

        fn()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            fn()
        e1.record()
        torch.cuda.synchronize()
        kms = e0.elapsed_time(e1) / 3
        del tt, d_g, d_res
        out["ngauss%d" % ngauss] = {
            "fits_per_s": n / (kms * 1e-3), "ms_per_step": kms,
            "fits_per_s_host_call": n / dt, "n": n,
            "calls_s": [float(x) for x in dts],
            "median_nfev": float(np.median(r["nfev"])),
            "nonzero_flags": int((r["flags"] != 0).sum())}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nobj", type=int, default=20000)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--nbdf", type=int, default=2000,
                    help="configs[3] sample: bdf fits on 64x64 stamps")
    ap.add_argument("--nmcal", type=int, default=2000,
                    help="configs[4] sample: objects with 5 sheared images")
    args = ap.parse_args()
    import torch
    from ngmix_b200 import _lib, sims
    _lib.init(0)
    dev = torch.device("cuda", 0)
    n = args.nobj
    b = sims.make_psf_batch(n, seed0=424242)
    t = {k: torch.from_numpy(np.ascontiguousarray(b[k])).to(dev)
         for k in ("images", "weights", "jac", "em_guess", "am_guess")}
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

    out = {"config": "configs[2]: %d PSF stamps 25x25, EM 3-gaussian "
                     "(tol 1e-5, miniter 40, maxiter 500) + admom" % n}
    for name, fn in (("em", em), ("admom", am)):
        fn()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        out[name] = {"fits_per_s": n / (ms * 1e-3), "ms_per_step": ms}
    rem = np.frombuffer(d_em.cpu().numpy().tobytes(),
                        dtype=_lib.EM_RESULT_DTYPE)
    ram = np.frombuffer(d_am.cpu().numpy().tobytes(),
                        dtype=_lib.ADMOM_RESULT_DTYPE)
    out["em"]["median_numiter"] = float(np.median(rem["numiter"]))
    out["em"]["frac_maxiter"] = float(np.mean(rem["flags"] == 32))
    # gaussian-pixel E+M evaluations per second (3 gaussians x 625 px x iters)
    out["em"]["gauss_pixel_evals_per_s"] = float(
        rem["numiter"].sum() * 3.0 * 625 / (out["em"]["ms_per_step"] * 1e-3))
    out["admom"]["median_numiter"] = float(np.median(ram["numiter"]))
    if not args.no_cpu:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import oracle
        oracle.build()
        cores = os.cpu_count()
        ns = min(n, 40 * cores)
        t0 = time.perf_counter()
        ores, _ = oracle.em_go_batch(b["images"][:ns], b["weights"][:ns],
                                     b["jac"][:ns], b["em_guess"][:ns],
                                     nthreads=cores)
        dt = time.perf_counter() - t0
        out["em"]["cpu_port_fits_per_s"] = ns / dt
        out["em"]["cpu_cores"] = cores
        out["em"]["numiter_equal_on_sample"] = int(
            (ores["numiter"] == rem["numiter"][:ns]).sum())
        out["em"]["flags_equal_on_sample"] = int(
            (ores["flags"] == rem["flags"][:ns]).sum())
        out["em"]["sample"] = ns
        na = min(n, 2000 * cores)
        t0 = time.perf_counter()
        oam = oracle.admom_go_batch(b["images"][:na], b["weights"][:na],
                                    b["jac"][:na], b["am_guess"][:na],
                                    nthreads=cores)
        dt = time.perf_counter() - t0
        out["admom"]["cpu_port_fits_per_s"] = na / dt
        out["admom"]["numiter_equal_on_sample"] = int(
            (oam["numiter"] == ram["numiter"][:na]).sum())
        out["admom"]["flags_equal_on_sample"] = int(
            (oam["flags"] == ram["flags"][:na]).sum())
        out["admom"]["sample"] = na
    out["gaussmom"] = bench_gaussmom(args, torch, _lib, b, t, dev)
    out["multi"] = bench_multi(args, _lib, sims, torch, dev)
    if args.nbdf > 0:
        out["bdf"] = bench_bdf(args, torch, _lib, sims, dev)
    if args.nmcal > 0:
        out["metacal"] = bench_metacal(args, torch, _lib, sims, dev)
    out["coellip"] = bench_coellip(args, _lib, b)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
