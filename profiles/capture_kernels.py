#!/usr/bin/env python
"""
One launch of every kernel of the path other than the headline k_lm_fit<0>,
at a realistic size, inside a cudaProfilerStart/Stop range, so that

  ncu --set full --clock-control none --import-source on \
      --profile-from-start off -o gpurun_out/prof_kernels_<tag> \
      python profiles/capture_kernels.py

captures exactly one instance of each (k_em, k_admom, the forward-difference
k_lm_fit<1> on bdf stamps, k_lm_fit_multi, the coelliptical kernel,
k_noise_corr, k_weighted_sums, k_template_flux, k_loglike_tiled).
`profiles/ncu_multi.py <rep>` prints the per-kernel table.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import ngmix_b200 as ng
    from ngmix_b200 import _lib, sims, sims_device
    import bench_configs
    _lib.init(0)
    dev = torch.device("cuda", 0)
    s = torch.cuda.current_stream().cuda_stream
    jobs = []

    # ---- EM + admom + weighted moments on 25x25 psf stamps
    n = 20000
    p = sims_device.make_psf_batch_device(n, 1, dev)
    d_pars = torch.zeros(n * 18, dtype=torch.float64, device=dev)
    d_em = torch.zeros(n * _lib.EM_RESULT_DTYPE.itemsize, dtype=torch.uint8,
                       device=dev)
    d_am = torch.zeros(n * _lib.ADMOM_RESULT_DTYPE.itemsize,
                       dtype=torch.uint8, device=dev)
    jobs.append(("em", lambda: _lib.em_device(
        n, p["images"].data_ptr(), p["weights"].data_ptr(),
        p["jac"].data_ptr(), 25, 25, p["em_guess"].data_ptr(), 3,
        d_pars.data_ptr(), d_em.data_ptr(), stream=s)))
    jobs.append(("admom", lambda: _lib.admom_device(
        n, p["images"].data_ptr(), p["weights"].data_ptr(),
        p["jac"].data_ptr(), 25, 25, p["am_guess"].data_ptr(),
        d_am.data_ptr(), stream=s)))
    gm = ng.GaussMom(1.2).weight
    wt = np.repeat(gm.get_data()[None, :], n, axis=0)
    d_wt = torch.from_numpy(
        np.frombuffer(wt.tobytes(), dtype=np.uint8).copy()).to(dev)
    maxrad = 100.0 * float(np.sqrt(gm.get_T() / 2.0))
    d_mr = torch.full((n,), maxrad, dtype=torch.float64, device=dev)
    d_mom = torch.zeros(n * _lib.MOM_RESULT_DTYPE.itemsize, dtype=torch.uint8,
                        device=dev)
    jobs.append(("weighted_sums", lambda: _lib.weighted_sums_device(
        n, 1, d_wt.data_ptr(), p["images"].data_ptr(),
        p["weights"].data_ptr(), p["jac"].data_ptr(), 25, 25,
        d_mr.data_ptr(), 6, d_mom.data_ptr(), stream=s)))

    # ---- bdf forward-difference LM on 64x64
    nb = 4000
    tb = sims_device.make_bdf_batch_device(nb, 2, dev)
    d_rb = torch.zeros(nb * _lib.LM_RESULT_DTYPE.itemsize, dtype=torch.uint8,
                       device=dev)
    prb = bench_configs.bdf_prior(_lib)
    jobs.append(("lm_bdf", lambda: _lib.lm_fit_device(
        "bdf", nb, tb["images"].data_ptr(), tb["weights"].data_ptr(),
        tb["jac"].data_ptr(), 64, 64, tb["psf"].data_ptr(), 3,
        tb["guess"].data_ptr(), d_rb.data_ptr(), prior=prb,
        analytic_jacobian=False, stream=s)))

    # ---- exp LM at noise 1.0 and the stand-alone fdiff pass
    nl = 10000
    tl = sims_device.make_lm_batch_device(nl, 3, dev, noise=1.0)
    d_rl = torch.zeros(nl * _lib.LM_RESULT_DTYPE.itemsize, dtype=torch.uint8,
                       device=dev)
    jobs.append(("lm_exp_noise1", lambda: _lib.lm_fit_device(
        "exp", nl, tl["images"].data_ptr(), tl["weights"].data_ptr(),
        tl["jac"].data_ptr(), 48, 48, tl["psf"].data_ptr(), 3,
        tl["guess"].data_ptr(), d_rl.data_ptr(), stream=s)))

    # the stand-alone fdiff / loglike / s2n pass on the same stamps, at the
    # guess mixtures (what bench.py reports as fdiff_kernel)
    gm0, _ = _lib.gmix_fill("exp", tl["h_guess"])
    gp, _ = _lib.gmix_fill("full", tl["h_psf"].reshape(nl, -1))
    gmc = _lib.gmix_convolve(gm0, gp)
    _lib.gmix_set_norms(gmc)
    d_gmc = torch.from_numpy(
        np.frombuffer(gmc.tobytes(), dtype=np.uint8).copy()).to(dev)
    d_o4 = torch.zeros(nl * 4, dtype=torch.float64, device=dev)
    d_fd = torch.zeros(nl * 48 * 48, dtype=torch.float64, device=dev)
    jobs.append(("loglike_tiled", lambda: _lib.loglike_device(
        nl, 18, d_gmc.data_ptr(), tl["images"].data_ptr(),
        tl["weights"].data_ptr(), tl["jac"].data_ptr(), 48, 48,
        d_o4.data_ptr(), d_fdiff=d_fd.data_ptr(), stream=s)))

    # ---- multi-epoch / multi-band: 2 bands x 2 epochs
    nobj = nl // 4
    starts = torch.arange(0, 4 * nobj + 1, 4, dtype=torch.int32, device=dev)
    bands = torch.tensor([0, 0, 1, 1], dtype=torch.int32,
                         device=dev).repeat(nobj)
    g = tl["h_guess"][::4][:nobj]
    d_gm = torch.from_numpy(np.ascontiguousarray(
        np.column_stack([g, g[:, 5]]))).to(dev)
    d_rm = torch.zeros(nobj * _lib.LM_RESULT_DTYPE.itemsize,
                       dtype=torch.uint8, device=dev)
    jobs.append(("lm_multi", lambda: _lib.lm_fit_multi_device(
        "exp", nobj, starts.data_ptr(), bands.data_ptr(), 2,
        tl["images"].data_ptr(), tl["weights"].data_ptr(),
        tl["jac"].data_ptr(), 48, 48, tl["psf"].data_ptr(), 3,
        d_gm.data_ptr(), d_rm.data_ptr(), stream=s)))

    # ---- host-buffer calls whose kernels have no device wrapper in _lib:
    # coelliptical psf fits, noise covariance, template flux
    nh = 2000
    hb = sims.make_psf_batch(nh, seed0=5)
    T = ng.moments.fwhm_to_T(0.263 * 3.5)
    rng = np.random.RandomState(11)
    cg = np.zeros((nh, 8))
    cg[:, 0:2] = rng.uniform(-0.01, 0.01, (nh, 2))
    cg[:, 2:4] = rng.uniform(-0.05, 0.05, (nh, 2))
    cg[:, 4:6] = T * np.array(ng.guessers._COELLIP_F[2])[None] \
        * rng.uniform(0.99, 1.01, (nh, 2))
    cg[:, 6:8] = hb["images"].sum(axis=(1, 2))[:, None] \
        * np.array(ng.guessers._COELLIP_P[2])[None] \
        * rng.uniform(0.99, 1.01, (nh, 2))
    jobs.append(("coellip2", lambda: _lib.lm_fit_coellip(
        2, hb["images"], hb["weights"], hb["jac"], None, cg)))
    lb = sims.make_lm_batch(512, seed0=9, noise=0.05)
    noise_im = np.random.RandomState(3).normal(0, 0.05, lb["images"].shape)
    jobs.append(("noise_cov", lambda: ng.fit_batch(
        "exp", lb["images"], lb["weights"], lb["jac"], lb["psf"],
        lb["guess"], noise_images=noise_im)))
    gmp = np.repeat(np.array([[1.0, 0, 0, 0.1, 0, 0.1]]), nh, axis=0)
    jobs.append(("template_flux", lambda: _lib.template_flux(
        np.arange(nh + 1, dtype="i4"), hb["images"], hb["weights"],
        hb["jac"], gmp)))

    for name, fn in jobs:            # warm-up outside the profiled range
        fn()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    for name, fn in jobs:
        fn()
        torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("captured:", ", ".join(n for n, _ in jobs))


if __name__ == "__main__":
    main()
