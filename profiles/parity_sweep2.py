"""Round-2 parity sweeps at the sample sizes of SURVEY.md 8d, against the
oracle on identical stamps: mismatch counts per sample and the true maximum
relative deviation of T and flux.

  config 2   100 000 exp fits at noise 0.01 and 1.0
  config 4   20 000 bdf fits on 64x64 stamps (PriorBDFSep, lmdif), plus the
             noise floor of the comparison: how often the SAME oracle
             compiled with fused multiply-adds disagrees with itself
  config 3   100 000 EM 3-gaussian fits and 100 000 adaptive moments

    python profiles/parity_sweep2.py [tag]   ->  gpurun_out/parity_sweep_<tag>.json
"""
import ctypes as C
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch  # noqa: E402
import oracle  # noqa: E402
import bench_configs  # noqa: E402
from ngmix_b200 import _lib, sims_device  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
_lib.init(0)
oracle.build()
dev = torch.device("cuda", 0)
cores = os.cpu_count()
out = {"cores": cores}


def lm_device(model, t, n, dims, prior=None, analytic=True):
    d_res = torch.zeros(n * _lib.LM_RESULT_DTYPE.itemsize, dtype=torch.uint8,
                        device=dev)
    _lib.lm_fit_device(model, n, t["images"].data_ptr(),
                       t["weights"].data_ptr(), t["jac"].data_ptr(), dims[0],
                       dims[1], t["psf"].data_ptr(), 3, t["guess"].data_ptr(),
                       d_res.data_ptr(), prior=prior,
                       analytic_jacobian=analytic)
    torch.cuda.synchronize()
    return np.frombuffer(d_res.cpu().numpy().tobytes(),
                         dtype=_lib.LM_RESULT_DTYPE)


# ---- config 2
n = 100000
for noise in (0.01, 1.0):
    t = sims_device.make_lm_batch_device(n, 5000000 + int(noise * 100), dev,
                                         noise=noise)
    res = lm_device("exp", t, n, (48, 48))
    t0 = time.perf_counter()
    ref = oracle.lm_fit_batch("exp", t["images"].cpu().numpy(),
                              t["weights"].cpu().numpy(), t["h_jac"],
                              t["h_psf"], t["h_guess"], nthreads=cores)
    dt = time.perf_counter() - t0
    o = bench_configs.lm_parity(res, ref, 6)
    o["cpu_port_fits_per_s"] = n / dt
    o["median_nfev"] = float(np.median(ref["nfev"]))
    out["config2_noise%g" % noise] = o
    print(out, flush=True)
    del t
    torch.cuda.empty_cache()

# ---- config 4
n = 20000
t = sims_device.make_bdf_batch_device(n, 4000000, dev)
res = lm_device("bdf", t, n, (64, 64), prior=bench_configs.bdf_prior(_lib),
                analytic=False)
im, wt = t["images"].cpu().numpy(), t["weights"].cpu().numpy()
t0 = time.perf_counter()
ref = oracle.lm_fit_batch("bdf", im, wt, t["h_jac"], t["h_psf"], t["h_guess"],
                          prior=oracle.make_prior("bdf"), nthreads=cores)
dt = time.perf_counter() - t0
o = bench_configs.lm_parity(res, ref, 7, Tcol=4, Fcol=6)
o["cpu_port_fits_per_s"] = n / dt
o["median_nfev"] = float(np.median(ref["nfev"]))
# noise floor: the oracle against itself with a*b+c contracted into FMAs
src = [os.path.join(ROOT, "oracle", f) for f in
       ("ora_gmix.c", "ora_minpack.c", "ora_lm.c", "ora_em_admom.c")]
so = "/tmp/liboracle_fma.so"
subprocess.check_call(["gcc", "-O2", "-fPIC", "-std=gnu11",
                       "-ffp-contract=fast", "-mfma", "-pthread", "-shared",
                       "-o", so] + src + ["-lm", "-lpthread"])
plain = oracle._lib
oracle._lib = C.CDLL(so)
nf = 4000
fma = oracle.lm_fit_batch("bdf", im[:nf], wt[:nf], t["h_jac"][:nf],
                          t["h_psf"][:nf], t["h_guess"][:nf],
                          prior=oracle.make_prior("bdf"), nthreads=cores)
oracle._lib = plain
refs = {k: ref[k][:nf] for k in ref.dtype.names}
o["fma_oracle_vs_oracle"] = bench_configs.lm_parity(fma, refs, 7, Tcol=4,
                                                    Fcol=6)
out["config4_bdf"] = o
print(out, flush=True)
del t, im, wt
torch.cuda.empty_cache()

# ---- config 3
n = 100000
t = sims_device.make_psf_batch_device(n, 3000000, dev)
d_pars = torch.zeros(n * 18, dtype=torch.float64, device=dev)
d_em = torch.zeros(n * _lib.EM_RESULT_DTYPE.itemsize, dtype=torch.uint8,
                   device=dev)
d_am = torch.zeros(n * _lib.ADMOM_RESULT_DTYPE.itemsize, dtype=torch.uint8,
                   device=dev)
_lib.em_device(n, t["images"].data_ptr(), t["weights"].data_ptr(),
               t["jac"].data_ptr(), 25, 25, t["em_guess"].data_ptr(), 3,
               d_pars.data_ptr(), d_em.data_ptr())
_lib.admom_device(n, t["images"].data_ptr(), t["weights"].data_ptr(),
                  t["jac"].data_ptr(), 25, 25, t["am_guess"].data_ptr(),
                  d_am.data_ptr())
torch.cuda.synchronize()
rem = np.frombuffer(d_em.cpu().numpy().tobytes(), dtype=_lib.EM_RESULT_DTYPE)
ram = np.frombuffer(d_am.cpu().numpy().tobytes(),
                    dtype=_lib.ADMOM_RESULT_DTYPE)
pem = d_pars.cpu().numpy().reshape(n, 3, 6)
im, wt = t["images"].cpu().numpy(), t["weights"].cpu().numpy()
t0 = time.perf_counter()
oem, opars = oracle.em_go_batch(im, wt, t["h_jac"], t["h_em_guess"],
                                nthreads=cores)
dt = time.perf_counter() - t0
good = (oem["flags"] == 0) & (rem["flags"] == 0)
rel = np.abs(pem[good] - opars[good]) / np.maximum(np.abs(opars[good]), 1e-12)
out["config3_em"] = {
    "n": n, "numiter_differ": int((oem["numiter"] != rem["numiter"]).sum()),
    "flags_differ": int((oem["flags"] != rem["flags"]).sum()),
    "max_rel_dev_pars_p_irr_icc": float(rel[:, :, [0, 3, 5]].max()),
    "frac_maxiter": float(np.mean(oem["flags"] == 32)),
    "cpu_port_fits_per_s": n / dt}
print(out, flush=True)
t0 = time.perf_counter()
oam = oracle.admom_go_batch(im, wt, t["h_jac"], t["h_am_guess"],
                            nthreads=cores)
dt = time.perf_counter() - t0
good = (oam["flags"] == 0) & (ram["flags"] == 0)
rel = np.abs(ram["pars"][good] - oam["pars"][good]) \
    / np.maximum(np.abs(oam["pars"][good]), 1e-12)
out["config3_admom"] = {
    "n": n, "numiter_differ": int((oam["numiter"] != ram["numiter"]).sum()),
    "flags_differ": int((oam["flags"] != ram["flags"]).sum()),
    "max_rel_dev_T_flux": float(rel[:, 4:6].max()),
    "cpu_port_fits_per_s": n / dt}
print(out, flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out",
                                 "parity_sweep_%s.json" % tag), "w"),
          indent=1)
