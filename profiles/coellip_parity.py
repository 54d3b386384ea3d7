"""coelliptical fits, 1 and 2 gaussians: evaluation counts of the CUDA kernel
against the oracle, beside the oracle's own FMA noise floor"""
import os, sys, subprocess, ctypes as C, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle as ora
from ngmix_b200 import sims, _lib
import ngmix_b200 as ngmix
_lib.init(0)
src = [os.path.join(ROOT, "oracle", f) for f in ("ora_gmix.c", "ora_minpack.c", "ora_lm.c", "ora_em_admom.c")]
so = "/tmp/liboracle_fma.so"
subprocess.check_call(["gcc", "-O2", "-fPIC", "-std=gnu11", "-ffp-contract=fast", "-mfma", "-pthread", "-shared", "-o", so] + src + ["-lm", "-lpthread"])
fma = C.CDLL(so)
n = 2000
b = sims.make_psf_batch(n, seed0=424242)
T = ngmix.moments.fwhm_to_T(0.263 * 3.5)
flux = b["images"].sum(axis=(1, 2))
out = {}
for ngauss in (1, 2):
    rng = np.random.RandomState(11)
    guess = np.zeros((n, 4 + 2 * ngauss))
    for i in range(n):
        g = np.zeros(4 + 2 * ngauss)
        g[0:2] = rng.uniform(-0.01, 0.01, 2); g[2:4] = rng.uniform(-0.05, 0.05, 2)
        fr = ngmix.guessers._COELLIP_F[ngauss]; pr = ngmix.guessers._COELLIP_P[ngauss]
        g[4:4 + ngauss] = T * np.array(fr) * rng.uniform(0.99, 1.01, ngauss)
        g[4 + ngauss:] = flux[i] * np.array(pr) * rng.uniform(0.99, 1.01, ngauss)
        guess[i] = g
    r = _lib.lm_fit_coellip(ngauss, b["images"], b["weights"], b["jac"], None, guess)
    res = {}
    for name, lib in (("plain", ora.lib()), ("fma", fma)):
        rr = np.zeros(n, dtype=ora.LM_RESULT_DTYPE)
        for i in range(n):
            prior = ora.make_prior("none"); prior.ncoellip = ngauss
            im, wt = b["images"][i], b["weights"][i]; jac = b["jac"][i]; gg = guess[i]
            lib.ora_lm_fit(C.c_int(ora.MODELS["coellip"]), im.ctypes.data_as(C.c_void_p), wt.ctypes.data_as(C.c_void_p), C.c_int(im.shape[0]), C.c_int(im.shape[1]), jac.ctypes.data_as(C.c_void_p), None, C.c_int(0), gg.ctypes.data_as(C.c_void_p), C.byref(prior), C.c_int(4000), C.c_double(1e-5), C.c_double(1e-5), C.c_int(1), C.c_int(1), rr[i:i + 1].ctypes.data_as(C.c_void_p))
        res[name] = rr
    np_ = 4 + 2 * ngauss
    o = res["plain"]
    good = (o["flags"] == 0) & (r["flags"] == 0)
    tol = np.maximum(1e-5 * np.abs(o["pars"][:, :np_]), 1e-3 * o["pars_err"][:, :np_])
    out["ngauss%d" % ngauss] = {
        "n": n,
        "gpu_vs_oracle_nfev_equal": int((r["nfev"] == o["nfev"]).sum()),
        "gpu_vs_oracle_flags_equal": int((r["flags"] == o["flags"]).sum()),
        "gpu_vs_oracle_pars_within_tol": int((np.abs(r["pars"][:, :np_] - o["pars"][:, :np_]) <= tol).all(axis=1)[good].sum()),
        "n_good": int(good.sum()),
        "fma_oracle_vs_oracle_nfev_equal": int((res["fma"]["nfev"] == o["nfev"]).sum()),
        "fma_oracle_vs_oracle_pars_within_tol": int((np.abs(res["fma"]["pars"][:, :np_] - o["pars"][:, :np_]) <= tol).all(axis=1)[good].sum()),
        "gpu_vs_oracle_pars_within_0.05sigma": int((np.abs(r["pars"][:, :np_] - o["pars"][:, :np_]) <= np.maximum(1e-5 * np.abs(o["pars"][:, :np_]), 0.05 * o["pars_err"][:, :np_])).all(axis=1)[good].sum()),
        "fma_oracle_vs_oracle_pars_within_0.05sigma": int((np.abs(res["fma"]["pars"][:, :np_] - o["pars"][:, :np_]) <= np.maximum(1e-5 * np.abs(o["pars"][:, :np_]), 0.05 * o["pars_err"][:, :np_])).all(axis=1)[good].sum()),
        "max_rel_dev_lnprob": float(np.max(np.abs(r["lnprob"][good] / o["lnprob"][good] - 1)))}
print(json.dumps(out), flush=True)
