"""flags / nfev / ier / parameter agreement with the oracle per 1e5 fits
(SURVEY.md 8d 'report mismatch count per 1e5'): config-2 stamps at two noise
levels, with and without a prior"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import bench, oracle
from ngmix_b200 import _lib
_lib.init(0); oracle.build()
n = 100000
out = {}
for noise in (0.01, 1.0):
    b = bench.make_inputs(n, seed0=5000000, noise=noise)
    for pk in ("none", "flat"):
        pr = None; opr = None
        if pk == "flat":
            pr = _lib.Prior(); pr.present = 1; pr.g_kind = 1
            pr.cen1 = pr.cen2 = 0.0; pr.sigma1 = pr.sigma2 = 0.263; pr.g_sigma = 0.3
            pr.T = _lib.Prior1D(1, 0, -1.0, 1.0e3, 0, 0); pr.F = _lib.Prior1D(1, 0, -100.0, 1.0e9, 0, 0)
            opr = oracle.make_prior("flat")
        res = _lib.lm_fit("exp", b["images"], b["weights"], b["jac"], b["psf"], b["guess"], prior=pr)
        t0 = time.perf_counter()
        ref = oracle.lm_fit_batch("exp", b["images"], b["weights"], b["jac"], b["psf"], b["guess"], prior=opr, nthreads=os.cpu_count())
        dt = time.perf_counter() - t0
        ok = (ref["flags"] == 0) & (res["flags"] == 0)
        err = np.abs(res["pars"][:, :6] - ref["pars"][:, :6])
        tol = np.maximum(1e-5 * np.abs(ref["pars"][:, :6]), 1e-3 * ref["pars_err"][:, :6])
        out["noise%g_%s" % (noise, pk)] = {
            "n": n, "flags_differ": int((res["flags"] != ref["flags"]).sum()),
            "nfev_or_ier_differ": int(((res["nfev"] != ref["nfev"]) | (res["ier"] != ref["ier"])).sum()),
            "pars_outside_tol": int((~np.all(err <= tol, axis=1) & ok).sum()),
            "nonzero_flags_ref": int((ref["flags"] != 0).sum()),
            "median_nfev": float(np.median(ref["nfev"])), "cpu_fits_per_s": n / dt}
        print(out, flush=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "parity_sweep_r1m.json"), "w"), indent=1)
