"""a small streamed host call of each format, for compute-sanitizer"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["NGX_STREAM_ALWAYS"] = "1"
import ngmix_b200 as ng
from ngmix_b200 import _lib, sims
_lib.init(0)
b = sims.make_lm_batch(400, seed0=1, noise=0.05)
pin = ng.pinned_copy
im32, wt32 = pin(b["images"].astype("f4")), pin(b["weights"].astype("f4"))
small = (pin(b["jac"]), pin(b["psf"]), pin(b["guess"]))
r = _lib.lm_fit("exp", im32, wt32, *small)
mask = np.random.RandomState(5).uniform(size=im32.shape) > 0.03
r2 = _lib.lm_fit_compact("exp", im32, pin(b["weights"][:, 0, 0].copy()), *small, mask=mask)
r3 = _lib.lm_fit("exp", b["images"], b["weights"], b["jac"], b["psf"], b["guess"])
print("flags == 0:", int((r["flags"] == 0).sum()), int((r2["flags"] == 0).sum()), int((r3["flags"] == 0).sum()))
