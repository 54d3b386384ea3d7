import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ngmix_b200 as ngmix
from ngmix_b200 import sims, _lib
_lib.init(0)
nobj = 2000
b = sims.make_lm_batch(4 * nobj, seed0=31000, noise=0.02)
starts = np.arange(0, 4 * nobj + 1, 4)
bands = np.tile(np.array([0, 0, 1, 1], dtype="i4"), nobj)
g = b["guess"][::4]
guesses = np.column_stack([g, g[:, 5]])
pim, pwt = ngmix.pinned_copy(b["images"]), ngmix.pinned_copy(b["weights"])
for label, im, wt in (("pageable", b["images"], b["weights"]), ("pinned", pim, pwt)):
    for rep in range(5):
        t0 = time.perf_counter()
        r = _lib.lm_fit_multi("exp", starts, bands, 2, im, wt, b["jac"], b["psf"], guesses)
        t1 = time.perf_counter()
        r2 = ngmix.fit_batch_multi("exp", starts, bands, 2, im, wt, b["jac"], b["psf"], guesses)
        t2 = time.perf_counter()
        print("%s _lib.lm_fit_multi %.1f ms   fit_batch_multi %.1f ms" % (label, (t1 - t0) * 1e3, (t2 - t1) * 1e3), flush=True)
