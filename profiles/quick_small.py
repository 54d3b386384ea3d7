import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from ngmix_b200 import _lib
import ngmix_b200 as ng
_lib.init(0)
b = bench.make_inputs(4000, 1)
for n in (1, 8, 100, 600, 2000, 4000):
    sl = slice(0, n)
    args = [np.ascontiguousarray(b[k][sl]) for k in ("images", "weights", "jac", "psf", "guess")]
    pargs = [ng.pinned_copy(a) for a in args]
    for label, aa in (("pageable", args), ("pinned", pargs)):
        line = "n=%5d %-8s" % (n, label)
        for mode in ("1", "0"):
            os.environ["NGX_NO_STREAM"] = mode
            for _ in range(5):
                _lib.lm_fit("exp", *aa)
            reps = 200 if n <= 100 else 30
            t0 = time.perf_counter()
            for _ in range(reps):
                _lib.lm_fit("exp", *aa)
            dt = (time.perf_counter() - t0) / reps
            line += "  NO_STREAM=%s %8.1f us/call" % (mode, dt * 1e6)
        print(line, flush=True)
