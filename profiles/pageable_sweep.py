"""pageable numpy stamps: chunked pipeline vs the streamed one with the
parallel staging copy (NGX_STAGE_THREADS is read once per process)"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from ngmix_b200 import _lib
n = 10000
_lib.init(0)
b = bench.make_inputs(n, 1)
im, wt = b["images"].copy(), b["weights"].copy()
im32, wt32 = im.astype(np.float32), wt.astype(np.float32)


def run(kind):
    if kind == "f64":
        return _lib.lm_fit("exp", im, wt, b["jac"], b["psf"], b["guess"])
    return _lib.lm_fit("exp", im32, wt32, b["jac"], b["psf"], b["guess"])


def timed(kind, reps=5):
    run(kind)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        run(kind)
    torch.cuda.synchronize()
    return n * reps / (time.perf_counter() - t0)


res = {}
for mode in ("1", "0"):
    os.environ["NGX_NO_STREAM"] = mode
    for kind in ("f64", "f32"):
        res[(mode, kind)] = run(kind).copy()
        print("threads=%s NO_STREAM=%s pageable %s %.3f M fits/s" % (
            os.environ.get("NGX_STAGE_THREADS", "default"), mode, kind,
            timed(kind) / 1e6), flush=True)
for kind in ("f64", "f32"):
    assert res[("1", kind)].tobytes() == res[("0", kind)].tobytes()
print("identical", flush=True)
