"""streamed host pipeline vs the chunked one: identical records, timings"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from ngmix_b200 import _lib
import ngmix_b200 as ng

n = 10000
_lib.init(0)
b = bench.make_inputs(n, 1)
pin = ng.pinned_copy
him, hwt = pin(b["images"]), pin(b["weights"])
him32, hwt32 = pin(b["images"].astype(np.float32)), pin(b["weights"].astype(np.float32))
hws = pin(b["weights"][:, 0, 0].copy())
hjac, hpsf, hguess = pin(b["jac"]), pin(b["psf"]), pin(b["guess"])
out = ng.pinned_empty((n,), _lib.LM_RESULT_DTYPE)
mask = np.random.RandomState(3).uniform(size=b["images"].shape) > 0.02


def run(kind, out=out):
    if kind == "f64":
        return _lib.lm_fit("exp", him, hwt, hjac, hpsf, hguess, out=out)
    if kind == "f32":
        return _lib.lm_fit("exp", him32, hwt32, hjac, hpsf, hguess, out=out)
    if kind == "compact":
        return _lib.lm_fit_compact("exp", him32, hws, hjac, hpsf, hguess, out=out)
    if kind == "mask":
        return _lib.lm_fit_compact("exp", him32, hws, hjac, hpsf, hguess,
                                   mask=mask, out=out)
    if kind == "pageable_res":
        return _lib.lm_fit("exp", him32, hwt32, hjac, hpsf, hguess)


def timed(kind, reps=10):
    for _ in range(2):
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
    for kind in ("f64", "f32", "compact", "mask", "pageable_res"):
        r = run(kind).copy()
        res[(mode, kind)] = r
        print("NO_STREAM=%s %-12s %.3f M fits/s  flags!=0: %d" % (
            mode, kind, timed(kind) / 1e6, int((r["flags"] != 0).sum())),
            flush=True)
for kind in ("f64", "f32", "compact", "mask", "pageable_res"):
    a, c = res[("1", kind)], res[("0", kind)]
    same = a.tobytes() == c.tobytes()
    print(kind, "records identical:", same, flush=True)
    assert same, kind
# several batches per call and a ragged tail
os.environ["NGX_NO_STREAM"] = "0"
os.environ["NGX_STREAM_BATCH"] = "3000"
for kind in ("f32", "mask", "pageable_res"):
    r = run(kind)
    assert r.tobytes() == res[("1", kind)].tobytes(), kind
    print(kind, "4 batches identical", flush=True)
del os.environ["NGX_STREAM_BATCH"]
# bdf (forward differences) through the same path
from ngmix_b200 import sims
bb = sims.make_bdf_batch(2000, seed0=5, dims=(40, 40)) if hasattr(sims, "make_bdf_batch") else None
if bb is not None:
    args = [pin(bb[k]) for k in ("images", "weights", "jac", "psf", "guess")]
    os.environ["NGX_NO_STREAM"] = "1"
    r1 = _lib.lm_fit("bdf", *args).copy()
    os.environ["NGX_NO_STREAM"] = "0"
    r0 = _lib.lm_fit("bdf", *args).copy()
    print("bdf identical:", r0.tobytes() == r1.tobytes(), flush=True)
    assert r0.tobytes() == r1.tobytes()
print("OK")
