"""streamed host pipeline: a call of several batches (70 000 stamps), and two
host threads fitting at once on one GPU (pinned and pageable stamps)"""
import os, sys, time, threading
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from ngmix_b200 import _lib
import ngmix_b200 as ng
_lib.init(0)
n = 10000
b = bench.make_inputs(n, 1)
rep = 7
big = {k: np.concatenate([b[k]] * rep) for k in ("images", "weights", "jac", "psf", "guess")}
im32, wt32 = big["images"].astype("f4"), big["weights"].astype("f4")
os.environ["NGX_NO_STREAM"] = "1"
want1 = _lib.lm_fit("exp", b["images"].astype("f4"), b["weights"].astype("f4"), b["jac"], b["psf"], b["guess"])
os.environ["NGX_NO_STREAM"] = "0"
for label, a, w in (("pageable", im32, wt32), ("pinned", ng.pinned_copy(im32), ng.pinned_copy(wt32))):
    _lib.lm_fit("exp", a, w, big["jac"], big["psf"], big["guess"])
    t0 = time.perf_counter()
    r = _lib.lm_fit("exp", a, w, big["jac"], big["psf"], big["guess"])
    dt = time.perf_counter() - t0
    ok = all(r[i * n:(i + 1) * n].tobytes() == want1.tobytes() for i in range(rep))
    print("%s 70000 stamps (3 batches): %.3f M fits/s, records identical to the chunked path: %s" % (label, rep * n / dt / 1e6, ok), flush=True)
    assert ok
# two host threads at once
res = {}
def work(tag, a, w):
    _lib.init(0)
    for _ in range(6):
        res[tag] = _lib.lm_fit("exp", a, w, b["jac"], b["psf"], b["guess"])
p32, pw32 = ng.pinned_copy(b["images"].astype("f4")), ng.pinned_copy(b["weights"].astype("f4"))
th = [threading.Thread(target=work, args=("pinned", p32, pw32)),
      threading.Thread(target=work, args=("pageable", b["images"].astype("f4"), b["weights"].astype("f4")))]
t0 = time.perf_counter()
for t in th: t.start()
for t in th: t.join()
dt = time.perf_counter() - t0
for tag in res:
    assert res[tag].tobytes() == want1.tobytes(), tag
print("two host threads, 6 calls each (buffers allocated in the first): %.3f M fits/s in total, records identical" % (12 * n / dt / 1e6), flush=True)
print("OK")
