#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_batch.py -x -q -m gpu 2>&1 | tail -15
# sanitizers over a small streamed call
cat > /tmp/san.py <<'PY'
import numpy as np, os, sys
sys.path.insert(0, os.getcwd())
import ngmix_b200 as ng
from ngmix_b200 import _lib, sims
_lib.init(0)
b = sims.make_lm_batch(400, seed0=1, noise=0.05)
pin = ng.pinned_copy
im32 = pin(b["images"].astype("f4")); wt32 = pin(b["weights"].astype("f4"))
r = _lib.lm_fit("exp", im32, wt32, pin(b["jac"]), pin(b["psf"]), pin(b["guess"]))
mask = np.random.RandomState(5).uniform(size=im32.shape) > 0.03
r2 = _lib.lm_fit_compact("exp", im32, pin(b["weights"][:, 0, 0].copy()), pin(b["jac"]), pin(b["psf"]), pin(b["guess"]), mask=mask)
print("flags0", int((r["flags"] == 0).sum()), int((r2["flags"] == 0).sum()))
PY
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool python /tmp/san.py > gpurun_out/sanitizer_stream_$tool.log 2>&1
  tail -4 gpurun_out/sanitizer_stream_$tool.log
done
