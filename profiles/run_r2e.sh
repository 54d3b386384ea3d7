set -x
TAG=${1:-r2e}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/gputest_$TAG.log 2>&1; tail -25 gpurun_out/gputest_$TAG.log
timeout 600 python - <<PY
import sys, time, numpy as np
sys.path.insert(0,".")
import bench_extra, ngmix_b200
from ngmix_b200 import _lib, sims
_lib.init(0)
b = sims.make_psf_batch(20000, seed0=424242)
class A: pass
for k in range(2):
    print("coellip", bench_extra.bench_coellip(A(), _lib, b))
PY
