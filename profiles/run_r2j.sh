set -x
TAG=${1:-r2j}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/gputest_$TAG.log 2>&1; tail -12 gpurun_out/gputest_$TAG.log
timeout 600 python - <<PY
import sys, time, numpy as np
sys.path.insert(0,".")
import torch, bench_extra
from ngmix_b200 import _lib, sims
_lib.init(0)
class A: steps=3; no_cpu=True
print("multi", bench_extra.bench_multi(A(), _lib, sims, torch, torch.device("cuda",0)))
PY
