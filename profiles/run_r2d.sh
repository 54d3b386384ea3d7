set -x
TAG=${1:-r2d}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/gputest_$TAG.log 2>&1; tail -5 gpurun_out/gputest_$TAG.log
timeout 900 python bench.py --configs 2b,4 --nbdf 200000 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; tail -c 1000 gpurun_out/bench_${TAG}.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_${TAG}.json"))
print("value",d["value"],"e2e",d["e2e"]["value"],"e2e32",d["e2e_f32_stamps"]["value"],"frac",d["roofline"]["frac"])
print("parity",d.get("parity_on_sample"))
for k,v in d["configs"].items(): print(k, {a:b for a,b in v.items() if a!="workload"})
PY
timeout 600 python - <<PY
import sys, time, numpy as np
sys.path.insert(0,".")
import bench_extra, ngmix_b200
from ngmix_b200 import _lib, sims
_lib.init(0)
b = sims.make_psf_batch(20000, seed0=424242)
class A: pass
print("coellip", bench_extra.bench_coellip(A(), _lib, b))
PY
