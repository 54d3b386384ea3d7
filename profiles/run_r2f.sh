set -x
TAG=${1:-r2f}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/gputest_$TAG.log 2>&1; tail -25 gpurun_out/gputest_$TAG.log
timeout 900 python bench.py --configs 1 --no-cpu > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; tail -c 1000 gpurun_out/bench_${TAG}.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_${TAG}.json"))
print("value",d["value"],"e2e",d["e2e"]["value"],"e2e32",d["e2e_f32_stamps"]["value"],"compact",d["e2e_compact"]["value"],"pageable",d["e2e_pageable"]["value"],"frac",d["roofline"]["frac"])
for k,v in d["configs"].items(): print(k, {a:b for a,b in v.items() if a!="workload"})
PY
