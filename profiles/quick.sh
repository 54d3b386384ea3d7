set -x
( timeout 900 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -5
timeout 600 python bench.py --configs none --no-cpu --sustain-seconds 0 > gpurun_out/bench_q.json 2>/dev/null
python - <<PY
import json
d=json.load(open("gpurun_out/bench_q.json"))
print("value",d["value"],"e2e",d["e2e"]["value"],"fdiff",d["fdiff_kernel"]["ms"],d["fdiff_kernel"]["frac_fp64_flops"],"h2d",d["host_h2d_ceiling"])
PY
