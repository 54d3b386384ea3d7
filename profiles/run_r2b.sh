set -x
TAG=${1:-r2b}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/gputest_$TAG.log 2>&1; tail -5 gpurun_out/gputest_$TAG.log
timeout 600 python bench.py --configs none --no-cpu > gpurun_out/bench_${TAG}_main.json 2> gpurun_out/bench_${TAG}_main.err; tail -c 1500 gpurun_out/bench_${TAG}_main.err; python - <<PY
import json
d=json.load(open("gpurun_out/bench_${TAG}_main.json"))
print("value",d["value"],"e2e",d["e2e"]["value"],"e2e32",d["e2e_f32_stamps"]["value"],"frac",d["roofline"]["frac"],"sust",d["sustained"]["fits_per_s"])
PY
timeout 600 compute-sanitizer --tool racecheck python __graft_entry__.py smoke > gpurun_out/racecheck_$TAG.log 2>&1; tail -4 gpurun_out/racecheck_$TAG.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_lm_fit -s 3 -c 1 -f -o gpurun_out/prof_lm_$TAG python bench.py --steps 2 --warmup 3 --no-cpu --configs none --sustain-seconds 0 > /dev/null 2>&1
python profiles/ncu_summary.py gpurun_out/prof_lm_$TAG.ncu-rep > gpurun_out/ncu_lm_fit_$TAG.txt 2>&1
python profiles/ncu_lines.py gpurun_out/prof_lm_$TAG.ncu-rep > gpurun_out/ncu_lm_fit_lines_$TAG.txt 2>&1
head -50 gpurun_out/ncu_lm_fit_$TAG.txt
if [ "$2" = "kernels" ]; then
timeout 900 ncu --set full --clock-control none --profile-from-start off -f -o gpurun_out/prof_kernels_$TAG python profiles/capture_kernels.py > gpurun_out/capture_$TAG.log 2>&1; tail -2 gpurun_out/capture_$TAG.log
python profiles/ncu_multi.py gpurun_out/prof_kernels_$TAG.ncu-rep > gpurun_out/ncu_kernels_$TAG.txt 2>&1
rm -f gpurun_out/prof_kernels_$TAG.ncu-rep
fi
ls -la gpurun_out; du -sh gpurun_out
