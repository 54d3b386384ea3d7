set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
nproc
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/gputest_r2a.log 2>&1; tail -5 gpurun_out/gputest_r2a.log
( time timeout 900 python bench.py ) > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; tail -c 3000 gpurun_out/bench_r2a.err; head -c 6000 gpurun_out/bench_r2a.json
timeout 300 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_ref_r2a.json 2> gpurun_out/bench_ref_r2a.err; cat gpurun_out/bench_ref_r2a.json
timeout 300 python bench.py --impl reference --kind port --steps 3 --warmup 1 > gpurun_out/bench_refport_r2a.json 2>> gpurun_out/bench_ref_r2a.err; cat gpurun_out/bench_refport_r2a.json
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/prof_kernels_r2a python profiles/capture_kernels.py > gpurun_out/capture_r2a.log 2>&1; tail -3 gpurun_out/capture_r2a.log
ls -la gpurun_out/*.ncu-rep
