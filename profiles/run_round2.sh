#!/bin/bash
# Produces the round-2 artifacts under gpurun_out/ (copied into profiles/ by
# hand): GPU tests, the bench lines of both arms, the ncu launch list, the
# full ncu capture of k_lm_fit and of every other kernel, the secondary
# measurements and the sanitizer logs.
#   gpurun --timeout 2400 -- 'bash profiles/run_round2.sh r2'
set -x
TAG=${1:-r2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader; nproc
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/gputest_$TAG.log 2>&1; tail -4 gpurun_out/gputest_$TAG.log
( time timeout 900 python bench.py ) > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -c 600 gpurun_out/bench_$TAG.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err
timeout 300 python bench.py --impl reference --kind port --steps 3 --warmup 1 > gpurun_out/bench_refport_$TAG.json 2>> gpurun_out/bench_ref_$TAG.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu --configs none --sustain-seconds 0 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_lm_fit -s 3 -c 1 -f -o gpurun_out/prof_lm_$TAG python bench.py --steps 2 --warmup 3 --no-cpu --configs none --sustain-seconds 0 > /dev/null 2>&1
python profiles/ncu_summary.py gpurun_out/prof_lm_$TAG.ncu-rep > gpurun_out/ncu_lm_fit_$TAG.txt 2>&1
python profiles/ncu_lines.py gpurun_out/prof_lm_$TAG.ncu-rep > gpurun_out/ncu_lm_fit_lines_$TAG.txt 2>&1
timeout 900 ncu --set full --clock-control none --profile-from-start off -f -o gpurun_out/prof_kernels_$TAG python profiles/capture_kernels.py > gpurun_out/capture_$TAG.log 2>&1
python profiles/ncu_multi.py gpurun_out/prof_kernels_$TAG.ncu-rep > gpurun_out/ncu_kernels_$TAG.txt 2>&1
rm -f gpurun_out/prof_kernels_$TAG.ncu-rep
timeout 900 python bench_extra.py > gpurun_out/bench_extra_$TAG.json 2> gpurun_out/bench_extra_$TAG.err; tail -c 300 gpurun_out/bench_extra_$TAG.err
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_chain.py tests/test_lm_bounds.py tests/test_lm_multi.py -m gpu -x -q > gpurun_out/sanitizer_memcheck_$TAG.log 2>&1; tail -3 gpurun_out/sanitizer_memcheck_$TAG.log
timeout 600 compute-sanitizer --tool racecheck python -m pytest tests/test_chain.py tests/test_lm_bounds.py "tests/test_lm_multi.py::test_gpu_multi_fd_matches_reference" -m gpu -x -q > gpurun_out/sanitizer_racecheck_$TAG.log 2>&1; tail -3 gpurun_out/sanitizer_racecheck_$TAG.log
timeout 300 compute-sanitizer --tool racecheck python __graft_entry__.py smoke > gpurun_out/sanitizer_racecheck_smoke_$TAG.log 2>&1; tail -2 gpurun_out/sanitizer_racecheck_smoke_$TAG.log
ls -la gpurun_out; du -sh gpurun_out
