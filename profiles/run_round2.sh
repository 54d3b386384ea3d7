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
NGX_NO_STREAM=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu --configs none --sustain-seconds 0 > /dev/null 2>&1
# the same command with the host calls streaming (pinned batches queue their copies before the launch, pageable ones fall back to chunks under the profiler)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_e2e_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu --configs none --sustain-seconds 0 > /dev/null 2>&1
NGX_NO_STREAM=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_lm_fit -s 3 -c 1 -f -o gpurun_out/prof_lm_$TAG python bench.py --steps 2 --warmup 3 --no-cpu --configs none --sustain-seconds 0 > /dev/null 2>&1
python profiles/ncu_summary.py gpurun_out/prof_lm_$TAG.ncu-rep > gpurun_out/ncu_lm_fit_$TAG.txt 2>&1
python profiles/ncu_lines.py gpurun_out/prof_lm_$TAG.ncu-rep > gpurun_out/ncu_lm_fit_lines_$TAG.txt 2>&1
timeout 900 ncu --set full --clock-control none --profile-from-start off -f -o gpurun_out/prof_kernels_$TAG python profiles/capture_kernels.py > gpurun_out/capture_$TAG.log 2>&1
python profiles/ncu_multi.py gpurun_out/prof_kernels_$TAG.ncu-rep > gpurun_out/ncu_kernels_$TAG.txt 2>&1
rm -f gpurun_out/prof_kernels_$TAG.ncu-rep
timeout 400 ncu --set full --clock-control none -k regex:k_admom -s 1 -c 1 -f -o gpurun_out/prof_admom_$TAG python profiles/capture_admom.py > /dev/null 2>&1
python profiles/ncu_multi.py gpurun_out/prof_admom_$TAG.ncu-rep > gpurun_out/ncu_admom_$TAG.txt 2>&1
timeout 900 python bench_extra.py > gpurun_out/bench_extra_$TAG.json 2> gpurun_out/bench_extra_$TAG.err; tail -c 300 gpurun_out/bench_extra_$TAG.err
# the host-buffer pipelines: streamed vs chunked, pageable arrays, several batches / two host threads
timeout 600 python profiles/stream_vs_chunked.py > gpurun_out/stream_vs_chunked_$TAG.txt 2>&1; tail -3 gpurun_out/stream_vs_chunked_$TAG.txt
timeout 300 python profiles/pageable_sweep.py > gpurun_out/pageable_$TAG.txt 2>&1
for t in 2 4 8 12; do NGX_STAGE_THREADS=$t timeout 300 python profiles/pageable_sweep.py >> gpurun_out/pageable_$TAG.txt 2>&1; done
timeout 600 python profiles/stress_stream.py > gpurun_out/stress_stream_$TAG.txt 2>&1; tail -2 gpurun_out/stress_stream_$TAG.txt
timeout 300 python profiles/multi_host_timing.py > gpurun_out/multi_host_$TAG.txt 2>&1
# parity at scale of the kernels changed in this tag
timeout 600 python profiles/admom_sweep.py > gpurun_out/admom_sweep_$TAG.json 2> gpurun_out/admom_sweep_$TAG.err
timeout 600 python profiles/coellip_parity.py 2>/dev/null | tail -1 > gpurun_out/coellip_parity_$TAG.json
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_chain.py tests/test_lm_bounds.py tests/test_lm_multi.py "tests/test_gpu_batch.py::test_streamed_pipeline_failure_releases_the_device" "tests/test_gpu_batch.py::test_em_admom_host_calls_pipelined" -m gpu -x -q > gpurun_out/sanitizer_memcheck_$TAG.log 2>&1; tail -3 gpurun_out/sanitizer_memcheck_$TAG.log
timeout 600 compute-sanitizer --tool racecheck python -m pytest tests/test_chain.py tests/test_lm_bounds.py "tests/test_lm_multi.py::test_gpu_multi_fd_matches_reference" -m gpu -x -q > gpurun_out/sanitizer_racecheck_$TAG.log 2>&1; tail -3 gpurun_out/sanitizer_racecheck_$TAG.log
for tool in memcheck racecheck; do timeout 600 compute-sanitizer --tool $tool python profiles/sanitize_stream.py > gpurun_out/sanitizer_stream_${tool}_$TAG.log 2>&1; tail -2 gpurun_out/sanitizer_stream_${tool}_$TAG.log; done
timeout 300 compute-sanitizer --tool racecheck python __graft_entry__.py smoke > gpurun_out/sanitizer_racecheck_smoke_$TAG.log 2>&1; tail -2 gpurun_out/sanitizer_racecheck_smoke_$TAG.log
ls -la gpurun_out; du -sh gpurun_out
