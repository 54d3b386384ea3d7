#!/bin/bash
# the host-pipeline measurements of run_round2.sh alone
TAG=${1:-r2c}
mkdir -p gpurun_out
( time timeout 900 python bench.py ) > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -c 300 gpurun_out/bench_$TAG.err
timeout 600 python profiles/stream_vs_chunked.py > gpurun_out/stream_vs_chunked_$TAG.txt 2>&1; tail -12 gpurun_out/stream_vs_chunked_$TAG.txt
timeout 300 python profiles/pageable_sweep.py > gpurun_out/pageable_$TAG.txt 2>&1
for t in 2 4 8 12; do NGX_STAGE_THREADS=$t timeout 300 python profiles/pageable_sweep.py >> gpurun_out/pageable_$TAG.txt 2>&1; done
grep -v identical gpurun_out/pageable_$TAG.txt
timeout 600 python profiles/stress_stream.py > gpurun_out/stress_stream_$TAG.txt 2>&1; tail -4 gpurun_out/stress_stream_$TAG.txt
timeout 300 python -m pytest tests/test_gpu_batch.py -m gpu -q 2>&1 | tail -2
for tool in memcheck racecheck; do timeout 600 compute-sanitizer --tool $tool python profiles/sanitize_stream.py > gpurun_out/sanitizer_stream_${tool}_$TAG.log 2>&1; tail -2 gpurun_out/sanitizer_stream_${tool}_$TAG.log; done
