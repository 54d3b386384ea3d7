set -x
TAG=${1:-r2h}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/gputest_$TAG.log 2>&1; tail -15 gpurun_out/gputest_$TAG.log
( time timeout 1200 python profiles/parity_sweep2.py $TAG ) > gpurun_out/sweep_$TAG.log 2>&1; tail -3 gpurun_out/sweep_$TAG.log | cut -c1-3000
