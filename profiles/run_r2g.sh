set -x
TAG=${1:-r2g}
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/gputest_$TAG.log 2>&1; tail -40 gpurun_out/gputest_$TAG.log
