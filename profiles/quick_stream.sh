#!/bin/bash
mkdir -p gpurun_out
nproc
for t in default 2 4 8 12; do
  if [ $t = default ]; then unset NGX_STAGE_THREADS; else export NGX_STAGE_THREADS=$t; fi
  timeout 300 python profiles/quick_pageable.py 2>&1 | grep -v Warn | tail -5
done | tee gpurun_out/pageable_quick.log
unset NGX_STAGE_THREADS
timeout 900 python -m pytest tests/test_gpu_batch.py -x -q -m gpu 2>&1 | tail -3
