#!/bin/bash
mkdir -p gpurun_out
tag=c31
timeout 900 python -m pytest tests/test_gpu_smoother.py -q -m gpu --timeout 600 -x > gpurun_out/${tag}_tests.log 2>&1; echo "smoother tests rc=$?"; tail -2 gpurun_out/${tag}_tests.log
for wm in 160 0; do
  RBSLAM_CHOL_WIDE_MAX=$wm timeout 300 python tools/chol_bench.py 100 10 2>&1 | tail -1 | cut -c1-220
  RBSLAM_CHOL_WIDE_MAX=$wm timeout 300 python tools/gemm_bench.py 100 192 3 2>&1 | tail -1 | cut -c1-260
done
timeout 300 python tools/chol_bench.py 4096 10 2>&1 | tail -1 | cut -c1-220
