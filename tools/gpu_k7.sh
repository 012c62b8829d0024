#!/bin/bash
mkdir -p gpurun_out
tag=${1:-c21}
for cfg in 0 1; do
  RBSLAM_CHOL_CFG=$cfg timeout 300 python -m pytest tests/test_gpu_smoother.py -q -m gpu --timeout 300 -x -k "ancestor_weights or information or cov_teacher" > gpurun_out/${tag}_tests_$cfg.log 2>&1; echo "cfg $cfg tests rc=$?"; tail -1 gpurun_out/${tag}_tests_$cfg.log
  RBSLAM_CHOL_CFG=$cfg timeout 300 python tools/chol_bench.py 4096 10 2>&1 | tail -1 | cut -c1-200
  RBSLAM_CHOL_CFG=$cfg timeout 300 python tools/chol_bench.py 100 10 2>&1 | tail -1 | cut -c1-200
done
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_chol_inv -s 2 -c 1 -o gpurun_out/${tag}_chol_inv python tools/chol_bench.py 2048 3 > gpurun_out/${tag}_ncu.log 2>&1; echo "ncu rc=$?"
