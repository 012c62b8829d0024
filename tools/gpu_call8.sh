#!/bin/bash
mkdir -p gpurun_out
tag=c8
timeout 900 python -m pytest tests/test_gpu_smoother.py -q -m gpu --timeout 600 > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/${tag}_tests.log
for nt in 128 256; do
  RBSLAM_CHOL_THREADS=$nt timeout 300 python tools/chol_bench.py 4096 10 2>&1 | tail -1
  RBSLAM_CHOL_THREADS=$nt timeout 300 python tools/chol_bench.py 100 10 2>&1 | tail -1
done
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_chol_solve -s 3 -c 1 -o gpurun_out/${tag}_chol python tools/chol_bench.py 2048 4 > gpurun_out/${tag}_ncu.log 2>&1; echo "ncu rc=$?"
