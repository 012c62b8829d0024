#!/bin/bash
mkdir -p gpurun_out
tag=${1:-c18}
timeout 900 python -m pytest tests/test_gpu_smoother.py -q -m gpu --timeout 600 -x > gpurun_out/${tag}_tests.log 2>&1; echo "smoother tests rc=$?"
tail -3 gpurun_out/${tag}_tests.log
timeout 300 python tools/chol_bench.py 4096 10 2>&1 | tail -1
timeout 300 python tools/chol_bench.py 100 10 2>&1 | tail -1
if [ -n "$NCU" ]; then
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_chol_inv -s 2 -c 1 -o gpurun_out/${tag}_chol_inv python tools/chol_bench.py 2048 3 > gpurun_out/${tag}_ncu.log 2>&1; echo "ncu rc=$?"
fi
