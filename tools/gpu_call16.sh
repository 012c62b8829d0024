#!/bin/bash
# K7 restructured (k_chol_inv): parity first, then speed against k_chol_solve, then one full ncu capture
mkdir -p gpurun_out
tag=c16
timeout 900 python -m pytest tests/test_gpu_smoother.py tests/test_gpu_packed.py "tests/test_gpu_kernels.py" -q -m gpu --timeout 600 -x > gpurun_out/${tag}_tests.log 2>&1; echo "tests (k_chol_inv default) rc=$?"
tail -5 gpurun_out/${tag}_tests.log
RBSLAM_CHOL_KERNEL=solve timeout 900 python -m pytest tests/test_gpu_smoother.py -q -m gpu --timeout 600 -k "ancestor_weights or information" > gpurun_out/${tag}_tests_solve.log 2>&1; echo "tests (k_chol_solve) rc=$?"
tail -2 gpurun_out/${tag}_tests_solve.log
for k in inv solve; do
  RBSLAM_CHOL_KERNEL=$k timeout 300 python tools/chol_bench.py 4096 10 2>&1 | tail -1
  RBSLAM_CHOL_KERNEL=$k timeout 300 python tools/chol_bench.py 100 10 2>&1 | tail -1
done
timeout 600 python bench.py --steps 60 --no-smoother --no-cpu-baseline --e2e-steps 40 2>/dev/null | python -c "
import sys,json
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); print('bench', round(d['value']), d['roofline']['phases_ms_per_step'])"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_chol_inv -s 6 -c 1 -o gpurun_out/${tag}_chol_inv python tools/chol_bench.py 2048 3 > gpurun_out/${tag}_ncu.log 2>&1; echo "ncu rc=$?"
