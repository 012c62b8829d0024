#!/bin/bash
mkdir -p gpurun_out
tag=c17
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_chol_inv -s 2 -c 1 -o gpurun_out/${tag}_chol_inv python tools/chol_bench.py 2048 3 > gpurun_out/${tag}_ncu.log 2>&1; echo "ncu rc=$?"
timeout 600 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_filter.py -q -m gpu --timeout 600 -x > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/${tag}_tests.log
timeout 600 python bench.py --steps 60 --no-smoother --no-cpu-baseline --e2e-steps 40 2>/dev/null | python -c "
import sys,json
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); print('bench', round(d['value']), d['roofline']['phases_ms_per_step'])"
