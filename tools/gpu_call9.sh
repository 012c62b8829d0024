#!/bin/bash
mkdir -p gpurun_out
tag=c9
RBSLAM_CHOL_PANEL_MIN=4 timeout 900 python -m pytest tests/test_gpu_smoother.py -q -m gpu --timeout 600 > gpurun_out/${tag}_tests.log 2>&1; echo "tests(panel path forced) rc=$?"
tail -5 gpurun_out/${tag}_tests.log
timeout 300 python tools/chol_bench.py 4096 10 2>&1 | tail -1
RBSLAM_CHOL_PANEL_MIN=50 timeout 300 python tools/chol_bench.py 100 10 2>&1 | tail -1
timeout 300 python tools/chol_bench.py 1024 10 2>&1 | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_chol -s 40 -c 24 --csv --log-file gpurun_out/${tag}_chol_launches.csv python tools/chol_bench.py 4096 4 > gpurun_out/${tag}_ncu.log 2>&1; echo "ncu rc=$?"
python - <<P
import csv
rows=[r for r in csv.reader(open("gpurun_out/c9_chol_launches.csv")) if len(r)>5 and r[0].isdigit()]
for r in rows: print(r[4][:40], r[-1], r[-3] if len(r)>3 else "")
P
