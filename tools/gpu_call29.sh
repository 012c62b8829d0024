#!/bin/bash
mkdir -p gpurun_out
tag=c29
timeout 900 python -m pytest tests/test_gpu_smoother.py -q -m gpu --timeout 600 -x > gpurun_out/${tag}_tests.log 2>&1; echo "smoother tests (lower-triangle GEMM) rc=$?"; tail -2 gpurun_out/${tag}_tests.log
timeout 300 python tools/gemm_bench.py 100 192 3 2>&1 | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_dgemm -c 800 --csv --log-file gpurun_out/${tag}_gemm_pipe.csv python tools/gemm_bench.py 100 192 2 > gpurun_out/${tag}_ncu1.log 2>&1; echo "ncu pipe rc=$?"
python - <<P
import csv
rows=[r for r in csv.reader(l for l in open("gpurun_out/${tag}_gemm_pipe.csv") if l.startswith('"'))]
h=rows[0]; ik=h.index("Kernel Name"); iv=h.index("Metric Value")
v=[float(r[iv].replace(",","")) for r in rows[1:]]
print("launches", len(v), "total ms", round(sum(v)/1e6,2))
P
