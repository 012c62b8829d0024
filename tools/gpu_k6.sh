#!/bin/bash
mkdir -p gpurun_out
tag=${1:-c24}
timeout 900 python -m pytest tests/test_gpu_smoother.py -q -m gpu --timeout 600 -x > gpurun_out/${tag}_tests.log 2>&1; echo "smoother tests (pipelined GEMM) rc=$?"; tail -2 gpurun_out/${tag}_tests.log
timeout 300 python tools/gemm_bench.py 100 192 3 2>&1 | tail -1
RBSLAM_GEMM_SYNC=1 timeout 300 python tools/gemm_bench.py 100 192 3 2>&1 | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_dgemm -c 800 --csv --log-file gpurun_out/${tag}_gemm_pipe.csv python tools/gemm_bench.py 100 192 2 > gpurun_out/${tag}_ncu1.log 2>&1; echo "ncu pipe rc=$?"
RBSLAM_GEMM_SYNC=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_dgemm -c 800 --csv --log-file gpurun_out/${tag}_gemm_sync.csv python tools/gemm_bench.py 100 192 2 > gpurun_out/${tag}_ncu2.log 2>&1; echo "ncu sync rc=$?"
python - <<P
import csv
for name in ("pipe","sync"):
    rows=[r for r in csv.reader(l for l in open("gpurun_out/${tag}_gemm_%s.csv"%name) if l.startswith('"'))]
    h=rows[0]; ik=h.index("Kernel Name"); iv=h.index("Metric Value"); ig=h.index("Grid Size")
    tot={}; 
    for r in rows[1:]:
        k="TA" if "true" in r[ik] else "NN"
        tot.setdefault(k,[]).append((r[ig], float(r[iv].replace(",",""))))
    for k,v in tot.items():
        print(name, k, "launches", len(v), "total ms", round(sum(x[1] for x in v)/1e6,2), "largest", v[len(v)//2+1] if len(v)>2 else v[0])
P
