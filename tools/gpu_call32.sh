#!/bin/bash
mkdir -p gpurun_out
tag=c32
timeout 300 python tools/chol_bench.py 100 10 2>&1 | tail -1 | cut -c1-220
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_chol_inv -s 2 -c 1 -o gpurun_out/${tag}_chol_wide python tools/chol_bench.py 100 3 > gpurun_out/${tag}_ncu.log 2>&1; echo "ncu rc=$?"
