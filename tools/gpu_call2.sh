#!/bin/bash
mkdir -p gpurun_out
tag=c2
timeout 900 python -m pytest tests/test_gpu_packed.py tests/test_gpu_smoother.py -q -m gpu --timeout 600 > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$?"
tail -6 gpurun_out/${tag}_tests.log
timeout 300 python tools/e2e_breakdown.py 0 24 2>&1 | tail -8
timeout 300 python tools/e2e_breakdown.py 7 24 2>&1 | tail -8
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --variant 7 --steps 12 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_$name.json 2> gpurun_out/${tag}_$name.err
  python - <<P
import json
try:
    d=json.load(open("gpurun_out/${tag}_$name.json"))
    print("$name", round(d["value"]), "ms", round(d["ms_per_step"],2), "kalman", round(d["roofline"]["kernel_ms_per_step"],2), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e: print("$name failed", e)
P
}
run base X=1
run ts32ns6 RBSLAM_PT_CFG=32,6
run ts24ns8 RBSLAM_PT_CFG=24,8
run ts64ns3 RBSLAM_PT_CFG=64,3
run ts96ns2 RBSLAM_PT_CFG=96,2
run nsplit1 RBSLAM_NSPLIT=1
run nsplit4 RBSLAM_NSPLIT=4
run nw7 RBSLAM_PT_NW=7
run nw7ns1 RBSLAM_PT_NW=7 RBSLAM_NSPLIT=1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_stream_fam_pt -s 6 -c 2 -o gpurun_out/${tag}_pt python bench.py --variant 7 --particles 2000 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/*.ncu-rep
