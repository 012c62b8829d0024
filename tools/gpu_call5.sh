#!/bin/bash
mkdir -p gpurun_out
tag=c5
timeout 900 python -m pytest tests/test_gpu_packed.py tests/test_gpu_filter.py -q -m gpu --timeout 600 -k "packed or 7" > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$?"
tail -4 gpurun_out/${tag}_tests.log
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --variant 7 --steps 12 --warmup 3 --no-cpu-baseline --no-smoother --e2e-steps 16 > gpurun_out/${tag}_$name.json 2> gpurun_out/${tag}_$name.err
  python - <<P
import json
try:
    d=json.load(open("gpurun_out/${tag}_$name.json"))
    print("$name", round(d["value"]), "ms", round(d["ms_per_step"],2), "kalman", round(d["roofline"]["kernel_ms_per_step"],2), "e2e", round(d["e2e"]["value"]), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e: print("$name failed", e); print(open("gpurun_out/${tag}_$name.err").read()[-600:])
P
}
run base X=1
run ts96ns2 RBSLAM_PT_CFG=96,2
run ts80ns3 RBSLAM_PT_CFG=80,3
run ts60ns4 RBSLAM_PT_CFG=60,4
run nsplit2 RBSLAM_NSPLIT=2
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_stream_fam_pt -s 6 -c 1 -o gpurun_out/${tag}_pt python bench.py --variant 7 --particles 2000 --steps 2 --warmup 3 --no-cpu-baseline --no-smoother --e2e-steps 8 > gpurun_out/${tag}_ncu.log 2>&1; echo "ncu rc=$?"
