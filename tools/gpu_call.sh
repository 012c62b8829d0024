#!/bin/bash
# one gpurun call: parity suite, then the bench lines of this round's kernels.  Everything
# is wrapped in `timeout` so that a hung kernel cannot hold the box.
mkdir -p gpurun_out
tag=${1:-c}
timeout 900 python -m pytest tests/test_gpu_packed.py -x -q -m gpu --timeout 600 > gpurun_out/${tag}_packed.log 2>&1; echo "packed rc=$?"
tail -5 gpurun_out/${tag}_packed.log
timeout 1500 python -m pytest tests -q -m gpu --timeout 900 --deselect tests/test_gpu_packed.py > gpurun_out/${tag}_all.log 2>&1; echo "all rc=$?"
tail -8 gpurun_out/${tag}_all.log
for v in 0 7; do
  timeout 600 python bench.py --variant $v --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_v$v.json 2> gpurun_out/${tag}_bench_v$v.err; echo "bench v$v rc=$?"
  python - <<P
import json
try:
    d=json.load(open("gpurun_out/${tag}_bench_v$v.json"))
    print("v$v", round(d["value"]), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"]), "kalman", round(d["roofline"]["kernel_ms_per_step"],2), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e: print("v$v failed", e)
P
done
