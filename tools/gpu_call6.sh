#!/bin/bash
mkdir -p gpurun_out
tag=c6
timeout 1800 python -m pytest tests -q -m gpu --timeout 900 > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$?"
tail -12 gpurun_out/${tag}_tests.log
timeout 900 python bench.py --steps 60 > gpurun_out/${tag}_bench_full.json 2> gpurun_out/${tag}_bench_full.err; echo "full bench rc=$?"; python - <<P
import json
d=json.load(open("gpurun_out/${tag}_bench_full.json"))
print(round(d["value"]), d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"].get("achieved_on_traffic"), d["smoother"], d["cpu_baseline"]["value"])
P
tail -3 gpurun_out/${tag}_bench_full.err
