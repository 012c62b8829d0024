#!/bin/bash
# final single-GPU evidence run: whole -m gpu suite, sanitizer passes on small invocations, launch list, bench lines
mkdir -p gpurun_out
tag=${1:-fin}
timeout 2400 python -m pytest tests -q -m gpu --timeout 1200 > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$?"
tail -6 gpurun_out/${tag}_tests.log
for tool in ${SAN_TOOLS:-memcheck racecheck synccheck}; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python tools/sanitize_small.py ${SAN_WHAT:-all} > gpurun_out/${tag}_san_$tool.log 2>&1; echo "sanitizer $tool rc=$?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|ok " gpurun_out/${tag}_san_$tool.log | tail -12
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-smoother --e2e-steps 8 > gpurun_out/${tag}_ncu_bench.log 2>&1; echo "launch list rc=$?"
timeout 900 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 40 --warmup 3 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err; echo "bench ref rc=$?"
python - <<P
import json
d=json.loads([l for l in open("gpurun_out/${tag}_bench.json").read().splitlines() if l.startswith("{")][-1])
print("value", round(d["value"]), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"],3), "on traffic", round(d["roofline"]["achieved_on_traffic"]), d["clocks"])
print("smoother", d.get("smoother"))
print("cpu", d.get("cpu_baseline"))
r=json.loads([l for l in open("gpurun_out/${tag}_bench_ref.json").read().splitlines() if l.startswith("{")][-1])
print("ref", r["value"], r["cpu_baseline"]["cores"])
P
