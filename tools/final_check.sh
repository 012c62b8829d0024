#!/bin/bash
# round-end verification on one B200: GPU test suite, smoke(), both bench arms, ncu launch list
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -6 | cut -c1-250
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -2
timeout -s KILL 600 python bench.py 2>/dev/null | tail -1 > gpurun_out/bench_final_n1.json
timeout -s KILL 600 python bench.py --impl reference 2>/dev/null | tail -1 > gpurun_out/bench_final_reference.json
python - <<'PY'
import json
for f in ("bench_final_n1", "bench_final_reference"):
    d = json.load(open("gpurun_out/%s.json" % f))
    print(f, round(d["value"], 1), round(d["ms_per_step"], 3), d.get("e2e", {}).get("value"), d.get("clocks"),
          d.get("roofline", {}).get("frac"), d.get("cpu_baseline", {}).get("value"), d.get("gpu_launches"))
PY
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
tail -30 gpurun_out/launches_final.csv | cut -d, -f5,15 | tail -14
