#!/bin/bash
run() {
  echo "== $*"
  env "$@" python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); r = d['roofline']
        print('value %.0f  kalman_ms %.3f  frac %.4f  step_ms %.3f' % (d['value'], r['kernel_ms_per_step'], r['frac'], d['ms_per_step']))
    elif 'rror' in l: print(l.strip())
"
}
run RBSLAM_NSPLIT=4
run RBSLAM_NSPLIT=8
run RBSLAM_NSPLIT=4 RBSLAM_STREAM_CFG=12,2
