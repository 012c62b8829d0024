#!/bin/bash
N=${1:-4000}
run() {
  echo "== $*"
  env "$@" python bench.py --steps 24 --warmup 4 --particles $N --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); r = d['roofline']
        print('value %.0f  kalman_ms %.3f  frac %.4f  step_ms %.3f' % (d['value'], r['kernel_ms_per_step'], r['frac'], d['ms_per_step']))
    elif 'rror' in l: print(l.strip())
"
}
run RBSLAM_STREAM_CFG=8,2
run RBSLAM_STREAM_CFG=8,2 RBSLAM_NSPLIT=1
run RBSLAM_STREAM_CFG=8,2 RBSLAM_NSPLIT=4
run RBSLAM_STREAM_CFG=4,4
run RBSLAM_STREAM_CFG=6,4
run RBSLAM_STREAM_CFG=8,3
run RBSLAM_STREAM_CFG=4,6
run RBSLAM_STREAM_CFG=12,2
run RBSLAM_STREAM_CFG=8,2 RBSLAM_NSPLIT=8
