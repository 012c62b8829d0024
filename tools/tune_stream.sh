#!/bin/bash
# Sweep the streaming-kernel knobs on the C4 shape (reduced N).  Usage: tools/tune_stream.sh [N]
N=${1:-4000}
run() {
  echo "== $*"
  env "$@" python bench.py --steps 8 --warmup 3 --particles $N --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); r = d['roofline']
        print('value %.0f  kalman_ms %.3f  frac %.4f  step_ms %.3f' % (d['value'], r['kernel_ms_per_step'], r['frac'], d['ms_per_step']))
    elif 'rror' in l: print(l.strip())
"
}
run RBSLAM_STREAM_CFG=4,4
run RBSLAM_STREAM_CFG=4,4 RBSLAM_STREAM_HINTS=1
run RBSLAM_STREAM_CFG=4,4 RBSLAM_STREAM_HINTS=2
run RBSLAM_STREAM_CFG=4,4 RBSLAM_STREAM_HINTS=3
run RBSLAM_STREAM_CFG=4,3 RBSLAM_CTAS_PER_SM=2
run RBSLAM_STREAM_CFG=4,3 RBSLAM_CTAS_PER_SM=1
run RBSLAM_STREAM_CFG=4,5
run RBSLAM_STREAM_CFG=4,6
run RBSLAM_STREAM_CFG=2,6 RBSLAM_CTAS_PER_SM=2
run RBSLAM_STREAM_CFG=2,8
run RBSLAM_STREAM_CFG=8,3
run RBSLAM_STREAM_CFG=8,2 RBSLAM_CTAS_PER_SM=1
run RBSLAM_STREAM_CFG=4,4 RBSLAM_NSPLIT=1
run RBSLAM_STREAM_CFG=4,4 RBSLAM_NSPLIT=4
