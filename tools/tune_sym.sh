#!/bin/bash
# symmetric kernel (kalman_variant 4) experiments: RBSLAM_SYM_FLAGS (2 = whole-column copy)
# and RBSLAM_SYM_CFG (columns per stage, stages)
mkdir -p gpurun_out
run() {
  RBSLAM_SYM_FLAGS=$1 RBSLAM_SYM_CFG=$2 timeout -s KILL 200 python bench.py --variant 4 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('flags=$1 cfg=$2', round(d['value']), round(d['ms_per_step'],2), 'kalman', round(d['roofline']['phases_ms_per_step']['kalman'],2), d['clocks']['sm_mhz'])"
}
for spec in "$@"; do run ${spec%%:*} ${spec##*:}; done
