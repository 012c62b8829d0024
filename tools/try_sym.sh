#!/bin/bash
# symmetric streaming kernels (kalman_variant 4 = SIMT, 5 = DMMA, 6 = DMMA with producer warp + deep ring):
# parity tests, then the C4 bench.  Variant 6 has not run on a GPU yet:
#   RBSLAM_TEST_UNVERIFIED=1 SYM_K=pipelined tools/try_sym.sh 6 5 0
mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests/test_gpu_symmetric.py -q -m gpu -k "${SYM_K:-dmma}" 2>&1 | tail -12 | cut -c1-250
for v in "$@"; do
  timeout -s KILL 200 python bench.py --variant $v --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('variant $v', round(d['value']), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'kalman', round(d['roofline']['phases_ms_per_step']['kalman'],2), d['clocks']['sm_mhz'], d['clocks']['reasons'])"
done
