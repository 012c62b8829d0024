#!/bin/bash
# default bench twice in a row: device-resident value, end-to-end value, Kalman phase
for i in 1 2; do
  timeout -s KILL 300 python bench.py --no-cpu-baseline "$@" 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('run $i', round(d['value']), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'kalman', round(d['roofline']['phases_ms_per_step']['kalman'],2), d['clocks']['sm_mhz'], d['clocks']['reasons'])"
done
