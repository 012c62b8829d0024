#!/bin/bash
# 8-GPU weak-scaling bench, migration overlap on (RBSLAM_OVERLAP=1) and off (default)
N=${1:-8}
for mode in overlap nooverlap; do
  if [ $mode = overlap ]; then export RBSLAM_OVERLAP=1; else unset RBSLAM_OVERLAP; fi
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 10 --warmup 3 2>/dev/null | tail -1 > gpurun_out/bench_n${N}_$mode.json
  python -c "
import json,sys; d=json.load(open('gpurun_out/bench_n${N}_$mode.json')); print('$mode', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['phases_ms_per_step'])"
done
