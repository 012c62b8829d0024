#!/bin/bash
# 4-GPU box: sharded filter with the migrant families walked first (default) vs last (RBSLAM_MIGRANTS_FIRST=0)
mkdir -p gpurun_out
for mf in 1 0; do
  RBSLAM_MIGRANTS_FIRST=$mf timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $((29700 + mf)) bench.py --gpus 4 --steps 60 --warmup 3 --e2e-steps 16 --no-smoother > gpurun_out/mf_$mf.json 2> gpurun_out/mf_$mf.err
  python - <<P
import json
d=json.loads([l for l in open("gpurun_out/mf_$mf.json").read().splitlines() if l.startswith("{")][-1])
print("migrants_first=$mf strong", round(d["value"]), round(d["ms_per_step"],3), "kalman", round(d["roofline"]["phases_ms_per_step"]["kalman"],3), "| weak", round(d["weak"]["value"]), round(d["weak"]["ms_per_step"],3), "kalman", round(d["weak"]["phases_ms_per_step"]["kalman"],3), "normalize", round(d["weak"]["phases_ms_per_step"]["normalize"],3))
P
done
timeout 300 python -m pytest tests/test_gpu_sharded.py -q -m gpu --timeout 200 -k "4-64 or group" 2>&1 | tail -2
