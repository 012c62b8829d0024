#!/bin/bash
# final build on an 8-GPU box: the driver's scaling runs (N = 8, 4, 2), reference arm under torchrun, multi-rank tests over NVLink
mkdir -p gpurun_out
tag=sf
nvidia-smi -L | wc -l
run() { # name nproc
  name=$1; np=$2
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) bench.py --gpus $np $EXTRA > gpurun_out/${tag}_$name.json 2> gpurun_out/${tag}_$name.err
  echo "$name rc=$?"
  python - <<P
import json
try:
    line=[l for l in open("gpurun_out/${tag}_$name.json").read().splitlines() if l.startswith("{")][-1]
    d=json.loads(line)
    print("$name", "strong", round(d["value"]), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), "| weak", round(d["weak"]["value"]), round(d["weak"]["ms_per_step"],3), d["clocks"])
    if "smoother" in d: print("   smoother", {k:(round(v,3) if isinstance(v,float) else v) for k,v in d["smoother"].items() if k!="c5"})
except Exception as e: print("$name failed", e); print(open("gpurun_out/${tag}_$name.err").read()[-800:])
P
}
EXTRA="" run n8 8
EXTRA="" run n4 4
EXTRA="--no-smoother" run n2 2
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --impl reference --gpus 2 --steps 6 --warmup 3 2>/dev/null | tail -1 | cut -c1-300
timeout 600 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_smoother.py -q -m gpu --timeout 300 -k "sharded or group or replicas" > gpurun_out/${tag}_tests.log 2>&1; echo "multi-GPU tests rc=$?"; tail -3 gpurun_out/${tag}_tests.log
