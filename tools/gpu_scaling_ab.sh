#!/bin/bash
# 8-GPU box: scaling of the sharded filter (strong: 10^4 particles in total, weak: 10^4 per GPU), fused vs
# fetch-then-pass migration, the C5 smoother slice on all GPUs
mkdir -p gpurun_out
tag=ab
export RBSLAM_CHOL_KERNEL=solve   # the verified K7 kernel: this call is about scaling
nvidia-smi -L | wc -l
run() { # name nproc env...
  name=$1; np=$2; shift 2
  env "$@" timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) bench.py --gpus $np --steps 40 --warmup 3 --e2e-steps 48 $EXTRA > gpurun_out/${tag}_$name.json 2> gpurun_out/${tag}_$name.err
  echo "$name rc=$?"
  python - <<P
import json
try:
    line=[l for l in open("gpurun_out/${tag}_$name.json").read().splitlines() if l.startswith("{")][-1]
    d=json.loads(line)
    print("$name", "strong", round(d["value"]), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), "| weak", round(d["weak"]["value"]), round(d["weak"]["ms_per_step"],3))
    print("   strong phases", {k:round(v,3) for k,v in d["roofline"]["phases_ms_per_step"].items()})
    print("   weak phases", {k:round(v,3) for k,v in d["weak"]["phases_ms_per_step"].items()})
    if "smoother" in d: print("   smoother", d["smoother"])
except Exception as e: print("$name failed", e); print(open("gpurun_out/${tag}_$name.err").read()[-800:])
P
}
EXTRA=""            run n8_fused 8 X=1
EXTRA="--no-smoother" run n8_unfused 8 RBSLAM_FUSED=0
EXTRA="--no-smoother" run n4_fused 4 X=1
EXTRA="--no-smoother" run n4_unfused 4 RBSLAM_FUSED=0
timeout 400 python -m pytest tests/test_gpu_sharded.py -q -m gpu --timeout 300 -k "4-64" > gpurun_out/${tag}_tests.log 2>&1; echo "4-rank tests over NVLink rc=$?"; tail -3 gpurun_out/${tag}_tests.log
