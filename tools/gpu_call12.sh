#!/bin/bash
# 2-GPU box: the sharded paths over NVLink (peer TMA reads of the fused migration), then the scaling bench
mkdir -p gpurun_out
tag=c12
nvidia-smi -L | head -3
timeout 900 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_smoother.py -q -m gpu --timeout 600 -k "sharded or group or replicas" > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$?"
tail -6 gpurun_out/${tag}_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 40 --warmup 3 --e2e-steps 48 > gpurun_out/${tag}_bench_n2.json 2> gpurun_out/${tag}_bench_n2.err; echo "bench n2 rc=$?"
tail -c 2500 gpurun_out/${tag}_bench_n2.json; tail -5 gpurun_out/${tag}_bench_n2.err
RBSLAM_FUSED=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 40 --warmup 3 --e2e-steps 48 --no-smoother > gpurun_out/${tag}_bench_n2_unfused.json 2> gpurun_out/${tag}_bench_n2_unfused.err; echo "bench n2 unfused rc=$?"
python - <<P
import json
for f in ["gpurun_out/c12_bench_n2.json","gpurun_out/c12_bench_n2_unfused.json"]:
    try:
        d=json.load(open(f)); print(f, round(d["value"]), d["scaling"], round(d["ms_per_step"],3), "weak", round(d["weak"]["value"]), round(d["weak"]["ms_per_step"],3), d["roofline"]["phases_ms_per_step"], d["weak"]["phases_ms_per_step"])
    except Exception as e: print(f, "failed", e)
P
