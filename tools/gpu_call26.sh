#!/bin/bash
# 4-GPU box: where does the replicated smoother spend its time?  + sharded filter sanity on the final build
mkdir -p gpurun_out
tag=c26
nvidia-smi -L | wc -l
timeout 600 python tools/replica_diag.py 1,2,4 24 2>&1 | grep "^{" 
timeout 300 python tools/replica_diag.py 4 48 2>&1 | grep "^{"
