#!/bin/bash
mkdir -p gpurun_out
tag=c13
timeout 1200 python -m pytest tests/test_gpu_sharded.py -q -m gpu --timeout 900 > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/${tag}_tests.log
