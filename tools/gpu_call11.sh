#!/bin/bash
mkdir -p gpurun_out
tag=c11
timeout 1800 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_filter.py tests/test_gpu_packed.py -q -m gpu --timeout 900 > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$?"
tail -8 gpurun_out/${tag}_tests.log
