#!/bin/bash
# symmetric streaming kernel (kalman_variant 4): parity tests, then the C4 bench next to the default kernel
mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests/test_gpu_symmetric.py -q -m gpu 2>&1 | tail -8 | cut -c1-250
tools/tune_sym.sh "$@"
