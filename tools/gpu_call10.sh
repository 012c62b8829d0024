#!/bin/bash
mkdir -p gpurun_out
tag=c10
timeout 1800 python -m pytest tests -q -m gpu --timeout 900 > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$?"
tail -8 gpurun_out/${tag}_tests.log
