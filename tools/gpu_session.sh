#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "mbconv or block_modules or backbone_levels or uavsal_call or constructor" > gpurun_out/r02y_tests.log 2>&1; echo "tests rc=$?"; tail -8 gpurun_out/r02y_tests.log | cut -c1-300
timeout 300 python tools/microbench.py mbconv 2>&1 | tail -8 | tee gpurun_out/r02y_mbconv.txt
