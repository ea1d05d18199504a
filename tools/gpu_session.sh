#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "glue or uavsal_call or config2 or constructor or scheduling" > gpurun_out/r02l_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r02l_tests.log | cut -c1-300
timeout 300 python tools/microbench.py bilinear 2>&1 | tail -4 | tee gpurun_out/r02l_bilinear.txt
