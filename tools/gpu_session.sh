#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "q16 or fused_depthwise_project" > gpurun_out/r02o_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02o_tests.log | cut -c1-300
timeout 300 python tools/microbench.py dwproj_q16 2>&1 | tail -4 | tee gpurun_out/r02o_dwproj.txt
