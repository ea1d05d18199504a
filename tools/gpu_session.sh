#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "stem or config1 or backbone_levels or block_modules or fused_depthwise_project" > gpurun_out/r02r_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02r_tests.log | cut -c1-300
timeout 300 python tools/microbench.py stem 2>&1 | tail -2 | tee gpurun_out/r02r_stem.txt
