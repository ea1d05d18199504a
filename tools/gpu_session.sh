#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r02q_tests.log 2>&1; echo "tests rc=$?"; tail -8 gpurun_out/r02q_tests.log | cut -c1-300
