#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "stay_inside" > gpurun_out/r02w_tests.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/r02w_tests.log | cut -c1-250
