#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r02m_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r02m_tests.log | cut -c1-300
