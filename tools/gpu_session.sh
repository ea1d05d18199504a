#!/bin/bash
set -u
mkdir -p gpurun_out
python tools/parity_report.py 2>&1 | tee gpurun_out/r02_parity_report.txt
