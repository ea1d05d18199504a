#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "fused_expand or block_modules or backbone_levels or uavsal_call" > gpurun_out/r02_expdw_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02_expdw_tests.log | cut -c1-300
python - <<'PY' 2>&1 | tee gpurun_out/r02_expdw_remap.txt
import sys; sys.argv=['x','none']; sys.path.insert(0,'tools')
import microbench as mb
mb.expdw(120, 180, 320, 16, 96, 2); mb.expdw(120, 90, 160, 24, 144, 2)
PY
