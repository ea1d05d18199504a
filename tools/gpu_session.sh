#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "pointwise or hidden or block_modules or conv3x3 or uavsal_call or recurrences or convlstm" > gpurun_out/r02_tests_gemm.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02_tests_gemm.log
echo "--- bias hoisted"; python tools/microbench.py f32set 2>&1 | tail -9
python tools/microbench.py r2 2>&1 | grep "pw\[" | head -12
( timeout 600 python bench.py --steps 6 --warmup 3 --skip-aux --skip-cpu ) > gpurun_out/r02_bench_biashoist.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_biashoist.json')); print('bench', round(d['value']), round(d['e2e']['value']), d['breakdown_per_plan']['uavsal_pw_gemm'], d['breakdown_per_plan']['uavsal_conv3x3'], d['breakdown_per_plan']['uavsal_twa_sequence'])"
