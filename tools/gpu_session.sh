#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "pointwise or hidden or block_modules or conv3x3 or uavsal_call or depthwise or fused or config2 or backbone" > gpurun_out/r02_tests_gemm.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r02_tests_gemm.log | cut -c1-200
echo "--- fragment-layout stores (new default)"; python tools/microbench.py f32set 2>&1 | tail -9
echo "--- staged (3=0x200000)"; UAVSAL_OPTIONS="3=0x200000" python tools/microbench.py f32set 2>&1 | tail -9
for o in 0 0x200000; do
( UAVSAL_OPTIONS="3=$o" timeout 600 python bench.py --steps 8 --warmup 3 --skip-aux --skip-cpu ) > gpurun_out/r02_bench_f32frag_$o.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_f32frag_$o.json')); print('opt $o', round(d['value']), round(d['e2e']['value']), d['breakdown_per_plan']['uavsal_pw_gemm'])"
done
