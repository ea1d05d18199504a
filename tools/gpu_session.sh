#!/bin/bash
# q16 hidden rows: new parity tests, A/B micro-benchmarks, bench with / without
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "q16 or readout or fused_depthwise_project or hidden" > gpurun_out/r02d_tests_q16.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/r02d_tests_q16.log | cut -c1-300
timeout 600 python tools/microbench.py q16 2>&1 | tee gpurun_out/r02d_microbench_q16.txt
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r02d_tests.log 2>&1; echo "all tests rc=$?"; tail -8 gpurun_out/r02d_tests.log | cut -c1-300
( timeout 600 python bench.py --steps 8 --warmup 3 --skip-aux --skip-cpu ) > gpurun_out/r02d_bench_q16.json 2> gpurun_out/r02d_bench_q16.err; echo "bench rc=$?"
python - <<'PY'
import json; d=json.load(open('gpurun_out/r02d_bench_q16.json'))
print(round(d['value']), round(d['e2e']['value']), d['clocks'], d['roofline']['frac'])
for k,v in d['breakdown_per_plan'].items(): print(k, v)
PY
