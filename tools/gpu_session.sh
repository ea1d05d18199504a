#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02_tests_all.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/r02_tests_all.log
tail -15 gpurun_out/r02_tests_all.log
for dt in f32 u8; do
  UAVSAL_OPTIONS="9=2" timeout 300 python tools/bench_metrics.py --pairs 2048 --dtype $dt > gpurun_out/r02_metrics_v4_${dt}.json 2> gpurun_out/r02_metrics_v4_${dt}.err
  echo "metrics v4 $dt rc=$?"; cat gpurun_out/r02_metrics_v4_${dt}.json
done
( time timeout 600 python __graft_entry__.py --smoke ) > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/r02_smoke.log
( time timeout 900 python bench.py --steps 10 --warmup 3 ) > gpurun_out/r02_bench_engine.json 2> gpurun_out/r02_bench_engine.err; echo "bench rc=$?"; cut -c1-700 gpurun_out/r02_bench_engine.json; tail -5 gpurun_out/r02_bench_engine.err
