#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "metrics or eval_driver or post_u8" > gpurun_out/r02_tests_new.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/r02_tests_new.log
tail -5 gpurun_out/r02_tests_new.log | cut -c1-300
for st in 3 4; do for pairs in 2048 16384; do for dt in f32 u8; do
  UAVSAL_OPTIONS="9=1,10=$st" timeout 300 python tools/bench_metrics.py --pairs $pairs --dtype $dt > gpurun_out/r02_metrics_async_st${st}_${dt}_${pairs}.json 2> gpurun_out/r02_metrics_async_st${st}_${dt}_${pairs}.err
  echo "async stream st=$st $dt pairs=$pairs rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/r02_metrics_async_st${st}_${dt}_${pairs}.json'));print(round(d['value']),d['ms'],d['roofline']['frac'])"
done; done; done
UAVSAL_OPTIONS="9=1" timeout 300 python tools/bench_metrics.py --pairs 2048 --reps 1 > gpurun_out/r02_plain.log 2>&1 &&
UAVSAL_OPTIONS="9=1" timeout 600 ncu --set full --clock-control none --import-source on -k regex:metrics4_stream -s 1 -c 1 -o gpurun_out/r02_metrics_stream_async python tools/bench_metrics.py --pairs 2048 --reps 1 > gpurun_out/r02_ncu.log 2>&1
echo "ncu rc=$?"
