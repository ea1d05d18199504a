#!/bin/bash
# One gpurun call: new/changed-kernel tests first (own process: a sticky CUDA fault must not take the rest down), then the
# remaining GPU suite, then the metrics micro-benchmark (A/B of the kernel variants) and an ncu capture of the resident kernel.
set -u
mkdir -p gpurun_out
K='metrics or auc or convlstm_config3 or eval_driver or post_u8'
timeout 600 python -m pytest tests -m gpu -q -k "$K" > gpurun_out/r02_tests_new.log 2>&1; echo "new tests rc=$?" | tee -a gpurun_out/r02_tests_new.log
tail -5 gpurun_out/r02_tests_new.log
timeout 900 python -m pytest tests -m gpu -q -k "not ($K)" > gpurun_out/r02_tests_rest.log 2>&1; echo "rest rc=$?" | tee -a gpurun_out/r02_tests_rest.log
tail -5 gpurun_out/r02_tests_rest.log
for opt in 2 1; do for dt in f32 u8; do
  UAVSAL_OPTIONS="9=$opt" timeout 300 python tools/bench_metrics.py --pairs 2048 --dtype $dt > gpurun_out/r02_metrics_opt${opt}_${dt}.json 2> gpurun_out/r02_metrics_opt${opt}_${dt}.err
  echo "metrics opt=$opt $dt rc=$?"; cat gpurun_out/r02_metrics_opt${opt}_${dt}.json
done; done
timeout 300 python tools/bench_metrics.py --pairs 512 --reps 1 > gpurun_out/r02_metrics_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:metrics4_tmem -s 1 -c 1 -o gpurun_out/r02_metrics_tmem python tools/bench_metrics.py --pairs 512 --reps 1 > gpurun_out/r02_metrics_ncu.log 2>&1
echo "ncu rc=$?"
