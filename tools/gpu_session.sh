#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest tests -m gpu -q -x -k "metrics_kernel_variants or pack_weights or resnet18 or auc_judd_dense or eval_driver_sum" > gpurun_out/r02_memcheck.log 2>&1; echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|passed|failed|Invalid|========= " gpurun_out/r02_memcheck.log | head -30
