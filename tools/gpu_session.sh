#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_tests_all.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/r02_tests_all.log
tail -25 gpurun_out/r02_tests_all.log | cut -c1-300
( time timeout 900 python bench.py --steps 6 --warmup 3 --skip-aux --skip-cpu ) > gpurun_out/r02_bench_quick.json 2> gpurun_out/r02_bench_quick.err; echo "bench rc=$?"; cut -c1-260 gpurun_out/r02_bench_quick.json; tail -3 gpurun_out/r02_bench_quick.err
