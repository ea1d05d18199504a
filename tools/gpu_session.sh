#!/bin/bash
set -u
mkdir -p gpurun_out
python tools/debug_pack.py > gpurun_out/r02_debug_pack.log 2>&1; cat gpurun_out/r02_debug_pack.log
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02_tests_all.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/r02_tests_all.log
tail -6 gpurun_out/r02_tests_all.log
( time timeout 1200 python bench.py ) > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"; cat gpurun_out/r02_bench_n1.json; tail -8 gpurun_out/r02_bench_n1.err
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/r02_bench_ref.json; tail -4 gpurun_out/r02_bench_ref.err
