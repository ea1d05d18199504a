#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_tests_all.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/r02_tests_all.log
tail -4 gpurun_out/r02_tests_all.log | cut -c1-300
for d in 1 3; do
( timeout 600 python bench.py --steps 8 --warmup 3 --skip-aux --skip-cpu --depth $d ) > gpurun_out/r02_bench_depth$d.json 2> gpurun_out/r02_bench_depth$d.err; echo "depth $d rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_depth$d.json')); print(round(d['value']), round(d['e2e']['value']))"
done
( timeout 600 python bench.py --steps 8 --warmup 3 --skip-aux --skip-cpu --clips-per-plan 1 ) > gpurun_out/r02_bench_cpp1.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_cpp1.json')); print('cpp1', round(d['value']), round(d['e2e']['value']))"
( timeout 600 python bench.py --steps 8 --warmup 3 --skip-aux --skip-cpu --precision fast ) > gpurun_out/r02_bench_fast.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_fast.json')); print('fast', round(d['value']), round(d['e2e']['value']))"
