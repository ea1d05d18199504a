#!/bin/bash
set -u
mkdir -p gpurun_out
for p in 1 0 1 0; do
( UAVSAL_BACK_PRIORITY=$p timeout 600 python bench.py --steps 8 --warmup 3 --skip-aux --skip-cpu ) > gpurun_out/r02j_bench_prio$p.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/r02j_bench_prio$p.json')); print('prio $p', round(d['value']), round(d['e2e']['value']), d['clocks']['sm_mhz'], d['weak_scaling']['value'])"
done
( timeout 600 python bench.py --steps 8 --warmup 3 --skip-aux --skip-cpu --clips-per-plan 4 ) > gpurun_out/r02j_bench_cpp4.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/r02j_bench_cpp4.json')); print('cpp4', round(d['value']), round(d['e2e']['value']), d['clocks']['sm_mhz'])"
( timeout 600 python bench.py --steps 8 --warmup 3 --skip-aux --skip-cpu --depth 3 ) > gpurun_out/r02j_bench_d3.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/r02j_bench_d3.json')); print('depth3', round(d['value']), round(d['e2e']['value']), d['clocks']['sm_mhz'])"
