#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "depthwise or readout or q16 or hidden or uavsal_call or config2" > gpurun_out/r02f_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r02f_tests.log | cut -c1-300
python - <<'PY' 2>&1 | tee gpurun_out/r02f_microbench.txt
import sys; sys.argv=['x','none']; sys.path.insert(0,'tools')
import microbench as mb
mb.readout(120,45,80,1536,q16=True); mb.readout(120,45,80,1536,q16=False)
mb.dw(2,24,45,80,1536,2,q16=True); mb.dw(2,24,45,80,1536,2,f32=True)
mb.dw(2,120,90,160,144,1,f32=True); mb.dw(2,120,23,40,576,1,f32=True); mb.dw(2,120,45,80,192,2,f32=True); mb.dw(2,20,45,80,1536,1,f32=True); mb.dw(2,20,45,80,1536,1)
PY
( timeout 600 python bench.py --steps 8 --warmup 3 --skip-aux --skip-cpu --dump-ops gpurun_out/r02f_ops.txt ) > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err; echo "bench rc=$?"
python - <<'PY'
import json; d=json.load(open('gpurun_out/r02f_bench.json'))
print(round(d['value']), round(d['e2e']['value']), d['clocks'], d['roofline']['frac'])
for k,v in d['breakdown_per_plan'].items(): print(k, v)
PY
