#!/bin/bash
# round-2 evidence refresh at HEAD: full GPU test suite, default bench line, ncu launch list of the 120-frame plan
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r02c_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r02c_tests.log | cut -c1-200
( timeout 900 python bench.py ) > gpurun_out/r02c_bench_n1.json 2> gpurun_out/r02c_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json; d=json.load(open('gpurun_out/r02c_bench_n1.json'))
print(round(d['value']), round(d['e2e']['value']), d['clocks'], d['roofline']['frac'], d['wall_s'])
PY
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none --csv --log-file gpurun_out/r02c_clip120_kernels.csv python tools/profile_call.py exact 120 > gpurun_out/r02c_ncu.log 2>&1; echo "ncu rc=$?"
python tools/summarize_ncu.py gpurun_out/r02c_clip120_kernels.csv gpurun_out/r02c_clip120 | tail -3
