#!/bin/bash
# round-2 evidence at HEAD: full GPU test suite, default bench line, ncu launch list of the 120-frame plan, ncu --set full of the top kernel
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r02i_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r02i_tests.log | cut -c1-200
( timeout 900 python bench.py --dump-ops gpurun_out/r02i_ops.txt ) > gpurun_out/r02i_bench_n1.json 2> gpurun_out/r02i_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json; d=json.load(open('gpurun_out/r02i_bench_n1.json'))
print(round(d['value']), round(d['e2e']['value']), d['clocks'], d['roofline'], d['wall_s'])
PY
( timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/r02i_bench_ref.json 2> gpurun_out/r02i_bench_ref.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/r02i_bench_ref.json
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none --csv --log-file gpurun_out/r02i_clip120_kernels.csv python tools/profile_call.py exact 120 > gpurun_out/r02i_ncu.log 2>&1; echo "ncu rc=$?"
python tools/summarize_ncu.py gpurun_out/r02i_clip120_kernels.csv gpurun_out/r02i_clip120 | tail -3
python tools/microbench.py pairq16prof
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc2_kernel --launch-skip 2 -c 1 -f -o gpurun_out/r02i_gemm_pair_q16 python tools/microbench.py pairq16prof > gpurun_out/r02i_ncu2.log 2>&1; echo "ncu2 rc=$?"
