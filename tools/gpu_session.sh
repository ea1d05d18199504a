#!/bin/bash
# round-2 evidence at HEAD (what profiles/r02_* were produced with): full GPU suite, default bench line (+ reference arm), ncu launch
# list of the 120-frame plan, smoke() launch list, parity-margin report.  Run through gpurun; tools/gpu_session_n.sh N for N > 1.
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r02z_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02z_tests.log | cut -c1-200
( timeout 900 python bench.py --dump-ops gpurun_out/r02z_ops.txt ) > gpurun_out/r02z_bench_n1.json 2> gpurun_out/r02z_bench_n1.err; echo "bench rc=$?"
( timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/r02z_bench_ref.json 2> gpurun_out/r02z_bench_ref.err; echo "ref rc=$?"
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none --csv --log-file gpurun_out/r02z_clip120_kernels.csv python tools/profile_call.py exact 120 > gpurun_out/r02z_ncu.log 2>&1; echo "ncu rc=$?"
python tools/summarize_ncu.py gpurun_out/r02z_clip120_kernels.csv gpurun_out/r02z_clip120 | tail -2
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02z_smoke_kernels.csv python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02z_smoke.log 2>&1; echo "smoke-ncu rc=$?"; tail -1 gpurun_out/r02z_smoke.log
python tools/parity_report.py > gpurun_out/r02z_parity_report.txt 2>&1; grep "saliency" gpurun_out/r02z_parity_report.txt
