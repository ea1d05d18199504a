#!/bin/bash
set -u
mkdir -p gpurun_out
python tools/profile_call.py exact 120 > gpurun_out/r02_profile_plain.log 2>&1 &&
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02_clip120_kernels.csv python tools/profile_call.py exact 120 > gpurun_out/r02_profile_ncu.log 2>&1
echo "ncu plan rc=$?"; tail -2 gpurun_out/r02_profile_ncu.log
timeout 300 python tools/bench_metrics.py --pairs 2048 --reps 1 > gpurun_out/r02_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:metrics4 -s 1 -c 1 --csv --log-file gpurun_out/r02_metrics_kernels.csv python tools/bench_metrics.py --pairs 2048 --reps 1 > gpurun_out/r02_ncu_m.log 2>&1
echo "ncu metrics rc=$?"
( time timeout 1200 python bench.py ) > gpurun_out/r02_bench_n1b.json 2> gpurun_out/r02_bench_n1b.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r02_bench_n1b.json
