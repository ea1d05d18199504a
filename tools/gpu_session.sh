#!/bin/bash
set -u
mkdir -p gpurun_out
for lib in base mr base mr; do
  if [ $lib = mr ]; then export UAVSAL_LIB=$PWD/gpurun_aux/libuavsal_b200_mr.so; else unset UAVSAL_LIB; fi
  echo "=== $lib"
  python - <<'PY' 2>&1 | grep -v "^$"
import sys; sys.argv=['x','none']; sys.path.insert(0,'tools')
import microbench as mb
M=432000
mb.gemm("tc", M, 256, 1536, q16=True); mb.gemm("tc", M, 320, 1920, q16=True); mb.gemm("tc", M, 256, 256, res=True); mb.gemm("tc", M, 32, 256, res=True)
mb.gemm("tc", 1728000, 24, 144, f32=True); mb.gemm("tc", 110400, 96, 576, f32=True); mb.gemm("tc", 28800, 1920, 256)
mb.dwproj(120, 45, 80, 1536, 256, res=True, q16=True); mb.dwproj(120, 45, 80, 1920, 256, q16=True); mb.dwproj(120, 45, 80, 1152, 64)
mb.mbblock(120, 45, 80, 64, 32, True); mb.mbblock(120, 45, 80, 32, 32, True); mb.mbblock(120, 23, 40, 64, 64, True)
mb.twa("tc", 60, 45, 80, 256, batch=2); mb.conv("tc", 120, 45, 80, 448, 256)
PY
done 2>&1 | tee gpurun_out/r02u_maxnreg.txt
for lib in base mr base mr; do
  if [ $lib = mr ]; then export UAVSAL_LIB=$PWD/gpurun_aux/libuavsal_b200_mr.so; else unset UAVSAL_LIB; fi
  ( timeout 600 python bench.py --steps 8 --warmup 3 --skip-aux --skip-cpu ) > gpurun_out/r02u_bench_$lib.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/r02u_bench_$lib.json')); print('$lib', round(d['value']), round(d['e2e']['value']), d['clocks']['sm_mhz'], d['whole_network']['sum_of_kernel_ms_per_plan'])"
done
