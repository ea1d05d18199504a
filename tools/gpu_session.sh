#!/bin/bash
set -u
mkdir -p gpurun_out
python tools/microbench.py glueprof || exit 1
timeout 900 ncu --set full --clock-control none -k regex:"dw3x3_tma_kernel|stem_kernel|bilinear_ac_kernel|dw3x3_img_kernel|mbconv_kernel" --launch-skip 5 --launch-count 5 -f -o gpurun_out/r02_glue python tools/microbench.py glueprof > gpurun_out/r02_glue_ncu.log 2>&1; echo "ncu rc=$?"
