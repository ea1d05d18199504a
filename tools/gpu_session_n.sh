#!/bin/bash
# multi-GPU bench line (torchrun + NCCL), N = $1
set -u
N=$1
mkdir -p gpurun_out
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 ) > gpurun_out/r02z_bench_n$N.json 2> gpurun_out/r02z_bench_n$N.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r02z_bench_n$N.json') if l.startswith('{')][-1])
print(d['n_gpus'], round(d['value']), round(d['e2e']['value']), d['clocks'], d['metrics_config5']['pairs_per_s'], d['metrics_config5'].get('allreduce'), d['timed_region_s'])
PY
