#!/bin/bash
set -u
mkdir -p gpurun_out
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 ) > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; echo "bench n2 rc=$?"; cut -c1-600 gpurun_out/r02_bench_n2.json; tail -5 gpurun_out/r02_bench_n2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_n2.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'weak',d['weak_scaling']['value'])
print(json.dumps(d['metrics_config5'])[:900])
PY
