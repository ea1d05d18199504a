#!/bin/bash
set -u
mkdir -p gpurun_out
for n in 8 4; do
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 3 ) > gpurun_out/r02_bench_n$n.json 2> gpurun_out/r02_bench_n$n.err; echo "bench n$n rc=$?"
python - <<PY
import json
lines=[l for l in open('gpurun_out/r02_bench_n$n.json') if l.startswith('{')]
d=json.loads(lines[-1])
print('N=$n value',round(d['value']),'e2e',round(d['e2e']['value']),'weak',round(d['weak_scaling']['value']),'timed',d['timed_region_s'],'clock samples',d['clocks']['samples'])
m=d['metrics_config5']; print('  metrics5',round(m['pairs_per_s']),m['ms'],m['allreduce'],m['allreduce_ms'],m['frac'],m['means_vs_single_rank_rel_err'])
PY
done
