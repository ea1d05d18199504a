#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_tests_all.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/r02_tests_all.log
tail -6 gpurun_out/r02_tests_all.log | cut -c1-300
( time timeout 600 python __graft_entry__.py --smoke ) > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/r02_smoke.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_smoke_launches.csv python __graft_entry__.py --smoke > gpurun_out/r02_smoke_ncu.log 2>&1; echo "smoke ncu rc=$?"
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r02_smoke_launches.csv', errors='replace')) if len(r)>10]
h=rows[0]; ik=h.index('Kernel Name')
c=collections.Counter(r[ik].split('(')[0][:60] for r in rows[1:])
print(len(rows)-1, 'launches in the first 400:'); [print('  %4d %s'%(v,k)) for k,v in c.most_common(25)]
PY
