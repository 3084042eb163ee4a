#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
(timeout 600 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; grep -E "passed|failed|Error|assert" gpurun_out/pytest_gpu.log | tail -12)
for dbg in 0 4; do
(WLB200_XA_DBG=$dbg timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -k regex:"cross_attn" --csv --log-file gpurun_out/xa_dbg$dbg.csv python tools/profile_step.py --streams 32 --tokens 3 --no-graph > gpurun_out/xa_dbg$dbg.log 2>&1; echo "dbg=$dbg: $(python - <<PY
import csv
rows=list(csv.reader(open('gpurun_out/xa_dbg$dbg.csv')))
hi=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
mv=rows[hi].index('Metric Value'); kn=rows[hi].index('Kernel Name')
v=[float(r[mv].replace(',','')) for r in rows[hi+2:] if len(r)>mv and 'combine' not in r[kn]]
c=[float(r[mv].replace(',','')) for r in rows[hi+2:] if len(r)>mv and 'combine' in r[kn]]
v=v[len(v)//2:]
print('n=%d avg %.1f us min %.1f ; combine avg %.1f'%(len(v),sum(v)/len(v)/1000,min(v)/1000, (sum(c)/max(1,len(c)))/1000))
PY
)")
done
(timeout 300 python tools/profile_step.py --streams 32 --tokens 24 > gpurun_out/step32.log 2>&1; echo "32 streams: $(tail -1 gpurun_out/step32.log | sed 's/.*mel ms/mel ms/')")
(timeout 300 python tools/profile_step.py --streams 8 --tokens 24 > gpurun_out/step8.log 2>&1; echo "8 streams: $(tail -1 gpurun_out/step8.log | sed 's/.*mel ms/mel ms/')")
(timeout 1500 python bench.py --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/bench_large.json 2> gpurun_out/bench_large.err; echo "bench large exit $?"; cat gpurun_out/bench_large.json | cut -c1-330; tail -5 gpurun_out/bench_large.err)
