#!/bin/bash
# every bench workload once (no ncu); JSON lines under gpurun_out/bench_<workload>.log
for w in ${@:-1080p 4k 480p 4k-q1}; do
  python bench.py --workload $w > gpurun_out/bench_$w.log 2> gpurun_out/bench_$w.err || tail -3 gpurun_out/bench_$w.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_$w.log').read().strip().splitlines()[-1])
    print('$w', round(d['value']), 'fps  e2e', round(d['e2e']['value']), ' ms/step', round(d['ms_per_step'],2), {k:round(v['ms'],2) for k,v in d['stages'].items()}, 'frac', round(d['roofline']['frac'],3), 'cpu', round(d['cpu_baseline']['value'],1))
except Exception as e: print('$w', 'failed', e)
PY
done
