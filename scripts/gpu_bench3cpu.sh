#!/bin/bash
# the three bench shapes WITH the CPU leg and the parity block (the lines that go under profiles/)
tag=${1:-r2c}
mkdir -p gpurun_out
for sc in full pe long; do
  timeout 900 python bench.py --scale $sc > gpurun_out/bench_${tag}_$sc.json 2> gpurun_out/bench_${tag}_$sc.log; echo "bench $sc rc=$?"
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_${tag}_$sc.json'))
    print('$sc', round(d['value']), round(d['e2e']['value']), {k:round(v['ms'],2) for k,v in d['stages'].items()}, d['parity'] and {k: d['parity'][k] for k in ('checked','mismatched','near_tie_reordered')}, d['cpu_baseline'] and round(d['cpu_baseline']['value'],1))
except Exception as e: print('no json', e)
PY
done
