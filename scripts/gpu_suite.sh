#!/bin/bash
# whole GPU suite (no -x: every failure is listed), then the three bench shapes with their CPU baseline + parity block
tag=${1:-s}; shift
shapes=${*:-full pe long}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_$tag.log
for sc in $shapes; do
  timeout 900 python bench.py --scale $sc --steps 6 --warmup 3 > gpurun_out/bench_${tag}_$sc.json 2> gpurun_out/bench_${tag}_$sc.log; echo "bench $sc rc=$?"
  tail -3 gpurun_out/bench_${tag}_$sc.log
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_${tag}_$sc.json'))
    print('$sc', round(d['value']), round(d['e2e']['value']), {k:round(v['ms'],2) for k,v in d['stages'].items()}, 'cpu', d['cpu_baseline'] and round(d['cpu_baseline']['value'],1), 'parity', d.get('parity'))
except Exception as e: print('no json', e)
PY
done
