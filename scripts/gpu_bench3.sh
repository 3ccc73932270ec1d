#!/bin/bash
# parity tests named on the command line (default: the fast parity files), then the three bench shapes without the CPU leg
tag=${1:-q}; shift
sel=${*:-tests/test_gpu_parity.py tests/test_gpu_more.py tests/test_gpu_sort.py tests/test_gpu_ref.py}
mkdir -p gpurun_out
timeout 1200 python -m pytest $sel -m gpu -q -x > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_$tag.log
for sc in full pe long; do
  timeout 600 python bench.py --scale $sc --steps 6 --warmup 3 --no-cpu > gpurun_out/bench_${tag}_$sc.json 2> gpurun_out/bench_${tag}_$sc.log; echo "bench $sc rc=$?"
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_${tag}_$sc.json'))
    print('$sc', round(d['value']), round(d['e2e']['value']), {k:round(v['ms'],2) for k,v in d['stages'].items()})
except Exception as e: print('no json', e)
PY
done
