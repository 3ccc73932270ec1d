#!/bin/bash
# A/B of an environment switch on the short bench: bash scripts/gpu_ab.sh VAR "v1 v2 ..." [bench args]
var=$1; vals=$2; shift; shift
mkdir -p gpurun_out
for v in $vals; do
  env $var=$v python bench.py --steps 4 --warmup 3 --no-cpu "$@" > gpurun_out/ab_${var}_$v.json 2> gpurun_out/ab_${var}_$v.log; echo "$var=$v rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/ab_${var}_$v.json')); print(round(d['value']), round(d['e2e']['value']), {k:round(x['ms'],2) for k,x in d['stages'].items()}, d.get('parity'))"
done
