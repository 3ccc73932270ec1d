#!/bin/bash
tag=${1:-b}
mkdir -p gpurun_out
python bench.py --steps 4 --warmup 3 --no-cpu > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.log; echo "bench rc=$?"
grep -a "hu_prof" gpurun_out/bench_$tag.log
python -c "
import json; d=json.load(open('gpurun_out/bench_$tag.json')); print(round(d['value']), round(d['e2e']['value']), {k:round(v['ms'],2) for k,v in d['stages'].items()})"
