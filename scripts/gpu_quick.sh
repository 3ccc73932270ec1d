#!/bin/bash
# quick check of a kernel change: the seed/ parity GPU tests named on the command line (default: all), then a short bench
tag=${1:-q}; shift
sel=${*:-tests}
mkdir -p gpurun_out
python -m pytest $sel -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_$tag.log
python bench.py --steps 4 --warmup 3 --no-cpu > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.log; echo "bench rc=$?"
grep -a "hu_prof" gpurun_out/bench_$tag.log
python -c "
import json; d=json.load(open('gpurun_out/bench_$tag.json')); print(round(d['value']), round(d['e2e']['value']), {k:round(v['ms'],2) for k,v in d['stages'].items()})"
