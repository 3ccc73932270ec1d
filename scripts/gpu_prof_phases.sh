#!/bin/bash
# phase clocks of the literal seed sort and the K3 pipeline (library built with HU_PROF=1)
mkdir -p gpurun_out
HU_PROF=1 python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_prof.json 2> gpurun_out/bench_prof.log; echo "rc=$?"
grep -a "hu_prof\|stage" gpurun_out/bench_prof.log; python -c "
import json; d=json.load(open('gpurun_out/bench_prof.json')); print(d['value'], {k:round(v['ms'],2) for k,v in d['stages'].items()})"
