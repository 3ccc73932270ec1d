#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 900 python bench.py --no-cpu --scale pe --batch 4096 --steps 4 --warmup 3 > $out/bench_pe.json 2> $out/bench_pe.log; echo "pe rc=$?"
tail -3 $out/bench_pe.log
python - <<PY
import json
try:
    d=json.loads(open("$out/bench_pe.json").read().strip().splitlines()[-1])
    print(round(d["value"]), round(d["e2e"]["value"]), {k:round(v["ms"],2) for k,v in d["stages"].items()}, d.get("viterbi_gcups"))
except Exception as e:
    print("no bench line:", e)
PY
bash scripts/gpu_quick.sh capw
