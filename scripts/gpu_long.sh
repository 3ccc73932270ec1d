#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 900 python bench.py --no-cpu --scale long --batch 1024 --steps 3 --warmup 3 > $out/bench_long.json 2> $out/bench_long.log; echo "long rc=$?"
tail -3 $out/bench_long.log
python - <<PY
import json
try:
    d=json.loads(open("$out/bench_long.json").read().strip().splitlines()[-1])
    print(round(d["value"]), round(d["e2e"]["value"]), {k:round(v["ms"],2) for k,v in d["stages"].items()}, d.get("viterbi_gcups"))
except Exception as e:
    print("no bench line:", e)
PY
bash scripts/gpu_quick.sh k4b
