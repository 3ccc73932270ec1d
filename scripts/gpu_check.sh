#!/bin/bash
# quick GPU check: parity tests then one bench line (no CPU baseline).  Usage: gpurun -- 'bash scripts/gpu_check.sh <tag> [pytest -k expr]'
tag=${1:-chk}
out=gpurun_out
mkdir -p $out
if [ -n "$2" ]; then sel=(-k "$2"); else sel=(); fi
timeout 900 python -m pytest tests -m gpu -x -q "${sel[@]}" > $out/pytest_$tag.log 2>&1; echo "pytest rc=$?" | tee -a $out/pytest_$tag.log
tail -15 $out/pytest_$tag.log
timeout 600 python bench.py --no-cpu > $out/bench_$tag.json 2> $out/bench_$tag.log; echo "bench rc=$?"
tail -3 $out/bench_$tag.log
python - <<PY
import json
try:
    d=json.loads(open("$out/bench_$tag.json").read().strip().splitlines()[-1])
    print(round(d["value"]), round(d["e2e"]["value"]), {k:round(v["ms"],2) for k,v in d["stages"].items()})
except Exception as e:
    print("no bench line:", e)
PY
