#!/bin/bash
# end-of-round verification: full GPU suite, smoke(), default bench (with CPU baseline), reference arm
out=gpurun_out; mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -x -q > $out/pytest_final.log 2>&1; echo "pytest rc=$?"; tail -3 $out/pytest_final.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke_final.log 2>&1; echo "smoke rc=$?"; tail -2 $out/smoke_final.log
( time timeout 900 python bench.py > $out/bench_final.json 2> $out/bench_final.log ) 2>&1 | grep real; echo "bench rc=$?"
( time timeout 900 python bench.py --impl reference > $out/bench_final_ref.json 2> $out/bench_final_ref.log ) 2>&1 | grep real; echo "ref rc=$?"
python - <<PY
import json
for f in ("bench_final.json","bench_final_ref.json"):
    d=json.loads(open("$out/"+f).read().strip().splitlines()[-1])
    print(f, round(d["value"],1), round(d["e2e"]["value"],1), d.get("cpu_baseline"))
PY
