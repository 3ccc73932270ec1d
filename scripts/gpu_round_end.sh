#!/bin/bash
# One gpurun call at the end of a round: GPU parity suite, smoke(), the bench line (with CPU baseline), the reference
# arm, the ncu launch list and one `--set full` capture of each hot kernel.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_round_end.sh <tag>'
tag=${1:-r1x}
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -3 $out/pytest_$tag.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke_$tag.log 2>&1; echo "smoke rc=$?"; tail -2 $out/smoke_$tag.log
timeout 600 python bench.py > $out/bench_$tag.json 2> $out/bench_$tag.log; echo "bench rc=$?"
timeout 600 python bench.py --impl reference > $out/bench_${tag}_ref.json 2> $out/bench_${tag}_ref.log; echo "ref rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu > $out/ncu_launch_$tag.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'viterbiBandedKernel|viterbiFullKernel|seedScanKernel|seedSortLiteralKernel|estPlanKernel|estimateStagedKernel|placeKernel' \
    --launch-skip 16 --launch-count 8 -o $out/prof_$tag -f python bench.py --steps 1 --warmup 1 --no-cpu > $out/ncu_full_$tag.log 2>&1; echo "ncu full rc=$?"
ncu -i $out/prof_$tag.ncu-rep --page raw --csv > $out/prof_${tag}_raw.csv 2>/dev/null
python - <<PY
import json
for f in ("bench_$tag.json","bench_${tag}_ref.json"):
    d=json.loads(open("$out/"+f).read().strip().splitlines()[-1])
    print(f, round(d["value"],1), round(d["e2e"]["value"],1), {k: round(v["ms"],2) for k,v in d.get("stages",{}).items()})
PY
