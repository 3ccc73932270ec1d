#!/bin/bash
# One gpurun call at the end of a round (about 15 minutes of box time): the whole GPU suite, smoke(), the three bench shapes with
# their CPU baseline and parity block, the reference arm, the ncu launch list of a step and one `--set full` capture of each hot
# kernel, and the command-line programs against each other on BASELINE.json configs[0] / [1].
#   gpurun --timeout 1700 -- 'bash scripts/gpu_final.sh <tag>'
tag=${1:-r2d}
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q > $out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -3 $out/pytest_$tag.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke_$tag.log 2>&1; echo "smoke rc=$?"; tail -1 $out/smoke_$tag.log
timeout 600 python bench.py > $out/bench_${tag}_full.json 2> $out/bench_${tag}_full.log; echo "bench full rc=$?"
timeout 600 python bench.py --impl reference > $out/bench_${tag}_ref.json 2> $out/bench_${tag}_ref.log; echo "bench ref rc=$?"
for sc in pe long; do
  timeout 700 python bench.py --scale $sc --steps 6 --warmup 3 > $out/bench_${tag}_$sc.json 2> $out/bench_${tag}_$sc.log; echo "bench $sc rc=$?"
done
python - <<PY
import json
for sc in ("full", "ref", "pe", "long"):
    try:
        d = json.loads(open("$out/bench_${tag}_%s.json" % sc).read().strip().splitlines()[-1])
        print(sc, round(d["value"], 1), round(d["e2e"]["value"], 1), {k: round(v["ms"], 2) for k, v in d.get("stages", {}).items()},
              "cpu", d.get("cpu_baseline") and round(d["cpu_baseline"]["value"], 1), "parity", d.get("parity"), "roofline", d.get("roofline", {}).get("kernel"), d.get("roofline", {}).get("frac"))
    except Exception as e:
        print(sc, "no json", e)
PY
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu > $out/ncu_launch_$tag.log 2>&1; echo "ncu launches rc=$?"
timeout 700 ncu --set full --clock-control none --import-source on -k regex:'viterbiBandedKernel|seedMmKernel|seedSortLiteralKernel|estimateStagedKernel|placeKernel' \
    --launch-skip 10 --launch-count 5 -o $out/prof_$tag -f python bench.py --steps 1 --warmup 1 --no-cpu > $out/ncu_full_$tag.log 2>&1; echo "ncu full rc=$?"
ncu -i $out/prof_$tag.ncu-rep --page raw --csv > $out/prof_${tag}_raw.csv 2>/dev/null
python profiles/stalls.py $out/prof_${tag}_raw.csv > $out/prof_${tag}_digest.txt 2>/dev/null; head -8 $out/prof_${tag}_digest.txt
timeout 600 bash scripts/gpu_c1_cli.sh 4000 > $out/c1_cli_$tag.log 2>&1; echo "c1 cli rc=$?"; cat $out/c1_cli.txt
