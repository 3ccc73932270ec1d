#!/bin/bash
# One gpurun call: GPU parity tests, the bench line (with CPU baseline), the ncu launch list and one
# `--set full` capture of each hot kernel.  Usage: gpurun --timeout 1700 -- 'bash scripts/gpu_profile.sh <tag>'
tag=${1:-r1x}
out=gpurun_out
mkdir -p $out
set -x
python -m pytest tests -m gpu -x -q > $out/pytest_$tag.log 2>&1; echo "pytest rc=$?" >> $out/pytest_$tag.log
tail -3 $out/pytest_$tag.log
python bench.py > $out/bench_$tag.json 2> $out/bench_$tag.log; echo "bench rc=$?"
cat $out/bench_$tag.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu > $out/ncu_launch_$tag.log 2>&1; echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'viterbiBandedKernel|viterbiFullKernel|seedScanKernel|seedSortLiteralKernel|estPlanKernel|estimateStagedKernel|placeKernel' \
    --launch-skip 16 --launch-count 8 -o $out/prof_$tag -f python bench.py --steps 1 --warmup 1 --no-cpu > $out/ncu_full_$tag.log 2>&1; echo "ncu full rc=$?"
ncu -i $out/prof_$tag.ncu-rep --page raw --csv > $out/prof_${tag}_raw.csv 2>/dev/null
ls -la $out
