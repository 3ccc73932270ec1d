#!/bin/bash
# one `--set full` capture of selected kernels.  Usage: gpurun -- 'bash scripts/gpu_ncu.sh <tag> <kernel regex> [skip] [count]'
tag=$1; rx=$2; skip=${3:-4}; cnt=${4:-2}
out=gpurun_out
mkdir -p $out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"$rx" --launch-skip $skip --launch-count $cnt \
    -o $out/prof_$tag -f python bench.py --steps 1 --warmup 1 --no-cpu > $out/ncu_full_$tag.log 2>&1; echo "ncu rc=$?"
ncu -i $out/prof_$tag.ncu-rep --page raw --csv > $out/prof_${tag}_raw.csv 2>/dev/null
ncu -i $out/prof_$tag.ncu-rep --page source --csv --print-source sass > $out/prof_${tag}_sass.csv 2>/dev/null
ls -la $out | tail -5
