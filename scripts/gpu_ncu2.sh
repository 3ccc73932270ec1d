#!/bin/bash
# `ncu --set full` of the kernels whose names match the regexes given (one launch each), raw pages as CSV + stall digest
tag=${1:-n}; shift
mkdir -p gpurun_out
for k in "$@"; do
  ncu --set full --clock-control none --import-source on -k regex:"$k" --launch-skip 2 --launch-count 1 -o gpurun_out/prof_${tag}_$k -f \
      python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_full_${tag}_$k.log 2>&1; echo "ncu $k rc=$?"
  ncu -i gpurun_out/prof_${tag}_$k.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_${k}_raw.csv 2>/dev/null
  python profiles/stalls.py gpurun_out/prof_${tag}_${k}_raw.csv 2>/dev/null > gpurun_out/prof_${tag}_${k}_digest.txt
done
