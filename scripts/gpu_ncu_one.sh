#!/bin/bash
# one `ncu --set full` capture of one kernel (regex $1, tag $2), raw page as CSV
k=${1:-seedSortLiteralKernel}; tag=${2:-one}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"$k" --launch-skip 2 --launch-count 1 -o gpurun_out/prof_$tag -f \
    python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_full_$tag.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/prof_$tag.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_raw.csv 2>/dev/null
python profiles/stalls.py gpurun_out/prof_${tag}_raw.csv 2>/dev/null | head -60
