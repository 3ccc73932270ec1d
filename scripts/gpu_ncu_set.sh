#!/bin/bash
# `ncu --set full` capture of the kernels matching $1 (one launch each after warm-up), digests under gpurun_out/; tag $2
#   gpurun --timeout 900 -- 'bash scripts/gpu_ncu_set.sh "seedMmKernel|estimateStagedKernel" r2c'
re=${1:-seedMmKernel}; tag=${2:-one}; cnt=${3:-1}
out=gpurun_out; mkdir -p $out
python bench.py --steps 1 --warmup 1 --no-cpu > $out/plain_$tag.log 2>&1 &&
timeout 800 ncu --set full --clock-control none --import-source on -k regex:"$re" --launch-skip $((2*cnt)) --launch-count $cnt -o $out/prof_$tag -f \
    python bench.py --steps 1 --warmup 1 --no-cpu > $out/ncu_full_$tag.log 2>&1; echo "ncu rc=$?"
ncu -i $out/prof_$tag.ncu-rep --page raw --csv > $out/prof_${tag}_raw.csv 2>/dev/null
python profiles/stalls.py $out/prof_${tag}_raw.csv 2>/dev/null | tee $out/prof_${tag}_digest.txt | head -70
python - <<PY
import csv
rows = list(csv.reader(open("$out/prof_${tag}_raw.csv")))
H = rows[0]
for key in ("sm__pipe_tensor", "sm__inst_executed_pipe_tensor", "tensor", "tmem", "lts__t_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum"):
    for i, h in enumerate(H):
        if key in h and ("pct" in h or "sum" in h):
            print(h, [r[i] for r in rows[2:]])
PY
