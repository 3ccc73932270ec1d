#!/bin/bash
# ncu launch list (device time per launch) of a short bench run, optionally restricted to kernels matching $2
#   gpurun --timeout 900 -- 'bash scripts/gpu_launches.sh <tag> [kernel regex] [bench args]'
tag=${1:-l}; re=${2:-.}; shift; shift
out=gpurun_out; mkdir -p $out
python bench.py --steps 2 --warmup 1 --no-cpu "$@" > $out/plain_$tag.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$re" -c 400 --csv --log-file $out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu "$@" > $out/ncu_launch_$tag.log 2>&1; echo "ncu launches rc=$?"
python - <<PY
import csv, collections
rows = [r for r in csv.reader(open("$out/launches_$tag.csv")) if len(r) > 10 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    name = r[4].split("(")[0][:60]
    v = float(r[-1].replace(",", ""))
    unit = r[-2]
    if unit == "ns": v /= 1e6
    elif unit in ("us", "usecond"): v /= 1e3
    elif unit == "s": v *= 1e3
    agg.setdefault(name, []).append(v)
for k, v in agg.items():
    print("%-62s n=%4d  mean %8.3f ms  last %8.3f ms" % (k, len(v), sum(v) / len(v), v[-1]))
PY
