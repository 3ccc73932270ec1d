#!/bin/bash
# stream-overlap experiment: workspaces x literal-sort CTAs per SM
out=gpurun_out; mkdir -p $out
for cfg in "2 2" "2 1" "3 1" "4 1" "4 2" "3 2"; do
  set -- $cfg
  HU_LIT_CTAS_PER_SM=$2 timeout 600 python bench.py --no-cpu --steps 12 --workspaces $1 > $out/exp_ws$1_lit$2.json 2>/dev/null
  python - <<PY
import json
d=json.loads(open("$out/exp_ws$1_lit$2.json").read().strip().splitlines()[-1])
print("ws=$1 lit=$2", round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"],2), {k:round(v["ms"],2) for k,v in d["stages"].items()})
PY
done
