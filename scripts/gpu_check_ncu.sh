#!/bin/bash
# parity tests + bench + one --set full capture of selected kernels.  Usage: ... gpu_check_ncu.sh <tag> <kernel regex> [skip] [count]
bash scripts/gpu_check.sh $1
bash scripts/gpu_ncu.sh $1 "$2" ${3:-2} ${4:-3}
