#!/bin/bash
# Usage: gpurun -- 'bash scripts/gpu_pytest.sh <tag> <pytest args...>'
tag=$1; shift
mkdir -p gpurun_out
timeout 1200 python -m pytest -m gpu -x -q "$@" > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?"
tail -40 gpurun_out/pytest_$tag.log
