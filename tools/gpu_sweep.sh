#!/bin/bash
# A/B sweep over experimental env knobs. Usage: bash tools/gpu_sweep.sh <tag> "ENV=.. ENV=.." ...
set -u
TAG=$1; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
for envs in "$@"; do
  echo "== $envs" | tee -a $OUT/sweep.txt
  env $envs timeout 300 python tools/opbench.py --config cfg2 --kind enc --dtype fp32 2>&1 | tail -1 | tee -a $OUT/sweep.txt
done
