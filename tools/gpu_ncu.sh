#!/bin/bash
# ncu full captures of the encoder-shaped fwd+bwd kernels under several settings.
# Usage: bash tools/gpu_ncu.sh <tag> "<name>|<env>|<opbench args>" ...
set -u
TAG=$1; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
for spec in "$@"; do
  IFS='|' read -r NAME ENVS ARGS <<< "$spec"
  echo "== ncu $NAME ($ENVS) $ARGS"
  env $ENVS timeout 600 ncu --set full --clock-control none --import-source on -k regex:msda_ -s 6 -c 2 -f -o $OUT/$NAME \
    python tools/opbench.py $ARGS --iters 2 --warmup 3 > $OUT/$NAME.log 2>&1
  tail -2 $OUT/$NAME.log
done
ls -la $OUT
