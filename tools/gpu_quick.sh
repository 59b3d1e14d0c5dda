#!/bin/bash
# Quick GPU visit: parity tests + op micro-benchmarks (A/B via env), optional ncu capture.
# Usage: bash tools/gpu_quick.sh <tag> [ncu]
set -u
TAG=${1:-q}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu" ; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee $OUT/pytest_gpu.txt
echo "== opbench"
for spec in "cfg2 enc fp32" "cfg2 dec fp32" "cfg2 enc bf16" "cfg3 enc bf16" "cfg4 enc fp32"; do
  set -- $spec
  timeout 300 python tools/opbench.py --config $1 --kind $2 --dtype $3 2>&1 | tail -1 | tee -a $OUT/opbench.txt
done
echo "== opbench, linear slot order (MSDA_NO_PATCHES=1)"
MSDA_NO_PATCHES=1 timeout 300 python tools/opbench.py --config cfg2 --kind enc --dtype fp32 2>&1 | tail -1 | tee -a $OUT/opbench.txt
echo "== opbench, jitter 0.5 px / 4 px"
timeout 300 python tools/opbench.py --config cfg2 --kind enc --dtype fp32 --jitter 0.5 2>&1 | tail -1 | tee -a $OUT/opbench.txt
timeout 300 python tools/opbench.py --config cfg2 --kind enc --dtype fp32 --jitter 4 2>&1 | tail -1 | tee -a $OUT/opbench.txt
echo "== reference CUDA kernels on this GPU"
timeout 300 python tools/opbench.py --config cfg2 --kind enc --dtype fp32 --ref 2>&1 | tail -1 | tee -a $OUT/opbench.txt
timeout 300 python tools/opbench.py --config cfg2 --kind dec --dtype fp32 --ref 2>&1 | tail -1 | tee -a $OUT/opbench.txt
if [ "${2:-}" = "ncu" ]; then
  echo "== ncu full (enc fwd+bwd)"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:msda_ -s 6 -c 2 -f -o $OUT/prof_enc_fp32 \
    python tools/opbench.py --config cfg2 --kind enc --dtype fp32 --iters 2 --warmup 3 > $OUT/ncu_enc.log 2>&1
  tail -3 $OUT/ncu_enc.log
fi
