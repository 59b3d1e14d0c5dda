#!/bin/bash
# One GPU visit: parity tests, smoke, op micro-benchmarks, bench line, ncu launch list + full capture.
# Usage (from the repo root, under gpurun):  bash tools/gpu_round.sh <tag>
set -u
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
nproc > $OUT/nproc.txt
echo "== pytest -m gpu" ; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee $OUT/pytest_gpu.txt
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/smoke.txt
echo "== opbench"
for spec in "cfg2 enc fp32" "cfg2 dec fp32" "cfg2 enc bf16" "cfg3 enc bf16" "cfg4 enc fp32"; do
  set -- $spec
  timeout 300 python tools/opbench.py --config $1 --kind $2 --dtype $3 2>&1 | tail -1 | tee -a $OUT/opbench.txt
done
echo "== bench" ; timeout 900 python bench.py --steps 10 --warmup 3 2>&1 | tail -3 | tee $OUT/bench.txt
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:msda_ -c 60 --csv --log-file $OUT/launches.csv \
  python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/bench_under_ncu.log 2>&1
echo "== ncu full (enc fwd+bwd)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:msda_ -s 6 -c 2 -f -o $OUT/prof_enc_fp32 \
  python tools/opbench.py --config cfg2 --kind enc --dtype fp32 --iters 2 --warmup 3 > $OUT/ncu_enc.log 2>&1
ls -la $OUT
