#!/bin/bash
set -u
TAG=${1:-r}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu" ; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench" ; timeout 900 python bench.py --steps 20 --warmup 5 2>&1 | tail -2 | tee $OUT/bench.txt
echo "== bench reference arm" ; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | tee $OUT/bench_ref.txt
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:msda_ -c 60 --csv --log-file $OUT/launches.csv \
  python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-frames > $OUT/bench_under_ncu.log 2>&1
tail -3 $OUT/launches.csv
