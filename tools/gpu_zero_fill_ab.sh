#!/bin/bash
# A/B of MSDA_KNOB_ZERO_FILL on a B200 (gpurun): three short op-only bench runs (0 = cudaMemsetAsync, 1 = msda_zero_fill,
# 2 = msda_zero_fill as PDL primary of the backward kernel), then the full -m gpu suite and the full bench line under the
# fastest setting (a non-zero level must win by >= 0.5 % to be picked).  Output: gpurun_out/$1/.
set -u
OUT=gpurun_out/${1:-zero_fill_ab}
mkdir -p "$OUT"
for lvl in 0 1 2 0 1 2; do
  MSDA_ZERO_FILL=$lvl timeout 120 python bench.py --steps 20 --warmup 5 --no-e2e --no-frames --no-cpu-baseline \
      --no-reference-cuda --no-configs 2>> "$OUT/quick.err" | tail -1 >> "$OUT/quick_$lvl.jsonl"
done
best=$(python - "$OUT" 2> "$OUT/quick_summary.txt" <<'PY'
import json, sys
out = sys.argv[1]
val = {}
for lvl in (0, 1, 2):
    xs = []
    try:
        for ln in open(f"{out}/quick_{lvl}.jsonl"):
            ln = ln.strip()
            if ln.startswith("{"):
                xs.append(json.loads(ln))
    except OSError:
        pass
    if xs:
        val[lvl] = max(x["value"] for x in xs)
        print(lvl, [(x["value"], x["kernels_ms"]) for x in xs], file=sys.stderr)
best = 0
for lvl in (1, 2):
    if lvl in val and 0 in val and val[lvl] >= 1.005 * val[0] and val[lvl] > val.get(best, 0):
        best = lvl
print(best)
PY
)
echo "best level: $best" | tee -a "$OUT/quick_summary.txt"
cat "$OUT/quick_summary.txt"
MSDA_ZERO_FILL=$best timeout 400 python -m pytest tests -m gpu -x -q > "$OUT/pytest_gpu.txt" 2>&1
tail -3 "$OUT/pytest_gpu.txt"
MSDA_ZERO_FILL=$best timeout 200 python bench.py --steps 20 --warmup 5 > "$OUT/bench.json" 2> "$OUT/bench.err"
tail -c 300 "$OUT/bench.json"
