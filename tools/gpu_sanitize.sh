#!/bin/bash
# compute-sanitizer memcheck + racecheck + initcheck over small forward/backward calls of every kernel family.
set -u
OUT=gpurun_out/${1:-san}
mkdir -p $OUT
cat > /tmp/san_driver.py <<'PY'
import sys, torch
sys.path.insert(0, ".")
from uninext_b200.dropin import MultiScaleDeformableAttention as MSDA
from uninext_b200.workloads import CONFIGS, make_inputs
import dataclasses
cfg = CONFIGS["cfg1"]
for kind in ("enc", "dec"):
    for dt in (torch.float32, torch.bfloat16, torch.float64):
        inp = make_inputs(cfg, kind, "cuda", dtype=dt, seed=1, wild_fraction=0.2)
        a = (inp["value"], inp["spatial_shapes"], inp["level_start_index"], inp["sampling_locations"], inp["attention_weights"])
        out = MSDA.ms_deform_attn_forward(*a, 64)
        g = MSDA.ms_deform_attn_backward(*a, inp["grad_output"], 64)
torch.cuda.synchronize()
print("driver done", float(out.float().abs().sum()))
PY
for tool in memcheck racecheck initcheck; do
  echo "== $tool"
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 --kernel-regex kns=msda python /tmp/san_driver.py > $OUT/$tool.log 2>&1
  echo "rc=$?" | tee -a $OUT/$tool.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|driver done|Error|hazard" $OUT/$tool.log | head -8
done
