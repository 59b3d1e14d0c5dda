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
# round 2: slab-ordered kernels (shared-memory window + lists: the racecheck target), lane-shape and mixed-bf16 knobs
from uninext_b200 import _cabi
lib = _cabi.load()
for knob, val, dts in ((_cabi.KNOB_SLAB, 1, (torch.float32, torch.bfloat16)), (_cabi.KNOB_SLAB, 2, (torch.float32, torch.bfloat16)),
                       (_cabi.KNOB_F32_VEC8_FWD, 1, (torch.float32,)),
                       (_cabi.KNOB_F32_VEC8_BWD, 1, (torch.float32,)), (_cabi.KNOB_BF16_FINE_ROWS, 500, (torch.bfloat16,))):
    old = lib.msda_set_knob(knob, val)
    for dt in dts:
        inp = make_inputs(cfg, "enc", "cuda", dtype=dt, seed=2, wild_fraction=0.2)
        a = (inp["value"], inp["spatial_shapes"], inp["level_start_index"], inp["sampling_locations"], inp["attention_weights"])
        out = MSDA.ms_deform_attn_forward(*a, 64)
        g = MSDA.ms_deform_attn_backward(*a, inp["grad_output"], 64)
    lib.msda_set_knob(knob, old)
# CondInst dynamic mask head + aligned_bilinear, forward and backward
from uninext_b200.modules.dynamic_mask_head import dynamic_mask_with_coords
feats = torch.randn(2, 8, 12, 20, device="cuda", requires_grad=True)
refs = (torch.rand(1, 21, 2, device="cuda") * 100).requires_grad_(True)
params = (torch.randn(1, 21, 169, device="cuda") * 0.3).requires_grad_(True)
dynamic_mask_with_coords(feats, refs, params, [17, 4], 8).square().mean().backward()
feats = torch.randn(2, 8, 9, 11, device="cuda", requires_grad=True)        # odd sizes: scalar tails, generic aligned_bilinear
dynamic_mask_with_coords(feats, refs, params, [0, 21], 8).square().mean().backward()
dynamic_mask_with_coords(feats, refs, params, [20, 1], 8, True, 2).square().mean().backward()
# W-stationary tcgen05 GEMM with mask + ReLU
from uninext_b200.functions.fused import tcgen05_linear_ex
c = tcgen05_linear_ex(torch.randn(300, 256, device="cuda"), torch.randn(256, 256, device="cuda"), torch.randn(256, device="cuda"),
                      torch.rand(300, device="cuda") < 0.3, True)
torch.cuda.synchronize()
print("driver done", float(out.float().abs().sum()), float(c.abs().sum()))
PY
for tool in memcheck racecheck; do
  echo "== $tool"
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 --kernel-regex kns=msda --kernel-regex kns=gemm python /tmp/san_driver.py > $OUT/$tool.log 2>&1
  echo "rc=$?" | tee -a $OUT/$tool.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|driver done|Error|hazard" $OUT/$tool.log | head -8
done
