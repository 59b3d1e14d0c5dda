#!/usr/bin/env python
"""Error and time of the bf16 backward's grad_value for several MSDA_KNOB_BF16_FINE_ROWS thresholds, against the fp32
kernel on the same bf16-rounded inputs (full size):   python tools/bf16_mixed_error.py --config cfg3"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from uninext_b200 import _cabi  # noqa: E402
from uninext_b200.dropin import MultiScaleDeformableAttention as MSDA  # noqa: E402
from uninext_b200.workloads import CONFIGS, make_inputs  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="cfg3")
a = ap.parse_args()
cfg = CONFIGS[a.config]
lib = _cabi.load()
inp = make_inputs(cfg, "enc", "cuda", dtype=torch.bfloat16, seed=0)
args = (inp["value"], inp["spatial_shapes"], inp["level_start_index"], inp["sampling_locations"], inp["attention_weights"])
ref = MSDA.ms_deform_attn_backward(inp["value"].float(), *args[1:], inp["grad_output"].float(), 64)[0]      # fp32 truth
scale = ref.abs().max().item()
starts = inp["level_start_index"].tolist() + [cfg.S]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for thr in (0, 8192, 4096, 1024, 1):
    lib.msda_set_knob(_cabi.KNOB_BF16_FINE_ROWS, thr)
    ts = []
    for i in range(8):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        gv = MSDA.ms_deform_attn_backward(*args, inp["grad_output"], 64)[0]
        e1.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(e0.elapsed_time(e1))
    err = (gv.float() - ref).abs()
    row = {"config": cfg.name, "fine_rows": thr, "bwd_ms": round(sorted(ts)[len(ts) // 2], 4),
           "max_err_over_scale": round(err.max().item() / scale, 5), "levels": []}
    for l in range(len(starts) - 1):
        e, r = err[:, starts[l]:starts[l + 1]], ref[:, starts[l]:starts[l + 1]]
        row["levels"].append({"rows": starts[l + 1] - starts[l], "rms_err_over_rms": round((e.square().mean().sqrt() / r.square().mean().sqrt()).item(), 5),
                              "max_err_over_level_max": round((e.max() / r.abs().max()).item(), 5)})
    print(json.dumps(row))
lib.msda_set_knob(_cabi.KNOB_BF16_FINE_ROWS, 0)
