import os, sys, torch
sys.path.insert(0, "/root/repo")
from uninext_b200 import _cabi
from uninext_b200.dropin import MultiScaleDeformableAttention as MSDA
from uninext_b200.workloads import CONFIGS, make_inputs
lib = _cabi.load()
cfg = CONFIGS["cfg3"]
inp = make_inputs(cfg, "enc", "cuda", dtype=torch.bfloat16, seed=0)
args = (inp["value"], inp["spatial_shapes"], inp["level_start_index"], inp["sampling_locations"], inp["attention_weights"])
for thr in (0, 8192, 1):
    lib.msda_set_knob(_cabi.KNOB_BF16_FINE_ROWS, thr)
    for _ in range(2):
        MSDA.ms_deform_attn_backward(*args, inp["grad_output"], 64)
    torch.cuda.synchronize()
