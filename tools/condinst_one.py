#!/usr/bin/env python
"""Three forward+backward passes of the CondInst head at 2 x 300 instances, 100 x 168 -- the ncu target:
    ncu --metrics gpu__time_duration.sum,sm__inst_executed_pipe_fma.sum,smsp__inst_executed.sum,dram__bytes_read.sum \
        --clock-control none -k regex:"condinst|aligned_bilinear" --csv --log-file out.csv python tools/condinst_one.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from uninext_b200.modules.dynamic_mask_head import dynamic_mask_with_coords
num_insts=[300,300]; hw=(100,168); total=600
g = torch.Generator().manual_seed(0)
feats = torch.randn(2, 8, *hw, generator=g).cuda().requires_grad_(True)
refs = (torch.rand(1, total, 2, generator=g) * torch.tensor([hw[1] * 8.0, hw[0] * 8.0])).cuda().requires_grad_(True)
params = (torch.randn(1, total, 169, generator=g) * 0.3).cuda().requires_grad_(True)
for _ in range(3):
    out = dynamic_mask_with_coords(feats, refs, params, num_insts, 8, True, 4)
    out.backward(torch.ones_like(out))
torch.cuda.synchronize()
