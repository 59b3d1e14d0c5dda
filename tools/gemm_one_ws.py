#!/usr/bin/env python
"""ncu target: value_proj-shaped GEMM (44 646 x 256 x 256) five times on the W-stationary tcgen05 kernel and on cuBLAS TF32.
    ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"linear_tf32|gemm|cutlass" --csv python tools/gemm_one_ws.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from uninext_b200 import _cabi

lib = _cabi.load()
torch.backends.cuda.matmul.allow_tf32 = True
m, n, k = 44646, 256, 256
a = torch.randn(m, k, device="cuda"); w = torch.randn(n, k, device="cuda") * 0.1; b = torch.randn(n, device="cuda")
c = torch.empty(m, n, device="cuda")
mask = (torch.rand(m, device="cuda") < 0.1).to(torch.uint8)
st = torch.cuda.current_stream().cuda_stream
for _ in range(5):
    assert lib.msda_linear_tf32_ex(a.data_ptr(), w.data_ptr(), b.data_ptr(), mask.data_ptr(), m, n, k, 0, c.data_ptr(), st) == 0
    torch.addmm(b, a, w.t(), out=c)
torch.cuda.synchronize()
