#!/usr/bin/env python
"""Phase timeline of the 2-CTA W-stationary GEMM (MSDA_GEMM_WS2=1 MSDA_GEMM_WS_DBG=1): where a CTA's time goes.
    MSDA_GEMM_WS2=1 MSDA_GEMM_WS_DBG=1 python tools/gemm_timeline.py [M N K]"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from uninext_b200 import _cabi

lib = _cabi.load()
m, n, k = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (44646, 256, 256)
a = torch.randn(m, k, device="cuda"); w = torch.randn(n, k, device="cuda") * 0.1; b = torch.randn(n, device="cuda")
c = torch.empty(m, n, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for _ in range(5):
    assert lib.msda_linear_tf32_ex(a.data_ptr(), w.data_ptr(), b.data_ptr(), None, m, n, k, 0, c.data_ptr(), st) == 0
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
lib.msda_linear_tf32_ex(a.data_ptr(), w.data_ptr(), b.data_ptr(), None, m, n, k, 0, c.data_ptr(), st)
e1.record(); torch.cuda.synchronize()
buf = np.zeros((160, 24), dtype=np.uint64)
assert lib.msda_debug_gemm_timeline(buf.ctypes.data_as(ctypes.c_void_p), buf.size) == 0
clk = torch.cuda.clock_rate() if hasattr(torch.cuda, "clock_rate") else 1965
print(f"M={m} N={n} K={k}: event time {e0.elapsed_time(e1) * 1e3:.1f} us (L2-warm), SM clock assumed {clk} MHz")
live = buf[:, 1] > 0
t = buf[live].astype(np.float64)
gt = t[:, 0]
print(f"CTAs {live.sum()}, entry skew (globaltimer) {(gt.max() - gt.min()) / 1e3:.2f} us")
rel = (t[:, 1:] - t[:, 1:2]) / clk                     # us since the CTA's own entry
names = {1: "set-up done", 2: "first stage landed", 19: "producer done", 20: "MMA issue done", 21: "exit"}
for i in range(8):
    names[3 + 2 * i] = f"tile {i} accumulated"
    names[4 + 2 * i] = f"tile {i} stored"
for j in range(1, 23):
    col = rel[:, j]
    ok = t[:, j + 1] > 0
    if ok.sum() == 0:
        continue
    print(f"  {names.get(j, j):22s} CTAs {int(ok.sum()):3d}   median {np.median(col[ok]):6.2f} us   min {col[ok].min():6.2f}   max {col[ok].max():6.2f}")
