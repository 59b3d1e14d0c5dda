#!/usr/bin/env python
"""Device time vs wall time of one DeformableStack training step at cfg2 (is the step launch-bound?)."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from uninext_b200.modules.deformable_layers import DeformableStack
from uninext_b200.workloads import CONFIGS, level_tensors
from torch.profiler import profile, ProfilerActivity

torch.backends.cuda.matmul.allow_tf32 = True
cfg = CONFIGS["cfg2"]; dev = "cuda"
ss, lsi = level_tensors(cfg.shapes, dev)
torch.manual_seed(0)
model = DeformableStack(num_layers=6, num_queries=cfg.dec_queries).to(dev)
src = torch.randn(cfg.batch, cfg.S, 256, device=dev)
pos = torch.randn(cfg.batch, cfg.S, 256, device=dev)

def step():
    for p in model.parameters(): p.grad = None
    model(src, pos, cfg.shapes, ss, lsi).square().mean().backward()

for _ in range(3): step()
torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(5): step()
torch.cuda.synchronize()
wall = (time.perf_counter() - t) / 5
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(2): step()
    torch.cuda.synchronize()
rows = sorted(((e.device_time_total / 2, e.count // 2, e.key) for e in prof.key_averages() if e.device_time_total > 0), reverse=True)
dev_ms = sum(r[0] for r in rows) / 1e3
print(f"wall {wall*1e3:.2f} ms/step, device busy {dev_ms:.2f} ms/step, kernels/step {sum(r[1] for r in rows)}")
for t_, c, k in rows[:28]:
    print(f"{t_:9.1f} us  x{c:<4d} {k[:100]}")
