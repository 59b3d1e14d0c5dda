#!/usr/bin/env python
"""Where does a MSDeformAttn module / encoder layer spend its time?  torch.profiler kernel table at cfg2 shapes."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from uninext_b200.modules import MSDeformAttn
from uninext_b200.modules.deformable_layers import DeformableTransformerEncoderLayer, encoder_reference_points
from uninext_b200.workloads import CONFIGS, level_tensors
from torch.profiler import profile, ProfilerActivity

what = sys.argv[1] if len(sys.argv) > 1 else "module"
tf32 = (sys.argv[2] == "tf32") if len(sys.argv) > 2 else True
torch.backends.cuda.matmul.allow_tf32 = tf32
cfg = CONFIGS["cfg2"]; dev = "cuda"
ss, lsi = level_tensors(cfg.shapes, dev)
torch.manual_seed(0)
src = torch.randn(cfg.batch, cfg.S, 256, device=dev, requires_grad=True)
pos = torch.randn(cfg.batch, cfg.S, 256, device=dev)
ref = encoder_reference_points(cfg.shapes, torch.ones(cfg.batch, 4, 2, device=dev), dev)
mod = (MSDeformAttn() if what == "module" else DeformableTransformerEncoderLayer(d_ffn=2048, dropout=0.0)).to(dev)

def step():
    if what == "module":
        out = mod(src + pos, ref, src, ss, lsi, None)
    else:
        out = mod(src, pos, ref, ss, lsi, None)
    out.square().mean().backward()

for _ in range(3): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3): step()
    torch.cuda.synchronize()
print(f"== {what} fwd+bwd, tf32={tf32}, per step (3 steps averaged)")
tab = prof.key_averages()
rows = sorted(((e.device_time_total / 3, e.count // 3, e.key) for e in tab if e.device_time_total > 0), reverse=True)
tot = sum(r[0] for r in rows)
print(f"total device time per step: {tot/1e3:.3f} ms in {sum(r[1] for r in rows)} kernels")
for t, c, k in rows[:22]:
    print(f"{t:9.1f} us  x{c:<3d} {k[:110]}")
