"""Synthetic workloads for the MSDeformAttn hot path (BASELINE.json configs; SURVEY.md section 8 shape table).

Geometry follows the reference pipeline: inputs padded to a multiple of 32 (uninext/util/misc.py:298-302), feature
levels at strides 8/16/32 plus one extra stride-2 3x3 conv level (deformable_detr.py:123-127); M=8 heads, D=32,
L=4 levels, P=4 points (uninext/config.py:156,170-174).

Sampling locations mimic what the module produces (ops/modules/ms_deform_attn.py:62-70,103-109):
  encoder -- query i is pixel i of the flattened pyramid; reference point = its centre, in every level
             (deformable_transformer.py:280-292 with valid_ratio 1); offsets = the module's ring initialisation
             (head k points in direction k, point p at distance p+1 pixels) plus Gaussian jitter, divided by (W_l, H_l).
  decoder -- reference boxes (cx, cy, w, h); loc = c + off / P * wh * 0.5 with off ~ N(0,1).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch


def _ceil_div(a, b):
    return -(-a // b)


def pyramid(height: int, width: int, pad: int = 32):
    """Level shapes [(H_l, W_l)] x4 for an image of (height, width)."""
    ph, pw = _ceil_div(height, pad) * pad, _ceil_div(width, pad) * pad
    shapes = [(_ceil_div(ph, s), _ceil_div(pw, s)) for s in (8, 16, 32)]
    h3, w3 = shapes[-1]
    shapes.append(((h3 - 1) // 2 + 1, (w3 - 1) // 2 + 1))      # 3x3 conv, stride 2, padding 1
    return shapes


@dataclass(frozen=True)
class OpConfig:
    name: str
    height: int
    width: int
    batch: int            # frames per GPU
    dec_queries: int
    heads: int = 8
    head_dim: int = 32
    points: int = 4

    @property
    def shapes(self):
        return pyramid(self.height, self.width)

    @property
    def S(self):
        return sum(h * w for h, w in self.shapes)

    def samples(self, kind: str) -> int:
        lq = self.S if kind == "enc" else self.dec_queries
        return self.batch * lq * self.heads * len(self.shapes) * self.points


# BASELINE.json `configs` (index = position in that list)
CONFIGS = {
    "cfg1": OpConfig("cfg1_320x320_q100", 320, 320, 1, 100),
    "cfg2": OpConfig("cfg2_coco_1333x800_r50_q300_b2", 800, 1333, 2, 300),
    "cfg3": OpConfig("cfg3_convnextL_1536x1024_q900_b2", 1024, 1536, 2, 900),
    "cfg4": OpConfig("cfg4_video_5x640x360_q300", 360, 640, 5, 300),
    "cfg5": OpConfig("cfg5_vith_1333x800_q300_b1", 800, 1333, 1, 300),
}


def level_tensors(shapes, device):
    ss = torch.as_tensor(shapes, dtype=torch.long, device=device)
    lsi = torch.cat((ss.new_zeros((1,)), ss.prod(1).cumsum(0)[:-1]))
    return ss, lsi


def encoder_reference_points(shapes, device, dtype=torch.float32):
    """[S, 2] pixel-centre (x, y) in [0,1] for every pixel of the flattened pyramid."""
    pts = []
    for h, w in shapes:
        ys = (torch.arange(h, device=device, dtype=dtype) + 0.5) / h
        xs = (torch.arange(w, device=device, dtype=dtype) + 0.5) / w
        yy, xx = torch.meshgrid(ys, xs, indexing="ij")
        pts.append(torch.stack((xx.reshape(-1), yy.reshape(-1)), -1))
    return torch.cat(pts, 0)


def ring_offsets(heads, levels, points, device, dtype=torch.float32):
    """The module's sampling_offsets bias (ms_deform_attn.py:64-70), in pixels: [M, L, P, 2]."""
    th = torch.arange(heads, device=device, dtype=dtype) * (2.0 * math.pi / heads)
    g = torch.stack((th.cos(), th.sin()), -1)
    g = g / g.abs().max(-1, keepdim=True)[0]
    g = g.view(heads, 1, 1, 2).repeat(1, levels, points, 1)
    g = g * torch.arange(1, points + 1, device=device, dtype=dtype).view(1, 1, points, 1)
    return g


def make_inputs(cfg: OpConfig, kind: str, device, dtype=torch.float32, seed: int = 0, jitter_px: float = 2.0,
                wild_fraction: float = 0.0):
    """Returns dict(value, spatial_shapes, level_start_index, sampling_locations, attention_weights, grad_output).

    kind: "enc" (Lq = S, pixel queries) or "dec" (Lq = cfg.dec_queries, box queries).
    wild_fraction: fraction of taps thrown uniformly into [-0.5, 1.5] to exercise the border predicates.
    """
    g = torch.Generator(device=device).manual_seed(seed)
    shapes = cfg.shapes
    L, M, D, P, N = len(shapes), cfg.heads, cfg.head_dim, cfg.points, cfg.batch
    ss, lsi = level_tensors(shapes, device)
    S = cfg.S
    f32 = torch.float32
    value = torch.randn(N, S, M, D, generator=g, device=device, dtype=f32).to(dtype)
    wh = torch.as_tensor([(w, h) for h, w in shapes], device=device, dtype=f32)          # [L, 2] = (W_l, H_l)
    if kind == "enc":
        Lq = S
        ref = encoder_reference_points(shapes, device)                                    # [S, 2]
        off = ring_offsets(M, L, P, device) + jitter_px * torch.randn(N, Lq, M, L, P, 2, generator=g, device=device)
        loc = ref.view(1, Lq, 1, 1, 1, 2) + off / wh.view(1, 1, 1, L, 1, 2)
    elif kind == "dec":
        Lq = cfg.dec_queries
        ctr = torch.rand(N, Lq, 1, 1, 1, 2, generator=g, device=device)
        box = 0.05 + 0.35 * torch.rand(N, Lq, 1, 1, 1, 2, generator=g, device=device)
        off = torch.randn(N, Lq, M, L, P, 2, generator=g, device=device)
        loc = ctr + off / P * box * 0.5
    else:
        raise ValueError(kind)
    if wild_fraction > 0:
        wild = torch.rand(loc.shape[:-1], generator=g, device=device) < wild_fraction
        rnd = torch.rand(loc.shape, generator=g, device=device) * 2.0 - 0.5
        loc = torch.where(wild[..., None], rnd, loc)
    logits = torch.randn(N, Lq, M, L * P, generator=g, device=device)
    attn = torch.softmax(logits, -1).view(N, Lq, M, L, P)
    grad_out = torch.randn(N, Lq, M * D, generator=g, device=device, dtype=f32).to(dtype)
    aux = torch.float64 if dtype == torch.float64 else f32
    return dict(value=value.contiguous(), spatial_shapes=ss, level_start_index=lsi,
                sampling_locations=loc.to(aux).contiguous(), attention_weights=attn.to(aux).contiguous(),
                grad_output=grad_out.contiguous())


def algorithmic_bytes(cfg: OpConfig, kind: str, elem: int, direction: str) -> int:
    """Compulsory traffic per op call (SURVEY.md section 8d / BASELINE.md section 3).
    fwd = value + loc + attn + out ;  bwd = fwd bytes (value, loc, attn, grad_out reads) + grad_value zero-fill and
    final write + grad_loc + grad_attn writes.  `elem` = bytes per value/out element; loc/attn are fp32."""
    N, S, M, D = cfg.batch, cfg.S, cfg.heads, cfg.head_dim
    lq = S if kind == "enc" else cfg.dec_queries
    smp = cfg.samples(kind)
    fwd = N * S * M * D * elem + smp * 8 + smp * 4 + N * lq * M * D * elem
    if direction == "fwd":
        return fwd
    gv_elem = 4 if elem == 2 else elem          # bf16 accumulates grad_value in fp32
    bwd = fwd + 2 * N * S * M * D * gv_elem + smp * 8 + smp * 4
    if direction == "bwd":
        return bwd
    return fwd + bwd
