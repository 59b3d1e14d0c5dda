"""The callers of the op: encoder / decoder layers of the deformable transformer, with the reference's sub-module
names so its checkpoints load unchanged (deformable_transformer.py:321-361 encoder layer, :364-416 decoder layer),
plus ``DeformableStack`` -- 6 + 6 layers wired the way the reference transformer wires them
(reference points: deformable_transformer.py:280-292; decoder boxes as 4-d reference points: :457-465) -- which is the
unit of the whole-model frames/s measurement (backbone, heads, matcher and losses are outside the hot path).
"""
from __future__ import annotations

import torch
from torch import nn

from uninext_b200.functions.fused import add_layer_norm, linear_colsum

from .ms_deform_attn import MSDeformAttn, batched_value_proj, use_batched_value_proj


def _add_pos(x, pos):
    return x if pos is None else x + pos


def fp32_under_autocast(forward):
    """The reference decorates every forward on the path with ``@custom_fwd(cast_inputs=torch.float32)``
    (deformable_transformer.py:351,398; _dino.py:360,407,441,511): inside autocast the WHOLE layer -- FFN GEMMs
    included -- runs in fp32 with autocast off.  Mirrored here for the reference-compatible configuration
    (``op_dtype is None``); with ``op_dtype=torch.bfloat16`` (new, no reference counterpart) autocast is left on."""
    import functools

    @functools.wraps(forward)
    def wrapper(self, *args, **kwargs):
        if getattr(self, "op_dtype", None) is None and torch.is_autocast_enabled("cuda"):
            cast = lambda t: t.float() if torch.is_tensor(t) and t.is_cuda and t.is_floating_point() else t
            with torch.autocast("cuda", enabled=False):
                return forward(self, *[cast(a) for a in args], **{k: cast(v) for k, v in kwargs.items()})
        return forward(self, *args, **kwargs)
    return wrapper


def _activation(name):
    """deformable_transformer.py `_get_activation_fn`: relu / gelu / glu, anything else raises."""
    import torch.nn.functional as F
    if name not in ("relu", "gelu", "glu"):
        raise RuntimeError(f"activation should be relu/gelu, not {name}.")
    return getattr(F, name)


def _ffn(x, linear1, act_name, drop_a, linear2):
    """linear2(dropout(activation(linear1(x)))) -- forward_ffn of both layers (deformable_transformer.py:345-349,
    392-396).  ReLU (every UNINEXT config) rides in the first GEMM's epilogue."""
    if act_name == "relu":
        return linear_colsum(drop_a(linear_colsum(x, linear1, relu=True)), linear2)
    return linear_colsum(drop_a(_activation(act_name)(linear_colsum(x, linear1))), linear2)


class DeformableTransformerEncoderLayer(nn.Module):
    """Constructor arguments in the reference's order (deformable_transformer.py:322-325); ``op_dtype`` is new."""

    def __init__(self, d_model=256, d_ffn=1024, dropout=0.1, activation="relu", n_levels=4, n_heads=8, n_points=4,
                 op_dtype=None):
        super().__init__()
        _activation(activation)
        self.activation_name = activation
        self.op_dtype = op_dtype
        self.self_attn = MSDeformAttn(d_model, n_levels, n_heads, n_points, op_dtype=op_dtype)
        self.dropout1, self.norm1 = nn.Dropout(dropout), nn.LayerNorm(d_model)
        self.linear1, self.dropout2 = nn.Linear(d_model, d_ffn), nn.Dropout(dropout)
        self.linear2, self.dropout3 = nn.Linear(d_ffn, d_model), nn.Dropout(dropout)
        self.norm2 = nn.LayerNorm(d_model)

    @fp32_under_autocast
    def forward(self, src, pos, reference_points, spatial_shapes, level_start_index, padding_mask=None):
        att = self.self_attn(_add_pos(src, pos), reference_points, src, spatial_shapes, level_start_index, padding_mask)
        src = add_layer_norm(src, self.dropout1(att), self.norm1)
        ffn = _ffn(src, self.linear1, self.activation_name, self.dropout2, self.linear2)
        return add_layer_norm(src, self.dropout3(ffn), self.norm2)


class DeformableTransformerDecoderLayer(nn.Module):
    """Constructor arguments in the reference's order (deformable_transformer.py:365-368); ``op_dtype`` is new."""

    def __init__(self, d_model=256, d_ffn=1024, dropout=0.1, activation="relu", n_levels=4, n_heads=8, n_points=4,
                 op_dtype=None):
        super().__init__()
        _activation(activation)
        self.activation_name = activation
        self.op_dtype = op_dtype
        self.cross_attn = MSDeformAttn(d_model, n_levels, n_heads, n_points, op_dtype=op_dtype)
        self.dropout1, self.norm1 = nn.Dropout(dropout), nn.LayerNorm(d_model)
        self.self_attn = nn.MultiheadAttention(d_model, n_heads, dropout=dropout)
        self.dropout2, self.norm2 = nn.Dropout(dropout), nn.LayerNorm(d_model)
        self.linear1, self.dropout3 = nn.Linear(d_model, d_ffn), nn.Dropout(dropout)
        self.linear2, self.dropout4 = nn.Linear(d_ffn, d_model), nn.Dropout(dropout)
        self.norm3 = nn.LayerNorm(d_model)

    @fp32_under_autocast
    def forward(self, tgt, query_pos, reference_points, src, src_spatial_shapes, level_start_index,
                src_padding_mask=None, attn_masks=None, *, projected_value=None):
        """``attn_masks``: the DINO variant's self-attention mask over the queries (denoising groups must not see each
        other; deformable_transformer_dino.py:407-412).  The non-DINO layer (deformable_transformer.py:398-401) is
        the same call with ``attn_masks=None``.  ``projected_value`` (keyword-only, new): this layer's
        ``cross_attn.value_proj(src)`` computed outside, batched over all decoder layers (``batched_value_proj``)."""
        qk = _add_pos(tgt, query_pos).transpose(0, 1)
        sa = self.self_attn(qk, qk, tgt.transpose(0, 1), attn_mask=attn_masks)[0].transpose(0, 1)
        tgt = add_layer_norm(tgt, self.dropout2(sa), self.norm2)
        ca = self.cross_attn(_add_pos(tgt, query_pos), reference_points, src, src_spatial_shapes, level_start_index,
                             src_padding_mask, projected_value=projected_value)
        tgt = add_layer_norm(tgt, self.dropout1(ca), self.norm1)
        ffn = _ffn(tgt, self.linear1, self.activation_name, self.dropout3, self.linear2)
        return add_layer_norm(tgt, self.dropout4(ffn), self.norm3)


def encoder_reference_points(spatial_shapes_list, valid_ratios, device):
    """Pixel centres of every level, normalised and scaled by the valid ratios -> [N, S, L, 2]
    (deformable_transformer.py:280-292)."""
    pts = []
    for lvl, (h, w) in enumerate(spatial_shapes_list):
        ys = torch.arange(h, dtype=torch.float32, device=device) + 0.5
        xs = torch.arange(w, dtype=torch.float32, device=device) + 0.5
        yy, xx = torch.meshgrid(ys, xs, indexing="ij")
        y = yy.reshape(-1)[None] / (valid_ratios[:, None, lvl, 1] * h)
        x = xx.reshape(-1)[None] / (valid_ratios[:, None, lvl, 0] * w)
        pts.append(torch.stack((x, y), -1))
    ref = torch.cat(pts, 1)
    return ref[:, :, None] * valid_ratios[:, None]


class DeformableStack(nn.Module):
    """num_layers encoder layers over the flattened pyramid + num_layers decoder layers over `num_queries` box queries
    (UNINEXT defaults: d_model 256, d_ffn 2048, 8 heads, 4 levels, 4 points; uninext/config.py:156-174)."""

    def __init__(self, d_model=256, d_ffn=2048, n_heads=8, n_levels=4, n_points=4, num_layers=6, num_queries=300,
                 dropout=0.0, op_dtype=None):
        super().__init__()
        kw = dict(d_model=d_model, d_ffn=d_ffn, dropout=dropout, n_levels=n_levels, n_heads=n_heads,
                  n_points=n_points, op_dtype=op_dtype)
        self.encoder = nn.ModuleList(DeformableTransformerEncoderLayer(**kw) for _ in range(num_layers))
        self.decoder = nn.ModuleList(DeformableTransformerDecoderLayer(**kw) for _ in range(num_layers))
        self.level_embed = nn.Parameter(torch.randn(n_levels, d_model) * 0.02)
        self.query_embed = nn.Embedding(num_queries, d_model * 2)
        # decoder reference boxes (cx, cy, w, h): constants here, as with the reference's box refinement, which feeds each
        # layer detached boxes (deformable_transformer.py: `reference_points = new_reference_points.detach()`)
        g = torch.Generator().manual_seed(0)
        boxes = torch.cat((torch.rand(num_queries, 2, generator=g), 0.05 + 0.35 * torch.rand(num_queries, 2, generator=g)), -1)
        self.register_buffer("reference_boxes", boxes)
        self.n_levels = n_levels

    def forward(self, src, pos, spatial_shapes_list, spatial_shapes, level_start_index, padding_mask=None):
        """src, pos: [N, S, C] flattened multi-scale features / position encodings; padding_mask [N, S] bool (the reference
        always passes its `mask_flatten`, deformable_transformer.py:190-205).  Returns decoder output [N, Q, C]."""
        n = src.shape[0]
        valid = torch.ones(n, self.n_levels, 2, device=src.device)
        ref = encoder_reference_points(spatial_shapes_list, valid, src.device)
        # per-level embedding added to the position encoding (deformable_transformer.py:193-195 lvl_pos_embed)
        pos = pos + torch.cat([self.level_embed[i].view(1, 1, -1).expand(1, h * w, -1)
                               for i, (h, w) in enumerate(spatial_shapes_list)], 1)
        memory = src
        for layer in self.encoder:
            memory = layer(memory, pos, ref, spatial_shapes, level_start_index, padding_mask)
        qpos, tgt = self.query_embed.weight.chunk(2, dim=-1)
        qpos, tgt = qpos[None].expand(n, -1, -1), tgt[None].expand(n, -1, -1)
        boxes = self.reference_boxes[None].expand(n, -1, -1)                              # [N, Q, 4] (cx, cy, w, h)
        ref_dec = boxes[:, :, None] * torch.cat((valid, valid), -1)[:, None]             # deformable_transformer.py:457-459
        out = tgt
        if use_batched_value_proj():                                 # memory is loop-invariant: one [256 -> 6 x 256] GEMM
            values = batched_value_proj([layer.cross_attn for layer in self.decoder], memory, padding_mask)
        else:
            values = [None] * len(self.decoder)
        for layer, val in zip(self.decoder, values):
            # `projected_value` only when there is one: any layer with the reference's forward signature (e.g. the
            # reference's own class, bench.py's reference_stack leg) can stand in `self.decoder`
            extra = {} if val is None else {"projected_value": val}
            out = layer(out, qpos, ref_dec, memory, spatial_shapes, level_start_index, padding_mask, **extra)
        return out
