"""``MSDeformAttnFunction`` -- same autograd boundary as the reference
(ops/functions/ms_deform_attn_func.py:21-40): ``apply(value, value_spatial_shapes, value_level_start_index,
sampling_locations, attention_weights, im2col_step)``, gradients for arguments 0, 3 and 4 only.

Two classes:
* ``MSDeformAttnFunction``      -- reference behaviour, including ``custom_fwd(cast_inputs=float32)`` (func.py:23):
                                   inside autocast every floating input is cast to fp32.
* ``MSDeformAttnFunctionBF16``  -- new: keeps bf16 ``value`` / output (fp32 locations and weights, fp32 accumulate).
"""
from __future__ import annotations

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from uninext_b200.dropin import MultiScaleDeformableAttention as MSDA


class MSDeformAttnFunction(Function):
    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights,
                im2col_step):
        ctx.im2col_step = im2col_step
        output = MSDA.ms_deform_attn_forward(value, value_spatial_shapes, value_level_start_index, sampling_locations,
                                             attention_weights, ctx.im2col_step)
        ctx.save_for_backward(value, value_spatial_shapes, value_level_start_index, sampling_locations,
                              attention_weights)
        return output

    @staticmethod
    @once_differentiable
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, grad_output):
        value, shapes, lsi, loc, attn = ctx.saved_tensors
        grad_value, grad_loc, grad_attn = MSDA.ms_deform_attn_backward(
            value, shapes, lsi, loc, attn, grad_output.contiguous(), ctx.im2col_step)
        return grad_value, None, None, grad_loc, grad_attn, None


class MSDeformAttnFunctionBF16(Function):
    """bf16 storage for value / output / grad_output; sampling locations and attention weights are fp32."""

    @staticmethod
    def forward(ctx, value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights,
                im2col_step):
        ctx.im2col_step = im2col_step
        ctx.value_dtype = value.dtype            # an fp32 value_proj output gets the fp32 accumulator back, un-rounded
        value = value.to(torch.bfloat16)
        loc = sampling_locations.float().contiguous()
        attn = attention_weights.float().contiguous()
        ctx.loc_dtype, ctx.attn_dtype = sampling_locations.dtype, attention_weights.dtype
        output = MSDA.ms_deform_attn_forward(value, value_spatial_shapes, value_level_start_index, loc, attn,
                                             im2col_step)
        ctx.save_for_backward(value, value_spatial_shapes, value_level_start_index, loc, attn)
        return output

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        value, shapes, lsi, loc, attn = ctx.saved_tensors
        grad_value, grad_loc, grad_attn = MSDA.ms_deform_attn_backward(
            value, shapes, lsi, loc, attn, grad_output.to(torch.bfloat16).contiguous(), ctx.im2col_step,
            grad_value_dtype=torch.float32 if ctx.value_dtype == torch.float32 else None)
        return grad_value.to(ctx.value_dtype), None, None, grad_loc.to(ctx.loc_dtype), grad_attn.to(ctx.attn_dtype), None
