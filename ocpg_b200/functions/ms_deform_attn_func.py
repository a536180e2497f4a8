"""``MSDeformAttnFunction`` -- same operator API as the reference's
models/ops/functions/ms_deform_attn_func.py:21-38:

    MSDeformAttnFunction.apply(value, value_spatial_shapes, value_level_start_index,
                               sampling_locations, attention_weights, im2col_step) -> output

Gradients flow to arguments 0, 3 and 4 only (:38); ``@once_differentiable`` (no double backward, :31).
The compiled module it calls is ocpg_b200.MultiScaleDeformableAttention (ctypes -> libmsda_sm100.so).
The reference's debug-only ``ms_deform_attn_core_pytorch`` (:41-61) is NOT part of the product; its
restatement lives in oracle/ as the checker.
"""
from __future__ import annotations

from torch.autograd import Function
from torch.autograd.function import once_differentiable

from .. import MultiScaleDeformableAttention as MSDA


class MSDeformAttnFunction(Function):
    @staticmethod
    def forward(ctx, value, value_spatial_shapes, value_level_start_index, sampling_locations,
                attention_weights, im2col_step):
        ctx.im2col_step = im2col_step
        output = MSDA.ms_deform_attn_forward(value, value_spatial_shapes, value_level_start_index,
                                             sampling_locations, attention_weights, ctx.im2col_step)
        ctx.save_for_backward(value, value_spatial_shapes, value_level_start_index, sampling_locations,
                              attention_weights)
        return output

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        value, shapes, start, sampling_locations, attention_weights = ctx.saved_tensors
        grad_value, grad_sampling_loc, grad_attn_weight = MSDA.ms_deform_attn_backward(
            value, shapes, start, sampling_locations, attention_weights, grad_output.contiguous(), ctx.im2col_step)
        return grad_value, None, None, grad_sampling_loc, grad_attn_weight, None
