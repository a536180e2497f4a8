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

import torch
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


class MSDeformAttnFusedFunction(Function):
    """The operator with MSDeformAttn.forward's softmax and sampling-location arithmetic
    (models/ops/modules/ms_deform_attn.py:100-110) inside the kernels.

        apply(value, spatial_shapes, level_start_index, sampling_offsets, attention_logits, reference_points,
              im2col_step, emit) -> (output, sampling_locations, attention_weights)

    ``sampling_offsets`` (N, Lq, M, L, P, 2) and ``attention_logits`` (N, Lq, M, L*P) are the raw Linear
    outputs.  With ``emit=False`` the last two results are ``None`` (the encoder discards them,
    deformable_transformer.py:251); with ``emit=True`` they are materialised for the caller (the decoder reads
    them, :365-375) but carry no gradient -- use the unfused module if a loss depends on them.
    Gradients flow to value, sampling_offsets, attention_logits and reference_points."""

    @staticmethod
    def forward(ctx, value, spatial_shapes, level_start_index, offsets, logits, reference_points, im2col_step, emit):
        ctx.im2col_step = im2col_step
        output, loc, attn = MSDA.ms_deform_attn_fused_forward(value, spatial_shapes, level_start_index, offsets, logits,
                                                              reference_points, im2col_step, bool(emit))
        ctx.save_for_backward(value, spatial_shapes, level_start_index, offsets, logits, reference_points)
        if emit:
            ctx.mark_non_differentiable(loc, attn)
        return output, loc, attn

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output, _grad_loc, _grad_attn):
        value, shapes, start, offsets, logits, ref = ctx.saved_tensors
        need_ref = ctx.needs_input_grad[5]
        gv, g_off, g_logits, g_loc = MSDA.ms_deform_attn_fused_backward(
            value, shapes, start, offsets, logits, ref, grad_output.contiguous(), ctx.im2col_step, need_ref)
        g_ref = None
        if need_ref:
            # loc = ref_xy + offsets-term: d/d ref_xy = sum over heads and points; 4-d boxes also feed (w, h):
            # loc = ref_xy + offsets / P * ref_wh * 0.5  ->  d/d ref_wh = sum(g_loc * offsets / P * 0.5)
            g_xy = g_loc.sum(dim=(2, 4))                                           # (N, Lq, L, 2)
            if ref.shape[-1] == 2:
                g_ref = g_xy
            else:
                g_wh = (g_loc * (offsets / offsets.shape[4] * 0.5)).sum(dim=(2, 4))
                g_ref = torch.cat((g_xy, g_wh), -1)
        return gv, None, None, g_off, g_logits, g_ref, None, None
