"""``MSDeformAttn`` nn.Module -- drop-in for the reference's models/ops/modules/ms_deform_attn.py:31-118.

Same constructor ``(d_model=256, n_levels=4, n_heads=8, n_points=4)``, same parameters and
state_dict keys (``sampling_offsets``, ``attention_weights``, ``value_proj``, ``output_proj``, each
``.weight`` / ``.bias``; :57-60) so reference checkpoints load unchanged, same initialisation (:62-78),
same ``forward`` signature and the same **3-tuple** return
``(output, sampling_locations, attention_weights)`` (:118 -- OCPG's decoder consumes the last two,
deformable_transformer.py:365-375).

Opt-in extension: ``module.fused = True`` runs the softmax and the
sampling-location arithmetic (:101-110) inside the kernels (MSDeformAttnFusedFunction) -- same parameters, same
result; ``module.emit_sampling = False`` additionally skips materialising ``sampling_locations`` /
``attention_weights`` (returned as ``None``), which is what the encoder layers want
(deformable_transformer.py:251 discards them).  The default (``fused = False``) is the reference's exact graph.

Host-side differences:
  * the sampling op is ocpg_b200's sm_100a kernels (through MSDeformAttnFunction);
  * the ``sum(H*W) == Len_in`` check (:94), which costs the reference one device->host sync per call, is
    done once per distinct ``input_spatial_shapes`` tensor and cached.
"""
from __future__ import annotations

import math
import warnings

import torch
import torch.nn.functional as F
from torch import nn

from .. import MultiScaleDeformableAttention as MSDA
from .. import epilogue
from .._shapes import shapes_on_host
from ..functions import MSDeformAttnFunction
from ..functions.ms_deform_attn_func import MSDeformAttnFusedFunction


def _is_power_of_2(n):
    if not isinstance(n, int) or n < 0:
        raise ValueError("invalid input for _is_power_of_2: {} (type: {})".format(n, type(n)))
    return n != 0 and (n & (n - 1)) == 0


class MSDeformAttn(nn.Module):
    def __init__(self, d_model=256, n_levels=4, n_heads=8, n_points=4):
        """Multi-scale deformable attention.

        :param d_model   hidden dimension
        :param n_levels  number of feature levels
        :param n_heads   number of attention heads
        :param n_points  number of sampling points per attention head per feature level
        """
        super().__init__()
        if d_model % n_heads != 0:
            raise ValueError("d_model must be divisible by n_heads, but got {} and {}".format(d_model, n_heads))
        if not _is_power_of_2(d_model // n_heads):
            warnings.warn("You'd better set d_model in MSDeformAttn to make the dimension of each attention head a "
                          "power of 2 which is more efficient in our CUDA implementation.")
        self.im2col_step = 64          # kept for API parity (reference :49); the kernels ignore it
        self.d_model, self.n_levels, self.n_heads, self.n_points = d_model, n_levels, n_heads, n_points

        self.sampling_offsets = nn.Linear(d_model, n_heads * n_levels * n_points * 2)
        self.attention_weights = nn.Linear(d_model, n_heads * n_levels * n_points)
        self.value_proj = nn.Linear(d_model, d_model)
        self.output_proj = nn.Linear(d_model, d_model)
        self.fused = False             # True: softmax + sampling locations inside the kernels
        self.emit_sampling = True      # fused only: materialise sampling_locations / attention_weights for the caller
        self._reset_parameters()

    def _reset_parameters(self):
        """Reference :62-78: zero offset weights, offset bias = one of n_heads compass directions
        (max-norm 1) scaled by the point index 1..n_points; zero attention logits; Xavier projections."""
        nn.init.constant_(self.sampling_offsets.weight.data, 0.0)
        angle = torch.arange(self.n_heads, dtype=torch.float32) * (2.0 * math.pi / self.n_heads)
        direction = torch.stack([angle.cos(), angle.sin()], -1)
        direction = direction / direction.abs().max(-1, keepdim=True)[0]
        grid = direction.view(self.n_heads, 1, 1, 2).repeat(1, self.n_levels, self.n_points, 1)
        grid = grid * torch.arange(1, self.n_points + 1, dtype=torch.float32).view(1, 1, self.n_points, 1)
        with torch.no_grad():
            self.sampling_offsets.bias = nn.Parameter(grid.reshape(-1))
        nn.init.constant_(self.attention_weights.weight.data, 0.0)
        nn.init.constant_(self.attention_weights.bias.data, 0.0)
        nn.init.xavier_uniform_(self.value_proj.weight.data)
        nn.init.constant_(self.value_proj.bias.data, 0.0)
        nn.init.xavier_uniform_(self.output_proj.weight.data)
        nn.init.constant_(self.output_proj.bias.data, 0.0)

    def _check_shapes(self, spatial_shapes, len_in):
        """Reference :94 -- ``assert (H * W).sum() == Len_in`` -- on the host copy of the shapes: one device->host copy per
        shapes tensor object and version (none at all for tensors built by ``flatten_levels``), not one per call."""
        total = sum(h * w for h, w in shapes_on_host(spatial_shapes))
        assert total == len_in, f"sum(H*W)={total} != Len_in={len_in}"

    def forward(self, query, reference_points, input_flatten, input_spatial_shapes, input_level_start_index,
                input_padding_mask=None):
        """See attend(); applies ``output_proj`` (:116) to its first result."""
        output, sampling_locations, weights = self.attend(query, reference_points, input_flatten, input_spatial_shapes,
                                                          input_level_start_index, input_padding_mask)
        return self.output_proj(output), sampling_locations, weights                     # :116, :118

    def attend(self, query, reference_points, input_flatten, input_spatial_shapes, input_level_start_index,
               input_padding_mask=None):
        """Everything of the reference's forward (:92-115) up to, not including, ``output_proj`` -- so that a caller can
        fold that projection's bias into its own epilogue (ocpg_b200/encoder.py).

        :param query                    (N, Length_query, C)
        :param reference_points         (N, Length_query, n_levels, 2) in [0, 1], top-left (0,0), bottom-right (1,1),
                                        including padding area; or (N, Length_query, n_levels, 4) = boxes (cx, cy, w, h)
        :param input_flatten            (N, sum_l H_l*W_l, C)
        :param input_spatial_shapes     (n_levels, 2) [(H_0, W_0), ...]
        :param input_level_start_index  (n_levels,)
        :param input_padding_mask       (N, sum_l H_l*W_l), True for padding elements

        :return (output before output_proj (N, Length_query, C), sampling_locations (N, Lq, M, L, P, 2),
                 attention_weights (N, Lq, M, L, P))
        """
        N, Len_q, _ = query.shape
        N, Len_in, _ = input_flatten.shape
        self._check_shapes(input_spatial_shapes, Len_in)
        M, L, P = self.n_heads, self.n_levels, self.n_points

        # fused mode: the Linears' bias gradients come from the streaming column-sum kernel (ocpg_b200/epilogue.py)
        lin = (lambda layer, x: epilogue.linear(x, layer.weight, layer.bias)) if self.fused else (lambda layer, x: layer(x))
        value = lin(self.value_proj, input_flatten)                                      # :96
        if input_padding_mask is not None:
            value = value.masked_fill(input_padding_mask[..., None], float(0))           # :97-98
        value = value.view(N, Len_in, M, self.d_model // M)
        offsets = lin(self.sampling_offsets, query).view(N, Len_q, M, L, P, 2)           # :100
        if reference_points.shape[-1] not in (2, 4):
            raise ValueError("Last dim of reference_points must be 2 or 4, but get {} instead.".format(
                reference_points.shape[-1]))
        if self.fused and value.is_cuda and MSDA.fused_supported(value, L, P):
            logits = lin(self.attention_weights, query).view(N, Len_q, M, L * P)
            # the fused kernels take fp32 offsets / logits / reference points whatever the dtype of `value` (a bf16 module
            # produces bf16 ones): the casts keep the gradient flowing back in the module's dtype
            return MSDeformAttnFusedFunction.apply(
                value.contiguous(), input_spatial_shapes, input_level_start_index, offsets.float().contiguous(),
                logits.float().contiguous(), reference_points.to(torch.float32).contiguous(), self.im2col_step, self.emit_sampling)
        weights = F.softmax(lin(self.attention_weights, query).view(N, Len_q, M, L * P), -1)  # :101-102
        weights = weights.view(N, Len_q, M, L, P)
        if reference_points.shape[-1] == 2:                                              # :104-107
            wh = torch.stack([input_spatial_shapes[..., 1], input_spatial_shapes[..., 0]], -1)
            sampling_locations = reference_points[:, :, None, :, None, :] + offsets / wh[None, None, None, :, None, :]
        elif reference_points.shape[-1] == 4:                                            # :108-110
            sampling_locations = reference_points[:, :, None, :, None, :2] \
                + offsets / P * reference_points[:, :, None, :, None, 2:] * 0.5
        else:
            raise ValueError("Last dim of reference_points must be 2 or 4, but get {} instead.".format(
                reference_points.shape[-1]))
        output = MSDeformAttnFunction.apply(value.contiguous(), input_spatial_shapes, input_level_start_index,
                                            sampling_locations.contiguous(), weights.contiguous(), self.im2col_step)
        return output, sampling_locations, weights
