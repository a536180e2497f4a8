"""The decoder-side callers of the hot path: OCPG's deformable-transformer decoder, re-hosted around ocpg_b200.MSDeformAttn.

Reference: models/deformable_transformer.py -- ``DeformableTransformerDecoderLayer`` (:293-340) and
``DeformableTransformerDecoder`` (:344-398).  Same constructor arguments, same sub-module names (``cross_attn``,
``self_attn``, ``norm1..3``, ``linear1/2``, ``layers.N``, ``bbox_embed``, ``class_embed``) so a reference state_dict loads
unchanged, same forward signatures and results.  This is the harness of BASELINE.json configs[3] (cross-attention with
5 object queries per frame over the full multi-level memory: the low-query-count regime) at layer level, and it hosts
SURVEY.md section 8f rank 3 -- the consumers of the attention's returned locations and weights:

  * ``scale_reference_points``  reference_points[:, :, None] * valid_ratios[:, None] (twice for boxes)   (:358-363), one launch
  * ``select_top_samples``      sampling_locations / valid_ratios, top-30 of the M*L*P weights, gather   (:368-375), one launch
                                instead of a division over every point, a sort-based topk and a repeat + gather.

Both are autograd Functions on the C ABI (``msda_decoder_*``, include/msda_sm100.h); the gradients the reference graph
defines for them (d reference_points; d sampling_locations through the gather) are tiny and use torch ops.
"""
from __future__ import annotations

import copy

import torch
import torch.nn.functional as F
from torch import nn
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _lib, _strict, epilogue
from .encoder import _activation
from .modules import MSDeformAttn

TOP_SAMPLES = 30          # the reference's hard-coded "hyperpara 30" (:371-372)


def _stream(t) -> int:
    return _lib.raw_stream(t.device)


def _native_ok(*tensors) -> bool:
    return all(t.is_cuda and t.dtype == torch.float32 for t in tensors) and not torch.is_autocast_enabled()


class _ScaleReferencePoints(Function):
    @staticmethod
    def forward(ctx, reference_points, valid_ratios):
        N, Lq, rd = reference_points.shape
        L = valid_ratios.shape[1]
        ref, vr = reference_points.contiguous(), valid_ratios.contiguous()
        out = torch.empty(N, Lq, L, rd, dtype=torch.float32, device=ref.device)
        with _lib.on_device(ref.device):
            rc = _lib.lib().msda_decoder_reference_points_f32(ref.data_ptr(), vr.data_ptr(), N, Lq, L, rd, out.data_ptr(), _stream(ref))
        _lib.check(rc, "msda_decoder_reference_points_f32")
        ctx.save_for_backward(vr)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        (vr,) = ctx.saved_tensors
        scale = vr if g.shape[-1] == 2 else torch.cat([vr, vr], -1)
        return (g * scale[:, None]).sum(2), None


def scale_reference_points(reference_points: torch.Tensor, valid_ratios: torch.Tensor) -> torch.Tensor:
    """(N, Lq, 2|4) x (N, L, 2) -> (N, Lq, L, 2|4): the decoder's ``reference_points_input`` (:358-363)."""
    if reference_points.shape[-1] not in (2, 4):
        raise AssertionError("reference_points must have 2 or 4 coordinates")                       # :362
    if _native_ok(reference_points, valid_ratios):
        return _ScaleReferencePoints.apply(reference_points, valid_ratios)
    _strict.note_fallback("scale_reference_points", "needs fp32 CUDA tensors outside autocast")
    scale = valid_ratios if reference_points.shape[-1] == 2 else torch.cat([valid_ratios, valid_ratios], -1)
    return reference_points[:, :, None] * scale[:, None]


class _SelectTopSamples(Function):
    @staticmethod
    def forward(ctx, sampling_locations, attention_weights, valid_ratios, top):
        N, Lq, M, L, P, _ = sampling_locations.shape
        loc, aw, vr = sampling_locations.contiguous(), attention_weights.contiguous(), valid_ratios.contiguous()
        keep = torch.empty(N, Lq, top, 2, dtype=torch.float32, device=loc.device)
        weights = torch.empty(N, Lq, top, dtype=torch.float32, device=loc.device)
        idx = torch.empty(N, Lq, top, dtype=torch.int64, device=loc.device)
        with _lib.on_device(loc.device):
            rc = _lib.lib().msda_decoder_select_samples_f32(loc.data_ptr(), aw.data_ptr(), vr.data_ptr(), N, Lq, M, L, P, int(top),
                                                            keep.data_ptr(), weights.data_ptr(), idx.data_ptr(), _stream(loc))
        _lib.check(rc, "msda_decoder_select_samples_f32")
        ctx.save_for_backward(idx, vr)
        ctx.dims = (N, Lq, M, L, P)
        ctx.mark_non_differentiable(weights, idx)
        return keep, weights, idx

    @staticmethod
    @once_differentiable
    def backward(ctx, g_keep, _gw, _gi):
        idx, vr = ctx.saved_tensors
        N, Lq, M, L, P = ctx.dims
        level = (idx // P) % L                                                     # (N, Lq, top)
        scale = torch.gather(vr[:, None].expand(N, Lq, L, 2), 2, level[..., None].expand(-1, -1, -1, 2))
        g = torch.zeros(N, Lq, M * L * P, 2, dtype=g_keep.dtype, device=g_keep.device)
        g.scatter_(2, idx[..., None].expand(-1, -1, -1, 2), g_keep / scale)
        return g.view(N, Lq, M, L, P, 2), None, None, None


def select_top_samples(sampling_locations, attention_weights, valid_ratios, top: int = TOP_SAMPLES):
    """``samples_keep`` of the reference decoder (:366-375): the sampling locations (divided by their level's valid ratio)
    of the ``top`` highest-weighted of a query's M*L*P points, in descending weight order.
    Returns (samples_keep (N, Lq, top, 2), top_weights (N, Lq, top), top_idx (N, Lq, top))."""
    N, Lq, M, L, P, _ = sampling_locations.shape
    K = M * L * P
    if _native_ok(sampling_locations, attention_weights, valid_ratios) and top <= 32 and top <= K <= 256:
        return _SelectTopSamples.apply(sampling_locations, attention_weights, valid_ratios, top)
    _strict.note_fallback("select_top_samples", "needs fp32 CUDA tensors outside autocast, top <= 32 and top <= M*L*P <= 256")
    loc = sampling_locations / valid_ratios[:, None, None, :, None, :]
    weights, idx = attention_weights.reshape(N, Lq, -1).topk(top, dim=2)
    keep = torch.gather(loc.reshape(N, Lq, -1, 2), 2, idx.unsqueeze(-1).repeat(1, 1, 1, 2))
    return keep, weights, idx


class DeformableTransformerDecoderLayer(nn.Module):
    def __init__(self, d_model=256, d_ffn=1024, dropout=0.1, activation="relu", n_levels=4, n_heads=8, n_points=4, fused=True):
        super().__init__()
        self.cross_attn = MSDeformAttn(d_model, n_levels, n_heads, n_points)
        self.cross_attn.fused = bool(fused)            # the decoder READS the locations and weights: emit_sampling stays True
        self.fused = bool(fused)
        self.activation_name = activation
        self.dropout1 = nn.Dropout(dropout)
        self.norm1 = nn.LayerNorm(d_model)
        self.self_attn = nn.MultiheadAttention(d_model, n_heads, dropout=dropout)
        self.dropout2 = nn.Dropout(dropout)
        self.norm2 = nn.LayerNorm(d_model)
        self.linear1 = nn.Linear(d_model, d_ffn)
        self.activation = _activation(activation)
        self.dropout3 = nn.Dropout(dropout)
        self.linear2 = nn.Linear(d_ffn, d_model)
        self.dropout4 = nn.Dropout(dropout)
        self.norm3 = nn.LayerNorm(d_model)

    @staticmethod
    def with_pos_embed(tensor, pos):
        return tensor if pos is None else tensor + pos

    def forward_ffn(self, tgt):
        tgt2 = self.linear2(self.dropout3(self.activation(self.linear1(tgt))))
        return self.norm3(tgt + self.dropout4(tgt2))

    def _epilogue_ok(self, tgt):
        ps = (self.dropout1.p, self.dropout2.p, self.dropout3.p, self.dropout4.p)
        return (self.fused and all(0.0 <= p < 1.0 for p in ps) and self.activation_name == "relu"
                and tgt.shape[-1] in epilogue.LN_CHANNELS and epilogue.supported(tgt, self.norm1.weight))

    def forward(self, tgt, query_pos, reference_points, src, src_spatial_shapes, level_start_index, src_padding_mask=None):
        q = k = self.with_pos_embed(tgt, query_pos)
        tgt2 = self.self_attn(q.transpose(0, 1), k.transpose(0, 1), tgt.transpose(0, 1))[0].transpose(0, 1)          # :326
        if not self._epilogue_ok(tgt):
            if self.fused:
                _strict.note_fallback("DeformableTransformerDecoderLayer epilogue", "needs ReLU, dropout in [0, 1), fp32 CUDA, d_model in %r" % (epilogue.LN_CHANNELS,))
            tgt = self.norm2(tgt + self.dropout2(tgt2))
            with torch.autocast(device_type=tgt.device.type, enabled=False):
                tgt2, sampling_locations, attention_weights = self.cross_attn(
                    self.with_pos_embed(tgt, query_pos), reference_points, src, src_spatial_shapes, level_start_index, src_padding_mask)
            tgt = self.norm1(tgt + self.dropout1(tgt2))
            return self.forward_ffn(tgt), sampling_locations, attention_weights
        p1, p2, p3, p4 = ((self.dropout1.p, self.dropout2.p, self.dropout3.p, self.dropout4.p) if self.training else (0.0,) * 4)
        rng = epilogue.new_rng(tgt.device) if max(p1, p2, p3, p4) > 0.0 else None
        tgt = epilogue.bias_residual_layer_norm(tgt2.contiguous(), None, tgt, self.norm2.weight, self.norm2.bias, self.norm2.eps,
                                                rng, 2, p2)                                                         # :327-328
        with torch.autocast(device_type=tgt.device.type, enabled=False):
            core, sampling_locations, attention_weights = self.cross_attn.attend(
                self.with_pos_embed(tgt, query_pos), reference_points, src, src_spatial_shapes, level_start_index, src_padding_mask)
        proj = self.cross_attn.output_proj
        tgt = epilogue.bias_residual_layer_norm(F.linear(core, proj.weight), proj.bias, tgt, self.norm1.weight, self.norm1.bias,
                                                self.norm1.eps, rng, 1, p1)                                         # :333-334
        hidden = epilogue.linear_relu(tgt, self.linear1.weight, self.linear1.bias, rng, 3, p3)                      # :318
        tgt = epilogue.bias_residual_layer_norm(F.linear(hidden, self.linear2.weight), self.linear2.bias, tgt, self.norm3.weight,
                                                self.norm3.bias, self.norm3.eps, rng, 4, p4)                        # :319-320
        return tgt, sampling_locations, attention_weights


def inverse_sigmoid(x, eps=1e-5):
    """util/misc.py:inverse_sigmoid of the reference."""
    x = x.clamp(min=0, max=1)
    return torch.log(x.clamp(min=eps) / (1 - x).clamp(min=eps))


class DeformableTransformerDecoder(nn.Module):
    def __init__(self, decoder_layer, num_layers, return_intermediate=False):
        super().__init__()
        self.layers = nn.ModuleList([copy.deepcopy(decoder_layer) for _ in range(num_layers)])
        self.num_layers = num_layers
        self.return_intermediate = return_intermediate
        self.bbox_embed = None          # set by the model for iterative box refinement (:350-352)
        self.class_embed = None

    def forward(self, tgt, reference_points, src, src_spatial_shapes, src_level_start_index, src_valid_ratios,
                query_pos=None, src_padding_mask=None):
        output = tgt
        intermediate, intermediate_reference_points, intermediate_samples = [], [], []
        samples_keep = None
        for lid, layer in enumerate(self.layers):
            reference_points_input = scale_reference_points(reference_points, src_valid_ratios)                      # :358-363
            output, sampling_locations, attention_weights = layer(output, query_pos, reference_points_input, src,
                                                                  src_spatial_shapes, src_level_start_index, src_padding_mask)
            samples_keep = select_top_samples(sampling_locations, attention_weights, src_valid_ratios, TOP_SAMPLES)[0]   # :366-375
            if self.bbox_embed is not None:                                                                          # :378-388
                tmp = self.bbox_embed[lid](output)
                if reference_points.shape[-1] == 4:
                    new_reference_points = (tmp + inverse_sigmoid(reference_points)).sigmoid()
                else:
                    new_reference_points = tmp
                    new_reference_points[..., :2] = tmp[..., :2] + inverse_sigmoid(reference_points)
                    new_reference_points = new_reference_points.sigmoid()
                reference_points = new_reference_points.detach()
            if self.return_intermediate:
                intermediate.append(output)
                intermediate_reference_points.append(reference_points)
                intermediate_samples.append(samples_keep)
        if self.return_intermediate:
            return torch.stack(intermediate), torch.stack(intermediate_reference_points), torch.stack(intermediate_samples)
        return output, reference_points, samples_keep


def build_decoder(num_layers=6, d_model=256, d_ffn=2048, dropout=0.0, n_levels=4, n_heads=8, n_points=4, fused=True,
                  return_intermediate=True):
    """The benchmark configuration of the decoder stack (opts.py: dec_layers, dim_feedforward 2048, 5 queries per frame)."""
    layer = DeformableTransformerDecoderLayer(d_model, d_ffn, dropout, "relu", n_levels, n_heads, n_points, fused=fused)
    return DeformableTransformerDecoder(layer, num_layers, return_intermediate)
