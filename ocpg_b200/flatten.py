"""Caller-side flattening around the encoder on the C ABI of include/msda_sm100.h.

Reference: DeformableTransformer.forward, models/deformable_transformer.py

    :149-169   per level  src.flatten(2).transpose(1, 2), pos_embed.flatten(2).transpose(1, 2) + level_embed[lvl],
               then torch.cat over the levels -> src_flatten, lvl_pos_embed_flatten  (N, S, C); spatial_shapes,
               level_start_index
    :205-212   memory[:, start:start + h*w].reshape(N, h, w, C).permute(0, 3, 1, 2).contiguous() for all but the last level

``flatten_levels`` and ``unflatten_levels`` do each side in one launch (tiled transpositions through shared memory) and
are each other's backward.  fp32 CUDA tensors; anything else takes the reference's torch formulation.
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence, Tuple

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _lib, _strict

MAX_LEVELS = 8


def _stream(t) -> int:
    return _lib.raw_stream(t.device)


def _native_ok(tensors: Sequence[torch.Tensor]) -> bool:
    return (0 < len(tensors) <= MAX_LEVELS and not torch.is_autocast_enabled()
            and all(t is not None and t.is_cuda and t.dtype == torch.float32 for t in tensors))


def _aligned(t: torch.Tensor) -> torch.Tensor:
    """contiguous and 16-byte aligned (a view at an odd storage offset is copied)."""
    t = t.contiguous()
    return t if t.data_ptr() % 16 == 0 else t.clone()


def _ptr_array(tensors):
    return (ctypes.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


def _int_array(values):
    return (ctypes.c_int * len(values))(*[int(v) for v in values])


def _flatten_native(srcs, poss, level_embed):
    N, C = srcs[0].shape[:2]
    hs, ws = [t.shape[2] for t in srcs], [t.shape[3] for t in srcs]
    S = sum(h * w for h, w in zip(hs, ws))
    srcs = [_aligned(t) for t in srcs]
    src_flat = torch.empty(N, S, C, dtype=torch.float32, device=srcs[0].device)
    pos_flat = None
    if poss is not None:
        poss = [_aligned(t) for t in poss]
        pos_flat = torch.empty_like(src_flat)
    with _lib.on_device(src_flat.device):
        rc = _lib.lib().msda_flatten_levels_f32(
            len(srcs), _ptr_array(srcs), None if poss is None else _ptr_array(poss),
            None if level_embed is None else level_embed.data_ptr(), _int_array(hs), _int_array(ws), N, C,
            src_flat.data_ptr(), None if pos_flat is None else pos_flat.data_ptr(), _stream(src_flat))
    _lib.check(rc, "msda_flatten_levels_f32")
    return src_flat, pos_flat


def _unflatten_native(flat, shapes):
    N, S, C = flat.shape
    flat = _aligned(flat)
    maps = [torch.empty(N, C, h, w, dtype=torch.float32, device=flat.device) for h, w in shapes]
    with _lib.on_device(flat.device):
        rc = _lib.lib().msda_unflatten_levels_f32(len(shapes), flat.data_ptr(), _int_array([h for h, _ in shapes]),
                                                  _int_array([w for _, w in shapes]), N, C, S, _ptr_array(maps), _stream(flat))
    _lib.check(rc, "msda_unflatten_levels_f32")
    return maps


class _FlattenLevels(Function):
    """args: L, has_pos, level_embed-or-None, *srcs, *poss."""

    @staticmethod
    def forward(ctx, L, has_pos, level_embed, *maps):
        srcs, poss = list(maps[:L]), (list(maps[L:]) if has_pos else None)
        le = None if level_embed is None else _aligned(level_embed)
        src_flat, pos_flat = _flatten_native(srcs, poss, le)
        ctx.L, ctx.has_pos, ctx.has_le = L, has_pos, level_embed is not None
        ctx.shapes = [tuple(t.shape[2:]) for t in srcs]
        if has_pos:
            return src_flat, pos_flat
        return src_flat

    @staticmethod
    @once_differentiable
    def backward(ctx, g_src, g_pos=None):
        grads = _unflatten_native(g_src, ctx.shapes)
        g_le = None
        if ctx.has_pos:
            grads += _unflatten_native(g_pos, ctx.shapes)
            if ctx.has_le:
                starts = [0]
                for h, w in ctx.shapes:
                    starts.append(starts[-1] + h * w)
                g_le = torch.stack([g_pos[:, a:b].sum((0, 1)) for a, b in zip(starts[:-1], starts[1:])])
        return (None, None, g_le, *grads)


def flatten_levels(srcs: List[torch.Tensor], pos_embeds: Optional[List[torch.Tensor]] = None,
                   level_embed: Optional[torch.Tensor] = None):
    """The encoder's inputs from the L feature maps (:149-169).

    srcs, pos_embeds: lists of (N, C, H_l, W_l); level_embed (L, C).  Returns (src_flatten (N, S, C),
    lvl_pos_embed_flatten (N, S, C) or None, spatial_shapes (L, 2) int64 on the maps' device, level_start_index (L,))."""
    shapes = [tuple(t.shape[2:]) for t in srcs]
    spatial_shapes = torch.as_tensor(shapes, dtype=torch.long, device=srcs[0].device)
    from ._shapes import remember_host_shapes                 # the encoder needs (H, W) on the host: no copy back later
    remember_host_shapes(spatial_shapes, shapes)
    level_start_index = torch.cat((spatial_shapes.new_zeros((1,)), spatial_shapes.prod(1).cumsum(0)[:-1]))
    tensors = list(srcs) + (list(pos_embeds) if pos_embeds is not None else []) + ([level_embed] if level_embed is not None else [])
    if _native_ok(list(srcs)) and all(t.is_cuda and t.dtype == torch.float32 for t in tensors):
        out = _FlattenLevels.apply(len(srcs), pos_embeds is not None, level_embed if pos_embeds is not None else None,
                                   *srcs, *(pos_embeds or []))
        src_flat, pos_flat = out if pos_embeds is not None else (out, None)
        return src_flat, pos_flat, spatial_shapes, level_start_index
    _strict.note_fallback("flatten_levels", "needs fp32 CUDA maps outside autocast")
    src_flat = torch.cat([s.flatten(2).transpose(1, 2) for s in srcs], 1)
    pos_flat = None
    if pos_embeds is not None:
        pos_flat = torch.cat([p.flatten(2).transpose(1, 2) + (0 if level_embed is None else level_embed[l].view(1, 1, -1))
                              for l, p in enumerate(pos_embeds)], 1)
    return src_flat, pos_flat, spatial_shapes, level_start_index


class _UnflattenLevels(Function):
    @staticmethod
    def forward(ctx, flat, shapes):
        ctx.shapes, ctx.flat_shape = shapes, tuple(flat.shape)
        return tuple(_unflatten_native(flat, shapes))

    @staticmethod
    @once_differentiable
    def backward(ctx, *g_maps):
        N, S, C = ctx.flat_shape
        covered = sum(h * w for h, w in ctx.shapes)
        maps = [g if g is not None else torch.zeros(N, C, h, w, device=g_maps[0].device) for g, (h, w) in zip(g_maps, ctx.shapes)]
        g_flat, _ = _flatten_native(maps, None, None)
        if covered < S:                                                    # the levels that were not converted (:207)
            g_flat = torch.cat((g_flat, g_flat.new_zeros(N, S - covered, C)), 1)
        return g_flat, None


def unflatten_levels(memory: torch.Tensor, shapes: Sequence[Tuple[int, int]]) -> List[torch.Tensor]:
    """memory (N, S, C) -> [(N, C, H_l, W_l)] for the given leading levels, contiguous (:205-212; the reference converts
    ``num_feature_level - 1`` of them)."""
    shapes = [(int(h), int(w)) for h, w in shapes]
    if _native_ok([memory]) and 0 < len(shapes) <= MAX_LEVELS:
        return list(_UnflattenLevels.apply(memory, tuple(shapes)))
    _strict.note_fallback("unflatten_levels", "needs an fp32 CUDA tensor outside autocast and 1..%d levels" % MAX_LEVELS)
    out, at = [], 0
    N, _, C = memory.shape
    for h, w in shapes:
        out.append(memory[:, at:at + h * w, :].reshape(N, h, w, C).permute(0, 3, 1, 2).contiguous())
        at += h * w
    return out
