"""Synthetic ReferFormer/OCPG-shaped inputs for the MSDeformAttn hot path.

The shapes come from the reference's own configuration:
hidden_dim=256, nheads=8, num_feature_levels=4, enc/dec_n_points=4 (opts.py:46,56,60,66-67); the four
feature levels are strides 8/16/32/64 of the padded input, each halving rounded up
(models/ocpg.py:98-122; models/backbone.py:69-70); encoder queries are the pixels themselves
(Lq = S, deformable_transformer.py:283-289); decoder cross-attention uses num_queries=5
(opts.py:64).

Sampling-location regimes:
  * ``"init"``  (R1): encoder reference points = pixel centres of the query's own level
    (deformable_transformer.py:269-281) plus offsets ~ N(0, (sigma px)^2) in each level's own pixel
    units (the module divides offsets by (W_l, H_l), ms_deform_attn.py:104-107).  Spatially
    coherent -- the realistic cache behaviour, and the headline regime.
  * ``"uniform"`` (R2): i.i.d. U(-0.1, 1.1): a random gather with ~17 % of the coordinates out of
    range -- the stress regime.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Tuple

import torch

N_HEADS, HEAD_DIM, N_LEVELS, N_POINTS = 8, 32, 4, 4


def level_geometry(height: int, width: int, n_levels: int = N_LEVELS, first_stride: int = 8):
    """[(H_l, W_l)] for an input of ``height x width`` pixels: stride 8, then ceil-halving."""
    h, w = math.ceil(height / first_stride), math.ceil(width / first_stride)
    out = []
    for _ in range(n_levels):
        out.append((h, w))
        h, w = math.ceil(h / 2), math.ceil(w / 2)
    return out


@dataclass(frozen=True)
class Workload:
    name: str
    n_frames: int                 # N = b*t
    levels: Tuple[Tuple[int, int], ...]
    n_queries: int                # Lq ( == S for the encoder)
    n_heads: int = N_HEADS
    head_dim: int = HEAD_DIM
    n_points: int = N_POINTS

    @property
    def S(self) -> int:
        return sum(h * w for h, w in self.levels)

    @property
    def L(self) -> int:
        return len(self.levels)

    @property
    def queries(self) -> int:
        """Queries per call: one query = one (n, q) pair, all heads."""
        return self.n_frames * self.n_queries

    def algorithmic_bytes(self, value_bytes: int = 4, out_bytes: int = 4):
        """(fwd, bwd) compulsory bytes, each tensor once:
        fwd = V + N*Lq*(B_loc + B_aw + C*so);  bwd = 2V + N*Lq*(C*so + 2*B_loc + 2*B_aw)."""
        C = self.n_heads * self.head_dim
        V = self.n_frames * self.S * C * value_bytes
        b_loc = self.n_heads * self.L * self.n_points * 2 * 4
        b_aw = self.n_heads * self.L * self.n_points * 4
        q = self.queries
        # decoder-like shapes touch only part of value on the way in (config 4)
        touched = q * self.n_heads * self.L * self.n_points * 4 * self.head_dim * value_bytes
        v_read = min(V, touched)
        fwd = v_read + q * (b_loc + b_aw + C * out_bytes)
        bwd = v_read + V + q * (C * out_bytes + 2 * b_loc + 2 * b_aw)
        return fwd, bwd


def encoder_workload(name: str, n_frames: int, height: int, width: int) -> Workload:
    lv = tuple(level_geometry(height, width))
    return Workload(name, n_frames, lv, sum(h * w for h, w in lv))


# BASELINE.json configs
A2D_ENCODER = encoder_workload("a2d_r101_encoder_N5_360x640", 5, 360, 640)          # configs[0], [1]
YTVOS_ENCODER = encoder_workload("ytvos_swinb_encoder_N10_640x1152", 10, 640, 1152)  # configs[2]
A2D_DECODER = Workload("a2d_decoder_cross_N5_Lq5", 5, A2D_ENCODER.levels, 5)        # configs[3]


def encoder_reference_points(levels, device, dtype=torch.float32):
    """(S, L, 2) reference points of the encoder with valid_ratios == 1: the pixel centres of each
    query's own level, normalised, replicated across levels (deformable_transformer.py:269-281)."""
    refs = []
    for h, w in levels:
        ys = (torch.arange(h, device=device, dtype=dtype) + 0.5) / h
        xs = (torch.arange(w, device=device, dtype=dtype) + 0.5) / w
        gy, gx = torch.meshgrid(ys, xs, indexing="ij")
        refs.append(torch.stack((gx.reshape(-1), gy.reshape(-1)), -1))
    ref = torch.cat(refs, 0)                       # (S, 2)
    return ref[:, None, :].expand(-1, len(levels), -1)


def make_inputs(wl: Workload, regime: str = "init", seed: int = 0, device="cpu", dtype=torch.float32,
                sigma_px: float = 2.0, value_dtype=None):
    """Returns dict(value, shapes, start, loc, attn, grad_out) for ``wl``.

    value ~ N(0,1); attn = softmax(N(0,1)) over L*P; grad_out ~ N(0,1); loc per ``regime``.
    ``value_dtype`` (e.g. torch.bfloat16) applies to value and grad_out only.
    """
    g = torch.Generator(device=device).manual_seed(seed)
    N, S, M, D, L, P, Lq = wl.n_frames, wl.S, wl.n_heads, wl.head_dim, wl.L, wl.n_points, wl.n_queries
    kw = dict(device=device, dtype=dtype, generator=g)
    value = torch.randn(N, S, M, D, **kw)
    shapes = torch.tensor(wl.levels, dtype=torch.int64, device=device)
    start = torch.cat((shapes.new_zeros(1), (shapes[:, 0] * shapes[:, 1]).cumsum(0)[:-1]))
    if regime == "uniform":
        loc = torch.rand(N, Lq, M, L, P, 2, **kw) * 1.2 - 0.1
    elif regime == "init":
        if Lq == S:
            ref = encoder_reference_points(wl.levels, device, dtype)        # (S, L, 2)
        else:  # decoder-like: a handful of object queries at random positions
            ref = torch.rand(Lq, 1, 2, **kw).expand(-1, L, -1)
        wh = torch.stack((shapes[:, 1], shapes[:, 0]), -1).to(dtype)        # (L, 2) = (W, H)
        off = torch.randn(N, Lq, M, L, P, 2, **kw) * sigma_px
        loc = ref[None, :, None, :, None, :] + off / wh[None, None, None, :, None, :]
    elif regime == "module_init":
        # exactly what a freshly initialised MSDeformAttn produces (ms_deform_attn.py:62-74, :104-107): zero offset weights,
        # offset bias = head m's compass direction (max-norm 1) times the point index 1..P, in pixels of every level;
        # neighbouring queries therefore sample a regular lattice (spacing 1, 1/2, 1/4, 1/8 px at levels 0..3 for
        # level-0 queries) -- far more row sharing between neighbours than the i.i.d. "init" regime
        if Lq == S:
            ref = encoder_reference_points(wl.levels, device, dtype)
        else:
            ref = torch.rand(Lq, 1, 2, **kw).expand(-1, L, -1)
        ang = torch.arange(M, device=device, dtype=dtype) * (2.0 * math.pi / M)
        direction = torch.stack([ang.cos(), ang.sin()], -1)
        direction = direction / direction.abs().max(-1, keepdim=True)[0]                         # (M, 2)
        off = direction.view(M, 1, 1, 2) * torch.arange(1, P + 1, device=device, dtype=dtype).view(1, 1, P, 1)
        wh = torch.stack((shapes[:, 1], shapes[:, 0]), -1).to(dtype)
        loc = (ref[None, :, None, :, None, :] + off.expand(M, L, P, 2)[None, None] / wh[None, None, None, :, None, :]).expand(N, Lq, M, L, P, 2)
    else:
        raise ValueError(f"unknown regime {regime!r}")
    attn = torch.softmax(torch.randn(N, Lq, M, L * P, **kw), -1).view(N, Lq, M, L, P)
    grad_out = torch.randn(N, Lq, M * D, **kw)
    if value_dtype is not None:
        value, grad_out = value.to(value_dtype), grad_out.to(value_dtype)
    return dict(value=value.contiguous(), shapes=shapes, start=start, loc=loc.contiguous(),
                attn=attn.contiguous(), grad_out=grad_out.contiguous())


def shard_frames(n_frames: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous block split of the flattened N = b*t axis: returns
    (first_frame, n_local).  Frames are independent units -- no exchange step on the op."""
    base, rem = divmod(n_frames, world_size)
    n_local = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, n_local


def all_workloads() -> List[Workload]:
    return [A2D_ENCODER, YTVOS_ENCODER, A2D_DECODER]
