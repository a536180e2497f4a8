"""The reference's CPU path, restated (TEST INFRASTRUCTURE -- see oracle/__init__.py).

The reference's only CPU-runnable implementation of multi-scale deformable attention is
``ms_deform_attn_core_pytorch`` (/root/reference/models/ops/functions/ms_deform_attn_func.py:41-61):
per level, view that level's rows of ``value`` as an image batch ``(N*M, D, H_l, W_l)``, sample it
with ``F.grid_sample(bilinear, zeros padding, align_corners=False)`` at ``2*loc - 1`` (:47, :55-56),
then weight by the attention weights and sum over levels and points (:59-60).

This is the same algorithm written independently (one einsum for the weighted sum instead of the
stack/flatten/sum chain), pinned against the imported reference function by tests/golden
(tests/test_oracle.py::test_grid_sample_port_matches_golden).  It is what bench.py times as the
``cpu_baseline`` / ``--impl reference`` arm, with all host threads torch can use.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def msda_grid_sample(value, spatial_shapes, sampling_locations, attention_weights):
    """value (N,S,M,D), spatial_shapes [(H,W)...], loc (N,Lq,M,L,P,2) (x,y), attn (N,Lq,M,L,P)
    -> (N, Lq, M*D), any float dtype, CPU or CUDA."""
    N, S, M, D = value.shape
    Lq, L, P = sampling_locations.shape[1], sampling_locations.shape[3], sampling_locations.shape[4]
    hw = [(int(h), int(w)) for h, w in (spatial_shapes.tolist() if hasattr(spatial_shapes, "tolist")
                                         else spatial_shapes)]
    assert len(hw) == L and sum(h * w for h, w in hw) == S
    # image batch index is (n, m); channels are D
    images = value.permute(0, 2, 3, 1).reshape(N * M, D, S)           # (N*M, D, S)
    grid = (2.0 * sampling_locations - 1.0).permute(0, 2, 1, 3, 4, 5)  # (N, M, Lq, L, P, 2)
    grid = grid.reshape(N * M, Lq, L, P, 2)
    sampled = []
    lo = 0
    for lvl, (h, w) in enumerate(hw):
        img = images[:, :, lo:lo + h * w].reshape(N * M, D, h, w)
        lo += h * w
        sampled.append(F.grid_sample(img, grid[:, :, lvl], mode="bilinear", padding_mode="zeros",
                                     align_corners=False))           # (N*M, D, Lq, P)
    sampled = torch.stack(sampled, dim=3)                            # (N*M, D, Lq, L, P)
    a = attention_weights.permute(0, 2, 1, 3, 4).reshape(N * M, Lq, L, P)
    out = torch.einsum("bdqlp,bqlp->bqd", sampled, a)                # (N*M, Lq, D)
    return out.reshape(N, M, Lq, D).permute(0, 2, 1, 3).reshape(N, Lq, M * D).contiguous()


def msda_grid_sample_fwd_bwd(value, spatial_shapes, sampling_locations, attention_weights, grad_output):
    """Forward + autograd backward: returns (out, grad_value, grad_loc, grad_attn)."""
    v = value.detach().clone().requires_grad_(True)
    s = sampling_locations.detach().clone().requires_grad_(True)
    a = attention_weights.detach().clone().requires_grad_(True)
    out = msda_grid_sample(v, spatial_shapes, s, a)
    out.backward(grad_output)
    return out.detach(), v.grad, s.grad, a.grad
