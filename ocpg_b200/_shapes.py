"""Host-side copies of ``spatial_shapes`` tensors (shared by the module's shape check and the encoder's reference points)."""
from __future__ import annotations

import torch

_HOST_SHAPES = {}      # id(tensor) -> (tensor, version, rows as Python ints); holds the tensor so its id / storage cannot be reused


def remember_host_shapes(spatial_shapes: torch.Tensor, shapes) -> torch.Tensor:
    """Attach the host-side copy of ``spatial_shapes`` to the tensor (``flatten_levels`` builds the tensor FROM host
    integers, so no device->host copy is ever needed for it)."""
    spatial_shapes._ocpg_host_shapes = (spatial_shapes._version, [tuple(int(v) for v in hw) for hw in shapes])
    return spatial_shapes


def shapes_on_host(spatial_shapes):
    """``spatial_shapes`` as a list of (H, W) ints.  The reference iterates the CUDA tensor (one device->host copy per
    call, deformable_transformer.py:269); here the copy happens at most once per tensor OBJECT and version -- which also
    keeps the forward free of synchronisation, so that a whole training step can be captured into a CUDA graph.  The cache
    keeps a reference to the tensor: a freed tensor's address can be handed to a new one with other contents, so identity
    without ownership would not be a safe key."""
    if not isinstance(spatial_shapes, torch.Tensor):
        return [tuple(int(v) for v in hw) for hw in spatial_shapes]
    tagged = getattr(spatial_shapes, "_ocpg_host_shapes", None)
    if tagged is not None and tagged[0] == spatial_shapes._version:
        return tagged[1]
    hit = _HOST_SHAPES.get(id(spatial_shapes))
    if hit is None or hit[0] is not spatial_shapes or hit[1] != spatial_shapes._version:
        if len(_HOST_SHAPES) > 32:
            _HOST_SHAPES.clear()
        hit = _HOST_SHAPES[id(spatial_shapes)] = (spatial_shapes, spatial_shapes._version,
                                                  [tuple(hw) for hw in spatial_shapes.tolist()])
    return hit[2]
