"""Strict mode for the helpers around the operator (epilogue / flatten / decoder consumers).

Those helpers cover the layouts the production model uses with CUDA kernels and take the equivalent torch
formulation for anything else (CPU tensors in the CPU-side tests, other dtypes, autocast, non-ReLU activations).
A benchmark must not take that route unnoticed: with strict mode on -- ``ocpg_b200.set_strict(True)`` or
``OCPG_B200_STRICT=1`` -- every such fallback raises instead, saying which helper and why.  ``fallback_counts()``
reports how often each helper fell back in non-strict mode.  (The operator itself, MSDeformAttnFunction /
MSDeformAttn, never falls back: it raises without the CUDA library.)
"""
from __future__ import annotations

import collections
import os

_STRICT = os.environ.get("OCPG_B200_STRICT", "0") not in ("", "0", "false", "False")
_COUNTS: "collections.Counter[str]" = collections.Counter()


def set_strict(on: bool = True) -> None:
    global _STRICT
    _STRICT = bool(on)


def is_strict() -> bool:
    return _STRICT


def fallback_counts() -> dict:
    return dict(_COUNTS)


def note_fallback(helper: str, reason: str) -> None:
    """Called by a helper right before it runs its torch formulation."""
    _COUNTS[helper] += 1
    if _STRICT:
        raise RuntimeError(f"ocpg_b200 strict mode: {helper} would run its torch formulation instead of the CUDA kernel ({reason})")
