"""Generates tests/golden/*.npz from the REFERENCE's own Python implementation.

Run in the build container only (it reads /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

It imports ``ms_deform_attn_core_pytorch`` from
/root/reference/models/ops/functions/ms_deform_attn_func.py (:41-61) -- with an empty stub for the
compiled ``MultiScaleDeformableAttention`` module that file imports at :18 -- runs it in fp64
(forward + autograd backward) on fp32-valued inputs and stores inputs and outputs.  The committed
.npz files are what pins oracle/ (tests/test_oracle.py) and the CUDA path (tests/test_parity_gpu.py).

Cases:
  ref_test      the reference's own test geometry and input recipe (models/ops/test.py:21-37, seed 3)
  oob_ragged    ragged levels incl. a 1x2 one, locations in U(-0.4, 1.4): out-of-range points and
                partially out-of-bounds corners (cuh:56-78, :288), which the reference never tests
  d32_l4p4      the production head layout M=8, D=32, L=4, P=4 on a small pyramid, U(-0.1, 1.1)
  odd_dims      D=5, M=3, L=2 with degenerate 1xW / Hx1 levels, P=3
  single_px     L=1, a 1x1 level, P=1
"""
import os
import sys
import types

import numpy as np
import torch

REF_OPS = "/root/reference/models/ops"
HERE = os.path.dirname(os.path.abspath(__file__))


def load_reference():
    sys.modules.setdefault("MultiScaleDeformableAttention", types.ModuleType("MultiScaleDeformableAttention"))
    sys.path.insert(0, REF_OPS)
    from functions.ms_deform_attn_func import ms_deform_attn_core_pytorch  # noqa
    return ms_deform_attn_core_pytorch


def case_inputs(name):
    if name == "ref_test":
        torch.manual_seed(3)
        N, M, D, Lq, L, P = 1, 2, 2, 2, 2, 2
        shapes = [(6, 4), (3, 2)]
        S = sum(h * w for h, w in shapes)
        value = torch.rand(N, S, M, D) * 0.01
        loc = torch.rand(N, Lq, M, L, P, 2)
        attn = torch.rand(N, Lq, M, L, P) + 1e-5
        attn /= attn.sum(-1, keepdim=True).sum(-2, keepdim=True)
        return value, shapes, loc, attn
    spec = {
        "oob_ragged": (2, 2, 3, 7, [(6, 4), (3, 5), (1, 2)], 4, (-0.4, 1.4), 11),
        "d32_l4p4": (2, 8, 32, 33, [(7, 9), (4, 5), (2, 3), (1, 2)], 4, (-0.1, 1.1), 12),
        "odd_dims": (3, 3, 5, 4, [(1, 7), (4, 1)], 3, (-0.2, 1.2), 13),
        "single_px": (2, 1, 4, 3, [(1, 1)], 1, (-0.5, 1.5), 14),
    }[name]
    N, M, D, Lq, shapes, P, (lo, hi), seed = spec
    g = torch.Generator().manual_seed(seed)
    L = len(shapes)
    S = sum(h * w for h, w in shapes)
    value = torch.randn(N, S, M, D, generator=g)
    loc = torch.rand(N, Lq, M, L, P, 2, generator=g) * (hi - lo) + lo
    attn = torch.softmax(torch.randn(N, Lq, M, L * P, generator=g), -1).view(N, Lq, M, L, P)
    return value, shapes, loc, attn


def main():
    ref = load_reference()
    for name in ("ref_test", "oob_ragged", "d32_l4p4", "odd_dims", "single_px"):
        value, shapes, loc, attn = case_inputs(name)
        g = torch.Generator().manual_seed(1000 + len(name))
        shapes_t = torch.tensor(shapes, dtype=torch.long)
        start = torch.cat((shapes_t.new_zeros(1), shapes_t.prod(1).cumsum(0)[:-1]))
        v = value.double().requires_grad_(True)
        s = loc.double().requires_grad_(True)
        a = attn.double().requires_grad_(True)
        out = ref(v, shapes_t, s, a)
        grad_out = torch.randn(out.shape, generator=g)
        out.backward(grad_out.double())
        np.savez(os.path.join(HERE, f"{name}.npz"),
                 value=value.numpy(), shapes=shapes_t.numpy(), start=start.numpy(), loc=loc.numpy(),
                 attn=attn.numpy(), grad_out=grad_out.numpy(),
                 out=out.detach().numpy(), grad_value=v.grad.numpy(), grad_loc=s.grad.numpy(),
                 grad_attn=a.grad.numpy())
        print(name, "out", tuple(out.shape), "bytes", os.path.getsize(os.path.join(HERE, f"{name}.npz")))


if __name__ == "__main__":
    main()
