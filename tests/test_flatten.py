"""Caller-side flattening (ocpg_b200/flatten.py) against the reference's own statements
(models/deformable_transformer.py:149-169, :205-212) -- pure data movement plus one add, so the bar is bit-exact."""
import pytest
import torch


def reference_flatten(srcs, poss, level_embed):
    """deformable_transformer.py:149-169, verbatim in spirit."""
    src_flatten, lvl_pos = [], []
    for lvl, (src, pos) in enumerate(zip(srcs, poss)):
        src_flatten.append(src.flatten(2).transpose(1, 2))
        lvl_pos.append(pos.flatten(2).transpose(1, 2) + level_embed[lvl].view(1, 1, -1))
    return torch.cat(src_flatten, 1), torch.cat(lvl_pos, 1)


def reference_unflatten(memory, shapes):
    """deformable_transformer.py:205-212."""
    out, at = [], 0
    bs, _, c = memory.shape
    for h, w in shapes:
        out.append(memory[:, at:at + h * w, :].reshape(bs, h, w, c).permute(0, 3, 1, 2).contiguous())
        at += h * w
    return out


def make(levels, N, C, device, seed=0):
    g = torch.Generator(device=device).manual_seed(seed)
    srcs = [torch.randn(N, C, h, w, device=device, generator=g) for h, w in levels]
    poss = [torch.randn(N, C, h, w, device=device, generator=g) for h, w in levels]
    return srcs, poss, torch.randn(len(levels), C, device=device, generator=g)


def test_cpu_tensors_take_the_reference_formulation():
    from ocpg_b200.flatten import flatten_levels, unflatten_levels
    levels = [(5, 7), (3, 4), (1, 2)]
    srcs, poss, le = make(levels, 2, 12, "cpu")
    s, p, shapes, start = flatten_levels(srcs, poss, le)
    rs, rp = reference_flatten(srcs, poss, le)
    assert torch.equal(s, rs) and torch.equal(p, rp)
    assert shapes.tolist() == [list(l) for l in levels] and start.tolist() == [0, 35, 47]
    for a, b in zip(unflatten_levels(s, levels[:-1]), reference_unflatten(rs, levels[:-1])):
        assert torch.equal(a, b)
    s2, p2, _, _ = flatten_levels(srcs)
    assert torch.equal(s2, rs) and p2 is None


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import ocpg_b200
    ocpg_b200.lib()
    return torch.device("cuda:0")


@pytest.mark.gpu
@pytest.mark.parametrize("levels,N,C", [
    ([(45, 80), (23, 40), (12, 20), (6, 10)], 5, 256),          # configs[1]
    ([(5, 7), (3, 4), (1, 2)], 2, 12),                          # ragged tiles, C not a multiple of 32
    ([(1, 1)], 1, 1),
    ([(33, 31), (2, 65)], 3, 40),
    ([(9, 9)] * 8, 2, 64),                                      # the maximum level count
    ([(4, 5), (2, 2)], 2, 12),                                  # 16-byte kernels, partial tiles both ways
    ([(10, 18), (8, 9), (4, 17)], 1, 72),                       # 16-byte kernels, C = 64 + 8
    ([(80, 144), (40, 72), (20, 36), (10, 18)], 2, 256),        # configs[2] geometry
])
def test_flatten_unflatten_bit_exact_with_grads(dev, levels, N, C):
    import ocpg_b200
    from ocpg_b200.flatten import flatten_levels, unflatten_levels
    srcs, poss, le = make(levels, N, C, dev, seed=len(levels) + C)
    leaves = [t.clone().requires_grad_(True) for t in srcs + poss + [le]]
    L = len(levels)
    n0 = ocpg_b200.launch_count()
    s, p, shapes, start = flatten_levels(leaves[:L], leaves[L:2 * L], leaves[2 * L])
    assert ocpg_b200.launch_count() - n0 == 1
    ref_leaves = [t.clone().requires_grad_(True) for t in srcs + poss + [le]]
    rs, rp = reference_flatten(ref_leaves[:L], ref_leaves[L:2 * L], ref_leaves[2 * L])
    assert torch.equal(s, rs) and torch.equal(p, rp)
    gs, gp = torch.randn_like(rs), torch.randn_like(rp)
    torch.autograd.backward([s, p], [gs, gp])
    torch.autograd.backward([rs, rp], [gs, gp])
    for a, b in zip(leaves[:2 * L], ref_leaves[:2 * L]):
        assert torch.equal(a.grad, b.grad)
    assert torch.allclose(leaves[-1].grad, ref_leaves[-1].grad, rtol=1e-4, atol=1e-4)       # a sum: order differs
    # back to maps: all but the last level, as the reference does; gradient = flatten with zeros for the dropped level
    keep = levels[:-1] if L > 1 else levels
    mem = torch.randn_like(rs).requires_grad_(True)
    mem_ref = mem.detach().clone().requires_grad_(True)
    n0 = ocpg_b200.launch_count()
    maps = unflatten_levels(mem, keep)
    assert ocpg_b200.launch_count() - n0 == 1
    ref_maps = reference_unflatten(mem_ref, keep)
    gm = [torch.randn_like(m) for m in ref_maps]
    for a, b in zip(maps, ref_maps):
        assert a.is_contiguous() and torch.equal(a, b)
    torch.autograd.backward(maps, gm)
    torch.autograd.backward(ref_maps, gm)
    assert torch.equal(mem.grad, mem_ref.grad)
    # src only
    s2, p2, _, _ = flatten_levels(srcs)
    assert p2 is None and torch.equal(s2, rs.detach())


@pytest.mark.gpu
def test_flatten_argument_errors(dev):
    import ctypes
    import ocpg_b200
    L = ocpg_b200.lib()
    one = (ctypes.c_int * 1)(4)
    assert L.msda_flatten_levels_f32(0, None, None, None, one, one, 1, 8, None, None, None) == -2
    assert L.msda_flatten_levels_f32(9, None, None, None, one, one, 1, 8, None, None, None) == -2
    assert L.msda_flatten_levels_f32(1, None, None, None, one, one, 1, 8, None, None, None) == -1
    bad = (ctypes.c_int * 1)(0)
    assert L.msda_unflatten_levels_f32(1, None, bad, one, 1, 8, 0, None, None) == -1


@pytest.mark.gpu
def test_flatten_random_shapes(dev):
    """40 random pyramids (1-8 levels, odd and even sizes, channel counts that are and are not multiples of 4 / 32): both
    kernel families (4-byte and 16-byte) against the reference statements, forward and inverse, bit for bit."""
    import random
    from ocpg_b200.flatten import flatten_levels, unflatten_levels
    rnd = random.Random(7)
    for case in range(40):
        L = rnd.randint(1, 8)
        even = case % 2 == 0
        levels = [(rnd.randint(1, 20) * (2 if even else 1), rnd.randint(1, 20) * (2 if even else 1)) for _ in range(L)]
        C = rnd.choice([4, 12, 32, 64, 100] if even else [1, 3, 12, 33, 64])
        N = rnd.randint(1, 3)
        srcs, poss, le = make(levels, N, C, dev, seed=case)
        s, p, shapes, start = flatten_levels(srcs, poss, le)
        rs, rp = reference_flatten(srcs, poss, le)
        assert torch.equal(s, rs) and torch.equal(p, rp), (case, levels, C)
        keep = levels[:max(1, L - 1)]
        for a, b in zip(unflatten_levels(rs, keep), reference_unflatten(rs, keep)):
            assert torch.equal(a, b), (case, levels, C)
