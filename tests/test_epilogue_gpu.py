"""GPU parity of the encoder-layer epilogue kernels (ocpg_b200/epilogue.py) against the
torch operators of the reference's graph (models/deformable_transformer.py:243-260) evaluated in fp64."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import ocpg_b200
    ocpg_b200.lib()
    return torch.device("cuda:0")


def rel(a, b):
    a, b = a.detach().double(), b.detach().double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("C", [128, 256, 512, 1024])
@pytest.mark.parametrize("rows,with_bias", [(1, True), (37, False), (4099, True)])
def test_bias_residual_layer_norm(dev, C, rows, with_bias):
    from ocpg_b200 import epilogue
    g = torch.Generator(device=dev).manual_seed(C + rows)
    mk = lambda *s: torch.randn(*s, device=dev, generator=g)
    x, res = mk(rows, C) * 2, mk(rows, C) + 0.5
    bias = mk(C) if with_bias else None
    gamma, beta = 1 + 0.1 * mk(C), 0.1 * mk(C)
    dy = mk(rows, C)
    leaves = [t.clone().requires_grad_(True) for t in (x, res, gamma, beta)] + ([bias.clone().requires_grad_(True)] if with_bias else [])
    y = epilogue.bias_residual_layer_norm(leaves[0], leaves[4] if with_bias else None, leaves[1], leaves[2], leaves[3], 1e-5)
    assert y.grad_fn is not None and type(y.grad_fn).__name__.startswith("_BiasResidualLayerNorm")
    y.backward(dy)
    ref = [t.double().clone().requires_grad_(True) for t in (x, res, gamma, beta)] + ([bias.double().clone().requires_grad_(True)] if with_bias else [])
    z = ref[1] + (ref[0] + ref[4] if with_bias else ref[0])
    yr = F.layer_norm(z, (C,), ref[2], ref[3], 1e-5)
    yr.backward(dy.double())
    assert rel(y, yr) <= 2e-6
    for a, b in zip(leaves, ref):
        assert rel(a.grad, b.grad) <= 2e-5, (a.shape, rel(a.grad, b.grad))


@pytest.mark.parametrize("rows,cin,cout", [(0, 8, 4), (1, 256, 256), (777, 256, 128), (4099, 256, 2048), (513, 40, 12)])
def test_linear_and_linear_relu(dev, rows, cin, cout):
    from ocpg_b200 import epilogue
    g = torch.Generator(device=dev).manual_seed(rows + cout)
    mk = lambda *s: torch.randn(*s, device=dev, generator=g)
    x, w, b, dy = mk(rows, cin), mk(cout, cin) * 0.1, mk(cout), mk(rows, cout)
    for fn, ref_fn in ((epilogue.linear, F.linear), (epilogue.linear_relu, lambda x, w, b: F.relu(F.linear(x, w, b)))):
        if rows == 0 and fn is epilogue.linear_relu:
            continue
        a = [t.clone().requires_grad_(True) for t in (x, w, b)]
        r = [t.double().clone().requires_grad_(True) for t in (x, w, b)]
        y = fn(*a)
        yr = ref_fn(*r)
        y.backward(dy)
        yr.backward(dy.double())
        if rows:
            assert rel(y, yr) <= 1e-5
            for p, q in zip(a, r):
                assert rel(p.grad, q.grad) <= 1e-4 if fn is epilogue.linear else rel(p.grad, q.grad) <= 2e-3   # relu: fp32 sign flips at 0
        else:
            assert a[2].grad.abs().max() == 0


@pytest.mark.parametrize("rows,C", [(0, 4), (1, 4), (5, 12), (1000, 128), (24100, 256), (3001, 2048), (77, 3000)])
def test_column_sum(dev, rows, C):
    from ocpg_b200 import epilogue
    x = torch.randn(rows, C, device=dev)
    got = epilogue.column_sum(x)
    want = x.double().sum(0)
    assert got.shape == (C,)
    assert float((got.double() - want).abs().max()) <= 1e-5 * max(1.0, float(x.double().abs().sum(0).max()))


def test_relu_backward_exact(dev):
    """dpre = dh where h > 0, exactly; zero elsewhere."""
    import ocpg_b200
    L = ocpg_b200.lib()
    rows, C = 1234, 2048
    h = torch.relu(torch.randn(rows, C, device=dev))
    dh = torch.randn(rows, C, device=dev)
    dpre, db = torch.empty_like(dh), torch.empty(C, device=dev)
    rc = L.msda_relu_backward_column_sum_f32(dh.data_ptr(), h.data_ptr(), rows, C, dpre.data_ptr(), db.data_ptr(),
                                             torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    want = torch.where(h > 0, dh, torch.zeros_like(dh))
    assert torch.equal(dpre, want)
    assert rel(db, want.double().sum(0)) <= 1e-5


# ---- dropout inside the epilogue kernels (training: deformable_transformer.py:226-235) ----
# The kernels' mask is their own counter-based stream, so parity is: the reference formula, in fp64, evaluated with the
# mask the kernels report for the same (rng, salt, p) -- plus the statistics and independence of that mask.

def test_dropout_mask_statistics(dev):
    from ocpg_b200 import epilogue
    torch.manual_seed(5)
    rng = epilogue.new_rng(dev)
    n = 1 << 22
    for p in (0.0, 0.1, 0.5, 0.9):
        keep = epilogue.dropout_mask(rng, 1, p, (n,))
        frac = float(keep.float().mean())
        assert abs(frac - (1 - p)) <= 5 * (p * (1 - p) / n) ** 0.5 + 1e-12, (p, frac)
    a = epilogue.dropout_mask(rng, 1, 0.5, (n,))
    assert torch.equal(a, epilogue.dropout_mask(rng, 1, 0.5, (n,)))                 # a pure function of its arguments
    for other in (epilogue.dropout_mask(rng, 2, 0.5, (n,)), epilogue.dropout_mask(epilogue.new_rng(dev), 1, 0.5, (n,))):
        agree = float((a == other).float().mean())                                  # another salt / other words: independent
        assert abs(agree - 0.5) < 5 * 0.5 / n ** 0.5
    # the four lanes of a chunk and neighbouring chunks are uncorrelated
    f = a.float() - 0.5
    for lag in (1, 2, 3, 4, 128):
        assert abs(float((f[:-lag] * f[lag:]).mean())) < 5 * 0.25 / n ** 0.5
    torch.manual_seed(5)
    assert torch.equal(rng, epilogue.new_rng(dev))                                  # reproducible under torch.manual_seed


@pytest.mark.parametrize("C,rows,with_bias,p", [(256, 4099, True, 0.1), (128, 37, False, 0.5), (1024, 513, True, 0.25)])
def test_bias_residual_layer_norm_with_dropout(dev, C, rows, with_bias, p):
    from ocpg_b200 import epilogue
    g = torch.Generator(device=dev).manual_seed(C + rows)
    mk = lambda *s: torch.randn(*s, device=dev, generator=g)
    x, res = mk(rows, C) * 2, mk(rows, C) + 0.5
    bias = mk(C) if with_bias else None
    gamma, beta = 1 + 0.1 * mk(C), 0.1 * mk(C)
    dy = mk(rows, C)
    rng, salt = epilogue.new_rng(dev), 3
    leaves = [t.clone().requires_grad_(True) for t in (x, res, gamma, beta)] + ([bias.clone().requires_grad_(True)] if with_bias else [])
    y = epilogue.bias_residual_layer_norm(leaves[0], leaves[4] if with_bias else None, leaves[1], leaves[2], leaves[3], 1e-5,
                                          rng, salt, p)
    y.backward(dy)
    keep = epilogue.dropout_mask(rng, salt, p, (rows, C)).double()
    ref = [t.double().clone().requires_grad_(True) for t in (x, res, gamma, beta)] + ([bias.double().clone().requires_grad_(True)] if with_bias else [])
    z = ref[1] + (ref[0] + ref[4] if with_bias else ref[0]) * keep / (1 - p)
    yr = F.layer_norm(z, (C,), ref[2], ref[3], 1e-5)
    yr.backward(dy.double())
    assert rel(y, yr) <= 2e-6
    for a, b in zip(leaves, ref):
        assert rel(a.grad, b.grad) <= 2e-5, (a.shape, rel(a.grad, b.grad))
    # dropped elements pass no gradient to x, every element passes it to the residual
    assert float(leaves[0].grad[keep == 0].abs().max()) == 0.0
    assert float((leaves[1].grad != 0).float().mean()) > 0.99


@pytest.mark.parametrize("rows,cin,cout,p", [(777, 256, 128, 0.1), (4099, 256, 2048, 0.1), (513, 40, 12, 0.5)])
def test_linear_relu_with_dropout(dev, rows, cin, cout, p):
    from ocpg_b200 import epilogue
    g = torch.Generator(device=dev).manual_seed(rows + cout)
    mk = lambda *s: torch.randn(*s, device=dev, generator=g)
    x, w, b, dy = mk(rows, cin), mk(cout, cin) * 0.1, mk(cout), mk(rows, cout)
    rng, salt = epilogue.new_rng(dev), 2
    a = [t.clone().requires_grad_(True) for t in (x, w, b)]
    r = [t.double().clone().requires_grad_(True) for t in (x, w, b)]
    y = epilogue.linear_relu(*a, rng, salt, p)
    keep = epilogue.dropout_mask(rng, salt, p, (rows, cout)).double()
    yr = F.relu(F.linear(*r)) * keep / (1 - p)
    y.backward(dy)
    yr.backward(dy.double())
    assert rel(y, yr) <= 1e-5
    assert float(y.detach()[keep == 0].abs().max()) == 0.0
    for q, t in zip(a, r):
        assert rel(q.grad, t.grad) <= 2e-3       # fp32 sign flips of the pre-activation at 0, as without dropout


def test_encoder_layer_dropout_runs_in_the_epilogue_kernels(dev):
    """Training mode with p = 0.1 (the reference's setting) stays on the fused path: masks are re-drawn per call, follow
    torch.manual_seed, vanish in eval mode, and the output statistics match the reference graph's (nn.Dropout)."""
    import ocpg_b200
    from ocpg_b200.encoder import DeformableTransformerEncoderLayer
    from ocpg_b200.workloads import encoder_reference_points, encoder_workload
    torch.manual_seed(0)
    layer = DeformableTransformerEncoderLayer(256, 512, dropout=0.1).to(dev)
    wl = encoder_workload("t", 2, 96, 160)
    shapes = torch.tensor(wl.levels, dtype=torch.int64, device=dev)
    start = torch.cat((shapes.new_zeros(1), (shapes[:, 0] * shapes[:, 1]).cumsum(0)[:-1]))
    src, pos = torch.randn(2, wl.S, 256, device=dev), 0.1 * torch.randn(2, wl.S, 256, device=dev)
    ref = encoder_reference_points(wl.levels, dev)[None].expand(2, -1, -1, -1).contiguous()
    args = (src, pos, ref, shapes, start, None)
    layer.train()
    assert layer._epilogue_ok(src)
    n0 = ocpg_b200.launch_count()
    torch.manual_seed(11); y1 = layer(*args)
    launches = ocpg_b200.launch_count() - n0
    y2 = layer(*args)
    torch.manual_seed(11); y3 = layer(*args)
    assert launches >= 4                                   # operator + two LN epilogues + in-place dropout
    assert not torch.equal(y1, y2) and torch.equal(y1, y3)
    y1.square().mean().backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in layer.parameters())
    # same distribution as the reference graph with nn.Dropout: compare the mean squared deviation from the eval output
    layer.eval()
    y_eval = layer(*args)
    assert torch.equal(y_eval, layer(*args))
    layer.train()
    ours = torch.stack([(layer(*args) - y_eval).square().mean() for _ in range(8)]).mean()
    layer.fused = False
    layer.self_attn.fused = False
    assert not layer._epilogue_ok(src)
    theirs = torch.stack([(layer(*args) - y_eval).square().mean() for _ in range(8)]).mean()
    assert abs(float(ours / theirs) - 1.0) < 0.15, (float(ours), float(theirs))


def test_dropout_masks_change_between_cuda_graph_replays(dev):
    """The dropout key words live in device memory and come from torch's graph-safe generator: a captured training step draws
    a fresh mask on every replay (a mask baked into the graph would silently turn dropout into a fixed pruning)."""
    from ocpg_b200 import epilogue
    rows, C = 513, 256
    x, res = torch.randn(rows, C, device=dev), torch.randn(rows, C, device=dev)
    gamma, beta = torch.ones(C, device=dev), torch.zeros(C, device=dev)

    def f():
        rng = epilogue.new_rng(dev)
        return epilogue.bias_residual_layer_norm(x, None, res, gamma, beta, 1e-5, rng, 1, 0.5), rng

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        f()
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        y, rng = f()
    outs, keys = [], []
    for _ in range(3):
        g.replay()
        torch.cuda.synchronize()
        outs.append(y.clone()); keys.append(rng.clone())
    assert not torch.equal(keys[0], keys[1]) and not torch.equal(keys[1], keys[2])
    assert not torch.equal(outs[0], outs[1]) and not torch.equal(outs[1], outs[2])
    # and each replay is the reference formula under the mask of ITS key
    keep = epilogue.dropout_mask(keys[2], 1, 0.5, (rows, C)).double()
    want = F.layer_norm(res.double() + x.double() * keep / 0.5, (C,), gamma.double(), beta.double(), 1e-5)
    assert rel(outs[2], want) <= 2e-6
