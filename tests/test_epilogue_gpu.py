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


def test_encoder_layer_falls_back_with_dropout(dev):
    """Active dropout disables the fused epilogue (the kernels have no dropout): the layer then runs the reference graph."""
    from ocpg_b200.encoder import DeformableTransformerEncoderLayer
    layer = DeformableTransformerEncoderLayer(256, 512, dropout=0.1).to(dev)
    x = torch.randn(2, 321, 256, device=dev)
    layer.train()
    assert not layer._epilogue_ok(x)
    layer.eval()
    assert layer._epilogue_ok(x)
