"""Whole-step CUDA-graph capture (ocpg_b200/graph.py): a captured forward + backward of the re-hosted encoder gives the
eager step's gradients, replays follow in-place input updates, and dropout draws new masks per replay."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import ocpg_b200
    ocpg_b200.lib()
    return torch.device("cuda:0")


def _setup(dev, dropout):
    from ocpg_b200.encoder import build_encoder
    from ocpg_b200.workloads import encoder_workload
    torch.manual_seed(0)
    wl = encoder_workload("t", 2, 96, 160)
    enc = build_encoder(num_layers=2, d_ffn=256, dropout=dropout).to(dev).train()
    with torch.no_grad():
        for layer in enc.layers:
            layer.self_attn.sampling_offsets.weight.normal_(0, 0.02)
            layer.self_attn.attention_weights.weight.normal_(0, 0.05)
    shapes = torch.tensor(wl.levels, dtype=torch.int64, device=dev)
    start = torch.cat((shapes.new_zeros(1), shapes.prod(1).cumsum(0)[:-1]))
    src = torch.randn(2, wl.S, 256, device=dev, requires_grad=True)
    pos = 0.1 * torch.randn(2, wl.S, 256, device=dev)
    vr = 0.8 + 0.2 * torch.rand(2, 4, 2, device=dev)
    g = torch.randn(2, wl.S, 256, device=dev)

    def step():
        src.grad = None
        out = enc(src, shapes, start, vr, pos, None)
        out.backward(g)
        return out
    return enc, src, step


def test_graphed_step_matches_eager_and_tracks_inputs(dev):
    from ocpg_b200.graph import GraphedStep
    enc, src, step = _setup(dev, 0.0)
    for p in enc.parameters():
        p.grad = None
    out_e = step().detach().clone()
    grads_e = [p.grad.clone() for p in enc.parameters()]
    gsrc_e = src.grad.clone()
    graphed = GraphedStep(step, params=enc.parameters())
    out_g = graphed()
    torch.cuda.synchronize()
    close = lambda a, b: float((a - b).abs().max()) <= 2e-5 * float(b.abs().max()) + 1e-7
    assert torch.equal(out_g, out_e)                                   # the forward has no atomics: bitwise
    assert close(src.grad, gsrc_e) and all(close(p.grad, ge) for p, ge in zip(enc.parameters(), grads_e))
    with torch.no_grad():                                              # new data in the same buffer
        src.mul_(0.5)
    out_g2 = graphed().clone()
    for p in enc.parameters():
        p.grad = None
    out_e2 = step().detach()
    assert torch.equal(out_g2, out_e2) and not torch.equal(out_g2, out_e)


def test_graphed_step_redraws_dropout(dev):
    from ocpg_b200.graph import GraphedStep
    enc, src, step = _setup(dev, 0.1)
    graphed = GraphedStep(step, params=enc.parameters())
    a = graphed().clone()
    b = graphed().clone()
    torch.cuda.synchronize()
    assert not torch.equal(a, b)
    assert all(torch.isfinite(p.grad).all() for p in enc.parameters())
