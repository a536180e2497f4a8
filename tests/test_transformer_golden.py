"""The callers either side of the hot path against golden vectors produced by the REFERENCE's own
encoder / decoder classes run on the CPU in fp64 (tests/golden/make_golden_transformer.py):

  CPU   the decoder-side consumers' torch formulation (the path CPU tensors take) reproduces the reference's
        ``samples_keep`` exactly from the reference's per-layer locations and weights; state_dict keys match;
  GPU   ocpg_b200's re-hosted DeformableTransformerEncoder / DeformableTransformerDecoder (fused operator, epilogue
        kernels, decoder consumer kernels) on the same seeded weights and inputs in fp32.
"""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import transformer_case as tc  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    with np.load(os.path.join(GOLDEN, f"transformer_{name}.npz")) as z:
        return {k: z[k] for k in z.files}


def rel(a, b):
    a, b = (torch.as_tensor(np.asarray(t.detach().cpu() if torch.is_tensor(t) else t)).double() for t in (a, b))
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


# ---------------------------------------------------------------- CPU
@pytest.mark.parametrize("case", ["decoder2", "decoder4"])
def test_consumers_torch_formulation_matches_reference(case):
    from ocpg_b200.decoder import scale_reference_points, select_top_samples
    g, x = load(case), tc.inputs("case")
    for layer in range(tc.N_LAYERS):
        keep, w, idx = select_top_samples(torch.from_numpy(g["loc"][layer]), torch.from_numpy(g["aw"][layer]), x["valid_ratios"], 30)
        assert keep.shape == (tc.N_FRAMES, tc.N_QUERIES, 30, 2) and torch.equal(keep, torch.from_numpy(g["samples"][layer]))
        assert bool((w[..., :-1] >= w[..., 1:]).all())
    ref = x["ref2" if case == "decoder2" else "ref4"]
    rpi = scale_reference_points(ref, x["valid_ratios"])
    vr = x["valid_ratios"]
    want = ref[:, :, None] * (vr if ref.shape[-1] == 2 else torch.cat([vr, vr], -1))[:, None]
    assert torch.equal(rpi, want)


def test_decoder_mirrors_reference_layout():
    """Same sub-module names as deformable_transformer.py:293-352: the golden generator's reference state_dict keys."""
    from ocpg_b200.decoder import build_decoder
    dec = build_decoder(num_layers=tc.N_LAYERS, d_ffn=tc.D_FFN)
    keys = set(dec.state_dict())
    g = load("decoder2")
    ref_keys = {k[len("pgrad/"):] for k in g if k.startswith("pgrad/")}
    assert keys == ref_keys, keys ^ ref_keys
    assert dec.bbox_embed is None and dec.class_embed is None and dec.return_intermediate


# ---------------------------------------------------------------- GPU
@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import ocpg_b200
    ocpg_b200.lib()
    return torch.device("cuda:0")


def to_dev(x, dev):
    return {k: (v.to(dev) if v.dtype in (torch.int64, torch.bool) else v.float().to(dev)) for k, v in x.items()}


def check_param_grads(module, g, tag, tol):
    worst = 0.0
    for k, p in module.named_parameters():
        want = float(g[f"pgrad/{k}"])
        got = tc.projection(f"{tag}/{k}", p.grad)
        scale = float(p.grad.double().norm()) * (p.numel() ** 0.0) + 1e-30       # |<g, r>| ~ ||g||
        worst = max(worst, abs(got - want) / scale)
        assert abs(got - want) <= tol * scale, (k, got, want, scale)
    return worst


@pytest.mark.gpu
@pytest.mark.parametrize("fused", [True, False])
def test_encoder_vs_reference_golden(dev, fused):
    from ocpg_b200.encoder import build_encoder
    g, x = load("encoder"), to_dev(tc.inputs("case"), dev)
    enc = build_encoder(num_layers=tc.N_LAYERS, d_ffn=tc.D_FFN, fused=fused)
    enc.load_state_dict({k: v.float() for k, v in tc.seeded_state_dict(enc, "enc").items()})
    enc = enc.to(dev)
    src = x["src"].clone().requires_grad_(True)
    out = enc(src, x["shapes"], x["start"], x["valid_ratios"], x["pos"], x["mask"])
    out.backward(x["grad_enc"])
    assert rel(out, g["out"]) <= 1e-4, rel(out, g["out"])
    assert rel(src.grad, g["grad_src"]) <= 1e-3, rel(src.grad, g["grad_src"])
    check_param_grads(enc, g, "enc", 3e-3)


@pytest.mark.gpu
@pytest.mark.parametrize("fused", [True, False])
@pytest.mark.parametrize("case", ["decoder2", "decoder4"])
def test_decoder_vs_reference_golden(dev, case, fused):
    import ocpg_b200
    from ocpg_b200.decoder import build_decoder
    g, x = load(case), to_dev(tc.inputs("case"), dev)
    dec = build_decoder(num_layers=tc.N_LAYERS, d_ffn=tc.D_FFN, fused=fused)
    dec.load_state_dict({k: v.float() for k, v in tc.seeded_state_dict(dec, "dec").items()})
    dec = dec.to(dev)
    tgt, memory, refp = (x[k].clone().requires_grad_(True) for k in ("tgt", "src", "ref2" if case == "decoder2" else "ref4"))
    n0 = ocpg_b200.launch_count()
    hs, refs, samples = dec(tgt, refp, memory, x["shapes"], x["start"], x["valid_ratios"], x["query_pos"], x["mask"])
    assert ocpg_b200.launch_count() - n0 >= 3 * tc.N_LAYERS          # operator + the two consumer kernels per layer
    hs.backward(x["grad_hs"])
    assert hs.shape == g["hs"].shape and samples.shape == g["samples"].shape
    assert rel(hs, g["hs"]) <= 1e-4, rel(hs, g["hs"])
    assert rel(refs, g["refs"]) <= 1e-6
    # the selection: fp32 weights can swap two near-equal neighbours of the fp64 order; everything else must coincide
    same = (samples.double().cpu() - torch.from_numpy(g["samples"])).abs().amax(-1) <= 1e-5
    assert float(same.float().mean()) >= 0.97, float(same.float().mean())
    assert rel(tgt.grad, g["grad_tgt"]) <= 1e-3 and rel(memory.grad, g["grad_memory"]) <= 1e-3
    assert rel(refp.grad, g["grad_ref"]) <= 1e-3, rel(refp.grad, g["grad_ref"])
    check_param_grads(dec, g, "dec", 3e-3)


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["decoder2", "decoder4"])
def test_consumer_kernels_match_reference(dev, case):
    """msda_decoder_select_samples_f32 / msda_decoder_reference_points_f32 on the reference's own per-layer locations and
    weights (rounded to fp32): selected points identical wherever the fp32 weights are distinct, locations divided by the
    valid ratios to fp32 accuracy, weights in descending order, indices consistent with both."""
    from ocpg_b200.decoder import _ScaleReferencePoints, _SelectTopSamples
    g, x = load(case), to_dev(tc.inputs("case"), dev)
    for layer in range(tc.N_LAYERS):
        loc = torch.from_numpy(g["loc"][layer]).float().to(dev)
        aw = torch.from_numpy(g["aw"][layer]).float().to(dev)
        keep, w, idx = _SelectTopSamples.apply(loc, aw, x["valid_ratios"], 30)
        N, Lq = loc.shape[:2]
        flat_w, flat_loc = aw.view(N, Lq, -1), (loc / x["valid_ratios"][:, None, None, :, None, :]).view(N, Lq, -1, 2)
        tw, ti = flat_w.topk(30, dim=2)
        assert torch.equal(w, tw)                                               # same multiset, same order of values
        assert torch.equal(torch.gather(flat_w, 2, idx), w)                     # indices point at those weights
        assert bool((idx.sort(-1).values[..., 1:] != idx.sort(-1).values[..., :-1]).all())     # no point taken twice
        assert torch.equal(keep, torch.gather(flat_loc, 2, idx[..., None].expand(-1, -1, -1, 2)))   # true division, bit-exact
        assert rel(keep, g["samples"][layer]) <= 1e-6 or float(((keep.double().cpu() - torch.from_numpy(g["samples"][layer])).abs().amax(-1) <= 1e-5).float().mean()) >= 0.97
    ref = x["ref2" if case == "decoder2" else "ref4"]
    vr = x["valid_ratios"]
    want = ref[:, :, None] * (vr if ref.shape[-1] == 2 else torch.cat([vr, vr], -1))[:, None]
    assert torch.equal(_ScaleReferencePoints.apply(ref, vr), want)


@pytest.mark.gpu
def test_select_samples_ties_and_shapes(dev):
    """All weights equal (the reference's initial state: zero attention-logit weights): ascending index order; K = 32 ... 256;
    gradient through the gather like the reference graph's."""
    from ocpg_b200.decoder import select_top_samples
    for M, L, P, top in ((8, 4, 4, 30), (2, 4, 4, 30), (8, 4, 8, 32), (1, 1, 32, 7), (4, 4, 4, 1)):
        N, Lq, K = 3, 7, M * L * P
        loc = torch.rand(N, Lq, M, L, P, 2, device=dev, requires_grad=True)
        vr = 0.5 + 0.5 * torch.rand(N, L, 2, device=dev)
        aw = torch.full((N, Lq, M, L, P), 1.0 / K, device=dev)
        keep, w, idx = select_top_samples(loc, aw, vr, top)
        assert torch.equal(idx, torch.arange(top, device=dev).expand(N, Lq, top))
        aw = torch.softmax(torch.randn(N, Lq, K, device=dev), -1).view(N, Lq, M, L, P)
        keep, w, idx = select_top_samples(loc, aw, vr, top)
        tw, ti = aw.view(N, Lq, -1).topk(top, dim=2)
        assert torch.equal(w, tw) and torch.equal(idx, ti)
        gk = torch.randn_like(keep)
        keep.backward(gk)
        loc2 = loc.detach().clone().requires_grad_(True)
        ref_keep = torch.gather((loc2 / vr[:, None, None, :, None, :]).view(N, Lq, -1, 2), 2, ti[..., None].expand(-1, -1, -1, 2))
        ref_keep.backward(gk)
        assert torch.equal(keep, ref_keep) and rel(loc.grad, loc2.grad) <= 1e-6
    with pytest.raises(RuntimeError, match="top"):
        import ocpg_b200
        from ocpg_b200 import _lib
        rc = ocpg_b200.lib().msda_decoder_select_samples_f32(1, 1, 1, 1, 1, 8, 4, 16, 30, 1, None, None, None)
        _lib.check(rc, "msda_decoder_select_samples_f32")


# ---------------------------------------------------------------- the whole DeformableTransformer (:26-217)
def _build_full(refine):
    from ocpg_b200.transformer import DeformableTransformer
    model = DeformableTransformer(d_model=tc.D_MODEL, nhead=8, num_encoder_layers=tc.N_LAYERS, num_decoder_layers=tc.N_LAYERS,
                                  dim_feedforward=tc.D_FFN, dropout=0.0, return_intermediate_dec=True)
    if refine:
        model.decoder.bbox_embed = torch.nn.ModuleList([torch.nn.Linear(tc.D_MODEL, 2) for _ in range(tc.N_LAYERS)])
    return model


def test_full_transformer_mirrors_reference_layout():
    g = load("full_refine")
    model = _build_full(True)
    assert set(model.state_dict()) == {k[len("pgrad/"):] for k in g if k.startswith("pgrad/")}
    from ocpg_b200.transformer import build_deforamble_transformer
    import types
    args = types.SimpleNamespace(hidden_dim=256, nheads=8, enc_layers=1, dec_layers=1, dim_feedforward=64, dropout=0.1,
                                 num_feature_levels=4, dec_n_points=4, enc_n_points=4, two_stage=False, num_queries=5)
    m = build_deforamble_transformer(args)
    assert m.decoder.return_intermediate and m.level_embed.shape == (4, 256) and m.reference_points.out_features == 2
    with pytest.raises(NotImplementedError):
        build_deforamble_transformer(types.SimpleNamespace(**{**vars(args), "two_stage": True}))
    # valid ratios as the reference computes them (:125-133)
    mask = torch.ones(2, 4, 6, dtype=torch.bool); mask[0, :3, :4] = False; mask[1, :, :3] = False
    assert torch.equal(m.get_valid_ratio(mask), torch.tensor([[4 / 6, 3 / 4], [3 / 6, 1.0]]))


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["full", "full_refine"])
def test_full_transformer_vs_reference_golden(dev, case):
    g = load(case)
    model = _build_full(case == "full_refine")
    model.load_state_dict({k: v.float() for k, v in tc.seeded_state_dict(model, "full").items()})
    model = model.to(dev).train()
    x = tc.full_inputs("full")
    f32 = lambda t: t.float().to(dev)
    srcs = [f32(s).requires_grad_(True) for s in x["srcs"]]
    tgt, qe = f32(x["tgt"]).requires_grad_(True), f32(x["query_embed"]).requires_grad_(True)
    masks, poss = [m.to(dev) for m in x["masks"]], [f32(p) for p in x["pos_embeds"]]
    hs, memory_features, init_ref, inter_refs, a, b, inter_samples = model(srcs, tgt, masks, poss, qe)
    assert a is None and b is None and len(memory_features) == 3
    loss = (hs * f32(x["grad_hs"])).sum() + sum((m * f32(gm)).sum() for m, gm in zip(memory_features, x["grad_maps"]))
    loss.backward()
    assert rel(hs, g["hs"]) <= 2e-4, rel(hs, g["hs"])
    assert rel(init_ref, g["init_ref"]) <= 1e-5 and rel(inter_refs, g["inter_refs"]) <= 1e-4
    for i, m in enumerate(memory_features):
        assert m.is_contiguous() and rel(m, g[f"memory_{i}"]) <= 2e-4, (i, rel(m, g[f"memory_{i}"]))
    same = (inter_samples.double().cpu() - torch.from_numpy(g["inter_samples"])).abs().amax(-1) <= 1e-4
    assert float(same.float().mean()) >= 0.95, float(same.float().mean())
    assert rel(tgt.grad, g["grad_tgt"]) <= 2e-3 and rel(qe.grad, g["grad_query_embed"]) <= 2e-3
    for i, s in enumerate(srcs):
        assert rel(s.grad, g[f"grad_src_{i}"]) <= 2e-3, (i, rel(s.grad, g[f"grad_src_{i}"]))
    for k, p in model.named_parameters():
        want = float(g[f"pgrad/{k}"])
        if want != want:                                    # NaN: the reference graph never reaches this parameter
            assert p.grad is None, k
            continue
        got = tc.projection(f"full/{k}", p.grad)
        assert abs(got - want) <= 5e-3 * (float(p.grad.double().norm()) + 1e-30), (k, got, want)
    if case == "full_refine":                               # the refinement really moved the reference points
        assert not np.allclose(g["inter_refs"], load("full")["inter_refs"])


@pytest.mark.gpu
def test_select_samples_random_with_ties(dev):
    """Weights quantised to a few values (many exact ties) and random sizes: the selected weights equal torch.topk's, the
    indices are distinct, point at those weights, and among equal weights come in ascending order."""
    import random
    from ocpg_b200.decoder import select_top_samples
    rnd = random.Random(3)
    for case in range(25):
        M, L, P = rnd.choice([(8, 4, 4), (4, 4, 4), (2, 4, 8), (1, 3, 5), (8, 2, 16), (3, 1, 7)])
        K = M * L * P
        top = rnd.randint(1, min(32, K))
        N, Lq = rnd.randint(1, 4), rnd.randint(1, 9)
        levels_q = rnd.choice([2, 3, 5, 1000])
        aw = (torch.rand(N, Lq, K, device=dev) * levels_q).floor().div(levels_q).view(N, Lq, M, L, P)
        if case % 5 == 0:
            aw = aw - 0.5                                   # negative weights order correctly too
        loc = torch.rand(N, Lq, M, L, P, 2, device=dev)
        vr = 0.5 + 0.5 * torch.rand(N, L, 2, device=dev)
        keep, w, idx = select_top_samples(loc, aw, vr, top)
        flat = aw.view(N, Lq, K)
        tw, _ = flat.topk(top, dim=2)
        assert torch.equal(w, tw), case
        assert torch.equal(torch.gather(flat, 2, idx), w)
        srt = idx.sort(-1).values
        assert bool((srt[..., 1:] != srt[..., :-1]).all()) if top > 1 else True
        if top > 1:
            tie = w[..., 1:] == w[..., :-1]
            assert bool((idx[..., 1:][tie] > idx[..., :-1][tie]).all()), case
        want = torch.gather((loc / vr[:, None, None, :, None, :]).view(N, Lq, K, 2), 2, idx[..., None].expand(-1, -1, -1, 2))
        assert torch.equal(keep, want)
