"""Seeded weights and inputs shared by tests/golden/make_golden_transformer.py (which runs the REFERENCE's encoder /
decoder classes on them, here, on the CPU, in fp64) and tests/test_transformer_golden_gpu.py (which runs ocpg_b200's
re-hosted classes on them on the GPU).  Everything is a pure function of names and shapes -- no dependence on module
construction order -- and is generated on the CPU in fp64, so both sides see identical values."""
import zlib

import torch

LEVELS = ((8, 12), (4, 6), (2, 3), (1, 2))          # S = 96 + 24 + 6 + 2 = 128
N_FRAMES, N_QUERIES, D_MODEL, D_FFN, N_LAYERS = 2, 5, 256, 512, 2


def _gen(name: str) -> torch.Generator:
    return torch.Generator().manual_seed(zlib.crc32(name.encode()))


def seeded_state_dict(module: torch.nn.Module, tag: str) -> dict:
    """A state_dict for ``module`` whose every tensor is a function of (tag, key, shape): LayerNorm weights near 1,
    biases small, matrices ~ N(0, 1/fan_in); the sampling-offset bias keeps the module's own ring initialisation (so
    points spread around the reference point) plus noise, and the attention-logit weights are NOT zero (at the reference
    initialisation all M*L*P weights are equal and the top-k is all ties)."""
    out = {}
    for key, ref in module.state_dict().items():
        r = torch.randn(ref.shape, generator=_gen(f"{tag}/{key}"), dtype=torch.float64)
        if "norm" in key and key.endswith("weight"):
            t = 1.0 + 0.1 * r
        elif key.endswith("sampling_offsets.bias"):
            t = ref.double() + 0.3 * r
        elif key.endswith("sampling_offsets.weight"):
            t = 0.05 * r
        elif key.endswith("bias") or "in_proj_bias" in key:
            t = 0.1 * r
        else:
            t = r / (ref.shape[-1] ** 0.5)
        out[key] = t
    return out


def inputs(tag: str) -> dict:
    g = lambda name, *shape: torch.randn(*shape, generator=_gen(f"{tag}/in/{name}"), dtype=torch.float64)
    u = lambda name, *shape: torch.rand(*shape, generator=_gen(f"{tag}/in/{name}"), dtype=torch.float64)
    S = sum(h * w for h, w in LEVELS)
    shapes = torch.tensor(LEVELS, dtype=torch.int64)
    start = torch.cat((shapes.new_zeros(1), shapes.prod(1).cumsum(0)[:-1]))
    d = dict(shapes=shapes, start=start,
             src=g("src", N_FRAMES, S, D_MODEL), pos=0.1 * g("pos", N_FRAMES, S, D_MODEL),
             valid_ratios=0.75 + 0.25 * u("vr", N_FRAMES, len(LEVELS), 2),
             mask=u("mask", N_FRAMES, S) < 0.06,
             tgt=g("tgt", N_FRAMES, N_QUERIES, D_MODEL), query_pos=0.5 * g("qpos", N_FRAMES, N_QUERIES, D_MODEL),
             ref2=0.15 + 0.7 * u("ref2", N_FRAMES, N_QUERIES, 2),
             grad_enc=g("genc", N_FRAMES, S, D_MODEL), grad_hs=g("ghs", N_LAYERS, N_FRAMES, N_QUERIES, D_MODEL))
    d["ref4"] = torch.cat((d["ref2"], 0.1 + 0.3 * u("refwh", N_FRAMES, N_QUERIES, 2)), -1)
    return d


def projection(name: str, t: torch.Tensor) -> float:
    """<t, r> with r a fixed N(0,1) tensor named ``name``: a one-number fingerprint of a gradient."""
    r = torch.randn(t.shape, generator=_gen(f"proj/{name}"), dtype=torch.float64)
    return float((t.detach().double().cpu() * r).sum())


def full_inputs(tag: str) -> dict:
    """Inputs of DeformableTransformer.forward: per-level maps (b*t, c, h, w), padding masks with a valid top-left region
    per frame (right / bottom padding like util/misc.py:nested_tensor_from_tensor_list), positional embeddings, the text
    queries tgt (b, t, q, c) and the learned query embedding (q, c); plus cotangents for hs and the returned memory maps."""
    g = lambda name, *shape: torch.randn(*shape, generator=_gen(f"{tag}/in/{name}"), dtype=torch.float64)
    b, t = 1, N_FRAMES
    valid = [(0.8, 0.9), (1.0, 0.7)]                      # per frame: (valid height, valid width) fraction
    srcs, masks, poss = [], [], []
    for l, (h, w) in enumerate(LEVELS):
        srcs.append(g(f"src{l}", b * t, D_MODEL, h, w))
        poss.append(0.1 * g(f"pos{l}", b * t, D_MODEL, h, w))
        m = torch.ones(b * t, h, w, dtype=torch.bool)
        for n, (fh, fw) in enumerate(valid):
            m[n, :max(1, round(fh * h)), :max(1, round(fw * w))] = False
        masks.append(m)
    return dict(srcs=srcs, masks=masks, pos_embeds=poss, tgt=g("tgt", b, t, N_QUERIES, D_MODEL),
                query_embed=g("qe", N_QUERIES, D_MODEL), grad_hs=g("ghs", N_LAYERS, b * t, N_QUERIES, D_MODEL),
                grad_maps=[g(f"gmap{l}", b * t, D_MODEL, h, w) for l, (h, w) in enumerate(LEVELS[:-1])])
