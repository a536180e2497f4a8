"""CPU suite for the drop-in boundary: the C-ABI library loads and exports every symbol the header
declares, the Python mirror of the reference interface has the reference's names / signatures / error
behaviour, and the product never touches oracle/.  No compute calls (no GPU here)."""
import ctypes
import inspect
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "msda_sm100.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(msda_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_header_symbol():
    import ocpg_b200
    from ocpg_b200 import _lib
    ocpg_b200.build()
    syms = header_symbols()
    assert syms == sorted(_lib.SYMBOLS), (syms, sorted(_lib.SYMBOLS))
    L = ctypes.CDLL(_lib.LIB_PATH)
    for s in syms:
        assert hasattr(L, s), s
    assert ocpg_b200.lib().msda_abi_version() == 4


def test_library_is_sm100a_only():
    import subprocess
    from ocpg_b200 import _lib
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_kernel_plan():
    import ocpg_b200
    L = ocpg_b200.lib()
    assert L.msda_kernel_plan(4, 8, 32, 4, 4) == 1       # production layout -> tiled sm_100a kernel
    assert L.msda_kernel_plan(2, 8, 32, 4, 4) == 1
    assert L.msda_kernel_plan(8, 8, 32, 4, 4) == 0       # fp64 -> generic
    assert L.msda_kernel_plan(4, 8, 64, 4, 4) == 0
    assert L.msda_kernel_plan(4, 8, 32, 4, 16) == 0


def test_argument_errors_do_not_touch_the_gpu():
    import ocpg_b200
    L = ocpg_b200.lib()
    rc = L.msda_forward_f32(None, None, None, None, None, 1, 0, 8, 32, 4, 1, 4, None, None)   # S = 0
    assert rc == -1 and b"dimensions" in L.msda_last_error()
    assert L.msda_set_option(b"no_such_key", 1) == -1
    # empty batch / no queries: success without any pointer
    assert L.msda_forward_f32(None, None, None, None, None, 0, 10, 8, 32, 4, 5, 4, None, None) == 0
    assert L.msda_forward_f64(None, None, None, None, None, 2, 10, 8, 32, 4, 0, 4, None, None) == 0


def test_product_never_imports_the_oracle():
    bad = []
    for base in ("ocpg_b200", "models"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".h", ".cuh", ".cpp")):
                    txt = open(os.path.join(dp, f)).read()
                    if re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M) or "oracle/_ref" in txt:
                        bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_operator_api_matches_reference_names():
    from ocpg_b200 import MSDeformAttnFunction, MSDeformAttn
    import ocpg_b200.MultiScaleDeformableAttention as MSDA
    # reference: functions/ms_deform_attn_func.py:23, :32 ; vision.cpp:14-15 ; modules/ms_deform_attn.py:32, :80
    assert list(inspect.signature(MSDeformAttnFunction.forward).parameters) == [
        "ctx", "value", "value_spatial_shapes", "value_level_start_index", "sampling_locations", "attention_weights",
        "im2col_step"]
    assert list(inspect.signature(MSDA.ms_deform_attn_forward).parameters) == [
        "value", "spatial_shapes", "level_start_index", "sampling_loc", "attn_weight", "im2col_step"]
    assert list(inspect.signature(MSDA.ms_deform_attn_backward).parameters) == [
        "value", "spatial_shapes", "level_start_index", "sampling_loc", "attn_weight", "grad_output", "im2col_step"]
    assert list(inspect.signature(MSDeformAttn.__init__).parameters) == ["self", "d_model", "n_levels", "n_heads", "n_points"]
    assert list(inspect.signature(MSDeformAttn.forward).parameters) == [
        "self", "query", "reference_points", "input_flatten", "input_spatial_shapes", "input_level_start_index",
        "input_padding_mask"]


def test_reference_import_paths_resolve():
    from models.ops.modules import MSDeformAttn                     # deformable_transformer.py:20
    from models.ops.functions import MSDeformAttnFunction           # modules/ms_deform_attn.py:22
    import MultiScaleDeformableAttention as MSDA                    # functions/ms_deform_attn_func.py:18
    import ocpg_b200
    assert MSDeformAttn is ocpg_b200.MSDeformAttn and MSDeformAttnFunction is ocpg_b200.MSDeformAttnFunction
    assert MSDA.ms_deform_attn_forward is ocpg_b200.MultiScaleDeformableAttention.ms_deform_attn_forward
    from models.deformable_transformer import build_deforamble_transformer, DeformableTransformer      # ocpg.py:17
    import ocpg_b200.transformer
    assert DeformableTransformer is ocpg_b200.transformer.DeformableTransformer and callable(build_deforamble_transformer)


def test_cpu_tensors_raise_like_the_reference():
    import ocpg_b200.MultiScaleDeformableAttention as MSDA
    v = torch.zeros(1, 6, 2, 4)
    shapes = torch.tensor([[2, 3]]); start = torch.tensor([0])
    loc = torch.zeros(1, 3, 2, 1, 2, 2); attn = torch.zeros(1, 3, 2, 1, 2)
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):       # ms_deform_attn.h:54
        MSDA.ms_deform_attn_forward(v, shapes, start, loc, attn, 64)
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        MSDA.ms_deform_attn_backward(v, shapes, start, loc, attn, torch.zeros(1, 3, 8), 64)
    with pytest.raises(RuntimeError, match="contiguous"):                       # cu:28
        MSDA.ms_deform_attn_forward(v.transpose(1, 2), shapes, start, loc, attn, 64)


def test_module_parameters_and_init_match_reference_scheme():
    import math
    from ocpg_b200 import MSDeformAttn
    torch.manual_seed(0)
    m = MSDeformAttn(256, 4, 8, 4)
    sd = m.state_dict()
    assert sorted(sd) == ["attention_weights.bias", "attention_weights.weight", "output_proj.bias", "output_proj.weight",
                          "sampling_offsets.bias", "sampling_offsets.weight", "value_proj.bias", "value_proj.weight"]
    assert sd["sampling_offsets.weight"].shape == (8 * 4 * 4 * 2, 256) and not sd["sampling_offsets.weight"].any()
    assert sd["attention_weights.weight"].shape == (8 * 4 * 4, 256) and not sd["attention_weights.weight"].any()
    assert not sd["attention_weights.bias"].any() and not sd["value_proj.bias"].any() and not sd["output_proj.bias"].any()
    bias = sd["sampling_offsets.bias"].view(8, 4, 4, 2)
    # reference ms_deform_attn.py:64-70: head h points along angle 2*pi*h/8 (max-norm 1), point i scaled by i+1
    for h in range(8):
        th = 2 * math.pi * h / 8
        d = torch.tensor([math.cos(th), math.sin(th)])
        d = d / d.abs().max()
        for i in range(4):
            assert torch.allclose(bias[h, :, i], (d * (i + 1)).expand(4, 2), atol=1e-6)
    assert m.im2col_step == 64
    with pytest.raises(ValueError):
        MSDeformAttn(250, 4, 8, 4)


def test_workload_shapes_and_bytes():
    from ocpg_b200.workloads import A2D_ENCODER, YTVOS_ENCODER, A2D_DECODER, shard_frames
    assert A2D_ENCODER.levels == ((45, 80), (23, 40), (12, 20), (6, 10)) and A2D_ENCODER.S == 4820
    assert YTVOS_ENCODER.levels == ((80, 144), (40, 72), (20, 36), (10, 18)) and YTVOS_ENCODER.S == 15300
    fwd, bwd = A2D_ENCODER.algorithmic_bytes()
    assert fwd == 24100 * 3584 and bwd == 24100 * 6144            # 3584 / 6144 B per query
    assert A2D_DECODER.queries == 25
    covered = []
    for r in range(3):
        f, n = shard_frames(10, 3, r)
        covered += list(range(f, f + n))
    assert covered == list(range(10))


def test_encoder_mirrors_reference_layout():
    """The encoder harness keeps the reference's sub-module names (deformable_transformer.py:220-290) so its
    state_dict keys are the reference's; 6 layers x (MSDeformAttn + 2 LayerNorm + FFN) = 7 693 056 weights
    (SURVEY.md section 8e)."""
    from ocpg_b200.encoder import DeformableTransformerEncoder, build_encoder
    enc = build_encoder(num_layers=6, d_ffn=2048)
    keys = set(enc.state_dict())
    for k in ("layers.0.self_attn.sampling_offsets.weight", "layers.5.self_attn.output_proj.bias", "layers.2.norm1.weight",
              "layers.3.linear1.weight", "layers.4.linear2.bias", "layers.1.norm2.bias"):
        assert k in keys, k
    assert sum(p.numel() for p in enc.parameters()) == 7693056
    shapes = torch.tensor([(3, 4), (2, 2)])
    ref = DeformableTransformerEncoder.get_reference_points(shapes, torch.ones(2, 2, 2), "cpu")
    assert ref.shape == (2, 16, 2, 2)
    assert torch.allclose(ref[0, 0, 0], torch.tensor([0.5 / 4, 0.5 / 3])) and torch.allclose(ref[0, 12, 1], torch.tensor([0.25, 0.25]))


def test_encoder_reference_points_match_reference_loop():
    """The cached / concatenated formulation gives, bit for bit, what the reference's per-level loop gives
    (deformable_transformer.py:268-281), valid ratios < 1 included; a second call hits the cache."""
    from ocpg_b200.encoder import DeformableTransformerEncoder
    torch.manual_seed(0)
    shapes = torch.tensor([(5, 7), (3, 4), (2, 2), (1, 1)])
    vr = 0.5 + 0.5 * torch.rand(3, 4, 2)

    def reference(spatial_shapes, valid_ratios):
        out = []
        for lvl, (H_, W_) in enumerate(spatial_shapes):
            ref_y, ref_x = torch.meshgrid(torch.linspace(0.5, H_ - 0.5, H_, dtype=torch.float32),
                                          torch.linspace(0.5, W_ - 0.5, W_, dtype=torch.float32), indexing="ij")
            ref_y = ref_y.reshape(-1)[None] / (valid_ratios[:, None, lvl, 1] * H_)
            ref_x = ref_x.reshape(-1)[None] / (valid_ratios[:, None, lvl, 0] * W_)
            out.append(torch.stack((ref_x, ref_y), -1))
        pts = torch.cat(out, 1)
        return pts[:, :, None] * valid_ratios[:, None]

    want = reference(shapes, vr)
    for _ in range(2):
        got = DeformableTransformerEncoder.get_reference_points(shapes, vr, "cpu")
        assert got.shape == want.shape == (3, 35 + 12 + 4 + 1, 4, 2) and torch.equal(got, want)
    assert torch.equal(DeformableTransformerEncoder.get_reference_points(shapes.tolist(), vr, "cpu"), want)


def test_host_shapes_cache_is_keyed_by_object_and_version():
    """The host copy of spatial_shapes is reused only for the same tensor object at the same version: a new tensor (even at
    a recycled address) or an in-place edit is read again; tensors built by flatten_levels never need the copy."""
    from ocpg_b200._shapes import remember_host_shapes, shapes_on_host
    a = torch.tensor([(3, 4), (2, 2)])
    assert shapes_on_host(a) == [(3, 4), (2, 2)] and shapes_on_host(a) is shapes_on_host(a)
    a[0, 0] = 5                                             # in-place: version bump
    assert shapes_on_host(a) == [(5, 4), (2, 2)]
    b = torch.tensor([(7, 1), (1, 1)])
    assert shapes_on_host(b) == [(7, 1), (1, 1)]
    assert shapes_on_host([(2, 3)]) == [(2, 3)] and shapes_on_host(torch.tensor([(2, 3)]).tolist()) == [(2, 3)]
    c = remember_host_shapes(torch.tensor([(9, 9)]), [(9, 9)])
    assert shapes_on_host(c) == [(9, 9)]
    c[0, 1] = 8                                             # the tag is version-checked too
    assert shapes_on_host(c) == [(9, 8)]
    from ocpg_b200.flatten import flatten_levels
    _, _, shapes, _ = flatten_levels([torch.zeros(1, 4, 2, 3), torch.zeros(1, 4, 1, 2)])
    assert getattr(shapes, "_ocpg_host_shapes")[1] == [(2, 3), (1, 2)]
    import ocpg_b200
    m = ocpg_b200.MSDeformAttn(64, 2, 2, 2)
    m._check_shapes(shapes, 8)
    with pytest.raises(AssertionError):
        m._check_shapes(shapes, 9)                          # reference :94


def test_side_kernel_argument_errors_do_not_touch_the_gpu():
    """Every entry point around the operator validates before it launches: bad arguments return MSDA_ERR_* (and empty work
    returns 0) without a device -- the C ABI's error contract (include/msda_sm100.h), checked here on the CPU."""
    import ctypes
    import ocpg_b200
    L = ocpg_b200.lib()
    INVALID, UNSUPPORTED = -1, -2
    one = 4096                                              # any non-null, 16-byte aligned "pointer": never dereferenced
    # epilogue
    assert L.msda_epilogue_ln_forward_f32(one, None, one, one, one, 1e-5, 5, 100, one, one, one, one, None) == UNSUPPORTED
    assert L.msda_epilogue_ln_forward_f32(one, None, one, one, one, 1e-5, -1, 256, one, one, one, one, None) == INVALID
    assert L.msda_epilogue_ln_forward_f32(None, None, None, None, None, 1e-5, 0, 256, None, None, None, None, None) == 0
    assert L.msda_epilogue_ln_forward_f32(one + 4, None, one, one, one, 1e-5, 5, 256, one, one, one, one, None) == INVALID
    assert b"aligned" in L.msda_last_error()
    assert L.msda_epilogue_ln_backward_f32(one, one, one, one, one, 5, 256, one, None, one, None, None) == INVALID   # no grad_gamma
    assert L.msda_column_sum_f32(one, 5, 6, one, None) == INVALID                                                     # channels % 4
    # dropout
    assert L.msda_epilogue_ln_dropout_forward_f32(one, None, one, one, one, 1e-5, 5, 256, None, 1, 0.1, one, one, one, one, None) == INVALID
    assert L.msda_epilogue_ln_dropout_forward_f32(one, None, one, one, one, 1e-5, 5, 256, one, 1, 1.0, one, one, one, one, None) == INVALID
    assert b"[0, 1)" in L.msda_last_error()
    assert L.msda_dropout_inplace_f32(one, 6, one, 1, 0.1, None) == INVALID                                           # n % 4
    assert L.msda_dropout_inplace_f32(None, 0, one, 1, 0.1, None) == 0
    assert L.msda_relu_dropout_backward_column_sum_f32(one, one, -0.5, 5, 8, one, one, None) == INVALID
    assert L.msda_dropout_mask_u8(one, 1, 0.5, 8, None, None) == INVALID
    # decoder consumers
    assert L.msda_decoder_select_samples_f32(one, one, one, 1, 1, 8, 4, 4, 33, one, None, None, None) == UNSUPPORTED  # top > 32
    assert L.msda_decoder_select_samples_f32(one, one, one, 1, 1, 8, 4, 16, 30, one, None, None, None) == UNSUPPORTED # K = 512
    assert L.msda_decoder_select_samples_f32(one, one, one, 1, 1, 1, 2, 2, 5, one, None, None, None) == UNSUPPORTED   # top > K
    assert L.msda_decoder_select_samples_f32(None, None, None, 0, 5, 8, 4, 4, 30, None, None, None, None) == 0
    assert L.msda_decoder_select_samples_f32(None, one, one, 1, 1, 8, 4, 4, 30, one, None, None, None) == INVALID
    assert L.msda_decoder_reference_points_f32(one, one, 1, 1, 4, 3, one, None) == INVALID                            # ref_dim
    assert L.msda_decoder_reference_points_f32(one + 8, one, 1, 1, 4, 4, one, None) == INVALID                        # 16-byte alignment for boxes
    # flattening
    hw = (ctypes.c_int * 2)(4, 4)
    ptrs = (ctypes.c_void_p * 2)(one, one)
    assert L.msda_flatten_levels_f32(9, ptrs, None, None, hw, hw, 1, 8, one, None, None) == UNSUPPORTED
    assert L.msda_flatten_levels_f32(2, ptrs, None, None, None, hw, 1, 8, one, None, None) == INVALID
    assert L.msda_flatten_levels_f32(2, ptrs, ptrs, None, hw, hw, 1, 8, one, None, None) == INVALID                   # pos without pos_flatten
    assert L.msda_flatten_levels_f32(2, ptrs, None, None, hw, hw, 0, 8, None, None, None) == 0
    bad = (ctypes.c_void_p * 2)(one, None)
    assert L.msda_flatten_levels_f32(2, bad, None, None, hw, hw, 1, 8, one, None, None) == INVALID
    assert L.msda_unflatten_levels_f32(2, one, hw, hw, 1, 8, 10, ptrs, None) == INVALID                               # spatial_size < pixels
    odd = (ctypes.c_void_p * 2)(one + 4, one)
    assert L.msda_unflatten_levels_f32(2, one, hw, hw, 1, 8, 0, odd, None) == INVALID                                 # 16-byte kernels need alignment


def test_strict_mode_raises_instead_of_falling_back_to_torch():
    """The helpers around the operator run a torch formulation for layouts their kernels do not cover (here: CPU
    tensors); in strict mode that is an error, so a benchmark cannot measure torch by accident."""
    import torch
    import ocpg_b200
    from ocpg_b200 import decoder, epilogue, flatten
    x, w, b = torch.randn(6, 256), torch.randn(8, 256), torch.randn(8)
    before = ocpg_b200.fallback_counts().get("linear", 0)
    assert epilogue.linear(x, w, b).shape == (6, 8)
    assert ocpg_b200.fallback_counts()["linear"] == before + 1
    ocpg_b200.set_strict(True)
    try:
        assert ocpg_b200.is_strict()
        for call in (lambda: epilogue.linear(x, w, b), lambda: epilogue.linear_relu(x, w, b),
                     lambda: epilogue.bias_residual_layer_norm(x, None, x, torch.ones(256), torch.zeros(256)),
                     lambda: flatten.flatten_levels([torch.randn(1, 4, 2, 2)]),
                     lambda: flatten.unflatten_levels(torch.randn(1, 4, 4), [(2, 2)]),
                     lambda: decoder.scale_reference_points(torch.rand(1, 2, 2), torch.rand(1, 4, 2)),
                     lambda: decoder.select_top_samples(torch.rand(1, 2, 8, 4, 4, 2), torch.rand(1, 2, 8, 4, 4), torch.rand(1, 4, 2))):
            with pytest.raises(RuntimeError, match="strict mode"):
                call()
    finally:
        ocpg_b200.set_strict(False)


def test_source_fingerprint_ignores_comments_and_tags_the_committed_capture():
    """profiles/traffic.json entries are only used by bench.py when they were captured from the kernel code built now."""
    import json
    import ocpg_b200
    from ocpg_b200 import _lib
    assert _lib._code_only('int a = 1; // note\n/* block */ const char *s = "// kept";\n') == 'inta=1;constchar*s="// kept";'
    fp = ocpg_b200.source_fingerprint()
    assert len(fp) == 16 and fp == ocpg_b200.source_fingerprint()
    with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
        table = json.load(f)
    entries = [v for k, v in table.items() if not k.startswith("_")]
    assert entries and all("kernel_sources" in e for e in entries)
