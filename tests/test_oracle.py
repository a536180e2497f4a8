"""CPU suite: pins oracle/ against the golden vectors produced by the reference's own Python
implementation (tests/golden/make_golden.py), and cross-checks the two restatements."""
import numpy as np
import pytest
import torch

import oracle
from oracle.compare import boundary_mask, rel_err


def test_c_oracle_f64_matches_golden(golden):
    g = golden
    out = oracle.c_forward(g["value"], g["shapes"], g["start"], g["loc"], g["attn"], np.float64)
    assert rel_err(out, g["out"]) < 1e-13
    gv, gl, ga = oracle.c_backward(g["value"], g["shapes"], g["start"], g["loc"], g["attn"], g["grad_out"], np.float64)
    assert rel_err(gv, g["grad_value"]) < 1e-13
    assert rel_err(ga, g["grad_attn"]) < 1e-12
    # fp32-valued inputs can sit exactly on a cell boundary only by accident; mask to be safe
    keep = ~boundary_mask(g["loc"], g["shapes"], 1e-9)
    assert rel_err(gl[keep], g["grad_loc"][keep]) < 1e-12
    assert keep.mean() > 0.99


def test_c_oracle_f32_within_reference_tolerance(golden):
    """The f32 instantiation models the reference CUDA kernel's rounding; north_star tolerances:
    forward 1e-5, backward 1e-4 (max-abs error over max-abs reference, vs the fp64 golden)."""
    g = golden
    out = oracle.c_forward(g["value"], g["shapes"], g["start"], g["loc"], g["attn"], np.float32)
    assert out.dtype == np.float32
    assert rel_err(out, g["out"]) < 1e-5
    gv, gl, ga = oracle.c_backward(g["value"], g["shapes"], g["start"], g["loc"], g["attn"], g["grad_out"], np.float32)
    assert rel_err(gv, g["grad_value"]) < 1e-4
    assert rel_err(ga, g["grad_attn"]) < 1e-4
    keep = ~boundary_mask(g["loc"], g["shapes"], 1e-4)
    assert rel_err(gl[keep], g["grad_loc"][keep]) < 1e-4


def test_reference_allclose_criteria_on_its_own_geometry():
    """models/ops/test.py:40 (fp64: torch.allclose default) and :56 (fp32: rtol 1e-2, atol 1e-3)."""
    from conftest import load_golden
    g = load_golden("ref_test")
    out64 = oracle.c_forward(g["value"], g["shapes"], g["start"], g["loc"], g["attn"], np.float64)
    assert torch.allclose(torch.from_numpy(out64), torch.from_numpy(g["out"]))
    out32 = oracle.c_forward(g["value"], g["shapes"], g["start"], g["loc"], g["attn"], np.float32)
    assert torch.allclose(torch.from_numpy(out32).double(), torch.from_numpy(g["out"]), rtol=1e-2, atol=1e-3)


def test_grid_sample_port_matches_golden(golden):
    g = golden
    t = lambda k: torch.from_numpy(g[k]).double()
    out, gv, gl, ga = oracle.msda_grid_sample_fwd_bwd(t("value"), g["shapes"], t("loc"), t("attn"), t("grad_out"))
    assert rel_err(out, g["out"]) < 1e-13
    assert rel_err(gv, g["grad_value"]) < 1e-13
    assert rel_err(gl, g["grad_loc"]) < 1e-12
    assert rel_err(ga, g["grad_attn"]) < 1e-12


@pytest.mark.parametrize("regime", ["init", "uniform"])
def test_restatements_agree_on_production_layout(regime):
    """C restatement vs grid_sample port on a production-layout pyramid (M=8, D=32, L=4, P=4) in
    both location regimes, fp64 (sizes the oracle finishes in well under a second)."""
    from ocpg_b200.workloads import encoder_workload, make_inputs
    wl = encoder_workload("tiny", 2, 72, 104)      # levels (9,13),(5,7),(3,4),(2,2)
    x = make_inputs(wl, regime, seed=5, dtype=torch.float64)
    out, gv, gl, ga = oracle.msda_grid_sample_fwd_bwd(x["value"], x["shapes"], x["loc"], x["attn"], x["grad_out"])
    c_out = oracle.c_forward(x["value"], x["shapes"], x["start"], x["loc"], x["attn"])
    c_gv, c_gl, c_ga = oracle.c_backward(x["value"], x["shapes"], x["start"], x["loc"], x["attn"], x["grad_out"])
    assert rel_err(c_out, out) < 1e-12
    assert rel_err(c_gv, gv) < 1e-12
    assert rel_err(c_ga, ga) < 1e-12
    keep = ~boundary_mask(x["loc"], x["shapes"], 1e-9)
    assert rel_err(c_gl[keep], gl.numpy()[keep]) < 1e-11


def test_empty_inputs():
    """Lq = 0 and N = 0 are legal shapes: outputs are empty / all-zero grads."""
    shapes = np.array([[2, 3]], dtype=np.int64)
    start = np.zeros(1, dtype=np.int64)
    value = np.ones((1, 6, 2, 4))
    loc = np.zeros((1, 0, 2, 1, 2, 2))
    attn = np.zeros((1, 0, 2, 1, 2))
    out = oracle.c_forward(value, shapes, start, loc, attn)
    assert out.shape == (1, 0, 8)
    gv, gl, ga = oracle.c_backward(value, shapes, start, loc, attn, np.zeros((1, 0, 8)))
    assert gv.shape == value.shape and not gv.any() and gl.size == 0 and ga.size == 0
