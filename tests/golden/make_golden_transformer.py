"""Generates tests/golden/transformer_{encoder,decoder2,decoder4}.npz by running the REFERENCE's own
DeformableTransformerEncoder / DeformableTransformerDecoder classes (models/deformable_transformer.py:220-398), unmodified,
on the CPU in fp64.

Run in the build container only (it reads /root/reference):    python tests/golden/make_golden_transformer.py

How the reference is made to run here: its compiled op does not exist (CUDA-only, sm_86 egg), so
``MultiScaleDeformableAttention`` is stubbed and the name ``MSDeformAttnFunction`` inside the reference's
models/ops/modules/ms_deform_attn.py is pointed at a shim whose ``apply`` calls the reference's OWN CPU formulation,
``ms_deform_attn_core_pytorch`` (functions/ms_deform_attn_func.py:41-61) -- the function the reference's test.py uses as
ground truth for that op.  ``models`` / ``models.ops`` are entered as bare namespace packages so that the reference's
models/__init__.py (which imports the whole model zoo) does not run.  Everything else -- MSDeformAttn, the encoder and
decoder layers, the reference-point arithmetic, the top-30 selection -- is the reference's code.

Weights and inputs: tests/golden/transformer_case.py (pure functions of names and shapes).  Stored: the outputs in fp64
and, for the gradients, full tensors for the inputs and one random projection per parameter.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import transformer_case as tc  # noqa: E402

REF = "/root/reference"


def load_reference():
    for name in [m for m in sys.modules if m == "models" or m.startswith("models.") or m == "util" or m.startswith("util.")]:
        del sys.modules[name]
    sys.path.insert(0, REF)
    sys.modules["MultiScaleDeformableAttention"] = types.ModuleType("MultiScaleDeformableAttention")
    for name, path in (("models", f"{REF}/models"), ("models.ops", f"{REF}/models/ops")):
        pkg = types.ModuleType(name)
        pkg.__path__ = [path]
        sys.modules[name] = pkg
    import models.ops.modules.ms_deform_attn as ref_mod
    from models.ops.functions.ms_deform_attn_func import ms_deform_attn_core_pytorch

    class CpuFunction:                                        # the reference's own CPU path behind the operator's signature
        @staticmethod
        def apply(value, shapes, start, loc, aw, im2col_step):
            return ms_deform_attn_core_pytorch(value, shapes, loc, aw)

    ref_mod.MSDeformAttnFunction = CpuFunction
    import models.deformable_transformer as dt
    assert dt.__file__.startswith(REF), dt.__file__
    return dt


def param_projections(module, tag):
    # parameters the graph never reaches (the refinement heads: their output is detached, :388) are recorded as NaN
    return {f"pgrad/{k}": (tc.projection(f"{tag}/{k}", p.grad) if p.grad is not None else float("nan"))
            for k, p in module.named_parameters()}


def main():
    dt = load_reference()
    torch.set_default_dtype(torch.float64)
    x = tc.inputs("case")
    # ---- encoder (deformable_transformer.py:220-290) ----
    layer = dt.DeformableTransformerEncoderLayer(tc.D_MODEL, tc.D_FFN, 0.0, "relu", len(tc.LEVELS), 8, 4)
    enc = dt.DeformableTransformerEncoder(layer, tc.N_LAYERS).double()
    enc.load_state_dict(tc.seeded_state_dict(enc, "enc"))
    src = x["src"].clone().requires_grad_(True)
    out = enc(src, x["shapes"], x["start"], x["valid_ratios"], x["pos"], x["mask"])
    out.backward(x["grad_enc"])
    np.savez_compressed(os.path.join(HERE, "transformer_encoder.npz"), out=out.detach().numpy(),
                        grad_src=src.grad.numpy().astype(np.float32), **param_projections(enc, "enc"))
    print("encoder", tuple(out.shape), float(out.abs().max()))
    # ---- decoder (deformable_transformer.py:293-398), 2-d reference points and 4-d boxes ----
    for name, ref_key in (("decoder2", "ref2"), ("decoder4", "ref4")):
        layer = dt.DeformableTransformerDecoderLayer(tc.D_MODEL, tc.D_FFN, 0.0, "relu", len(tc.LEVELS), 8, 4)
        dec = dt.DeformableTransformerDecoder(layer, tc.N_LAYERS, return_intermediate=True).double()
        dec.load_state_dict(tc.seeded_state_dict(dec, "dec"))
        tgt, memory, refp = (x[k].clone().requires_grad_(True) for k in ("tgt", "src", ref_key))
        seen = []           # every layer's (sampling_locations, attention_weights): the inputs of the rank-3 consumers
        for lyr in dec.layers:
            lyr.register_forward_hook(lambda m, a, o: seen.append((o[1].detach().numpy(), o[2].detach().numpy())))
        hs, refs, samples = dec(tgt, refp, memory, x["shapes"], x["start"], x["valid_ratios"], x["query_pos"], x["mask"])
        hs.backward(x["grad_hs"])
        np.savez_compressed(os.path.join(HERE, f"transformer_{name}.npz"), hs=hs.detach().numpy(), refs=refs.detach().numpy(),
                            loc=np.stack([a for a, _ in seen]), aw=np.stack([b for _, b in seen]),
                            samples=samples.detach().numpy(), grad_tgt=tgt.grad.numpy(),
                            grad_memory=memory.grad.numpy().astype(np.float32), grad_ref=refp.grad.numpy(),
                            **param_projections(dec, "dec"))
        print(name, tuple(hs.shape), tuple(samples.shape), float(hs.abs().max()))
    full_transformer(dt)


def full_transformer(dt):
    """The whole DeformableTransformer (deformable_transformer.py:26-217), plus the decoder's iterative refinement branch
    (:378-388) through a ``bbox_embed`` of plain Linears."""
    torch.manual_seed(0)
    for name, refine in (("full", False), ("full_refine", True)):
        model = dt.DeformableTransformer(d_model=tc.D_MODEL, nhead=8, num_encoder_layers=tc.N_LAYERS, num_decoder_layers=tc.N_LAYERS,
                                         dim_feedforward=tc.D_FFN, dropout=0.0, return_intermediate_dec=True).double()
        if refine:
            model.decoder.bbox_embed = torch.nn.ModuleList([torch.nn.Linear(tc.D_MODEL, 2) for _ in range(tc.N_LAYERS)]).double()
        model.load_state_dict(tc.seeded_state_dict(model, "full"))
        x = tc.full_inputs("full")
        srcs = [s.clone().requires_grad_(True) for s in x["srcs"]]
        tgt, qe = x["tgt"].clone().requires_grad_(True), x["query_embed"].clone().requires_grad_(True)
        hs, memory_features, init_ref, inter_refs, _, _, inter_samples = model(srcs, tgt, x["masks"], x["pos_embeds"], qe)
        loss_terms = [(hs * x["grad_hs"]).sum()] + [(m * g).sum() for m, g in zip(memory_features, x["grad_maps"])]
        sum(loss_terms).backward()
        out = dict(hs=hs.detach().numpy(), init_ref=init_ref.detach().numpy(), inter_refs=inter_refs.detach().numpy(),
                   inter_samples=inter_samples.detach().numpy(), grad_tgt=tgt.grad.numpy(), grad_query_embed=qe.grad.numpy())
        for i, m in enumerate(memory_features):
            out[f"memory_{i}"] = m.detach().numpy().astype(np.float32)
        for i, s_ in enumerate(srcs):
            out[f"grad_src_{i}"] = s_.grad.numpy().astype(np.float32)
        np.savez_compressed(os.path.join(HERE, f"transformer_{name}.npz"), **out, **param_projections(model, "full"))
        print(name, tuple(hs.shape), [tuple(m.shape) for m in memory_features], float(hs.abs().max()))


if __name__ == "__main__":
    main()
