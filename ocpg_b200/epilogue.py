"""Encoder-layer epilogue operators on the C ABI of include/msda_sm100.h.

The reference's DeformableTransformerEncoderLayer (models/deformable_transformer.py:243-260) surrounds the attention and
the FFN with ``norm(src + dropout(linear(x)))``.  In eager PyTorch the parameter gradients of that pattern -- LayerNorm's
gamma/beta kernel and one column-sum kernel per Linear bias -- take a third of a TF32 encoder step
(profiles/r1_encoder_kernel_breakdown.txt).  The autograd Functions below compute the same values with HBM-streaming
kernels; the GEMMs themselves stay with cuBLAS (``torch.matmul`` / ``F.linear``).

    bias_residual_layer_norm(x, bias, residual, gamma, beta, eps)   == F.layer_norm(residual + (x + bias), ...)
    linear(x, weight, bias)                                         == F.linear(x, weight, bias)
    linear_relu(x, weight, bias)                                    == F.relu(F.linear(x, weight, bias))

With dropout active (training, p = 0.1 in the reference: deformable_transformer.py:226-235) the same kernels take a
``(rng, salt, p)`` triple: ``rng = new_rng(device)`` draws two 64-bit words from torch's CUDA generator (one tiny
kernel per layer; reproducible under ``torch.manual_seed``, graph-capturable), the keep mask is a counter-based hash of
them that the backward regenerates, never a stored tensor.  The mask differs from the one ``torch.dropout`` would draw
-- as it does between any two dropout implementations -- so parity is checked against the reference formula evaluated
with the mask the kernels used (``dropout_mask``).

fp32 CUDA tensors only (the callers fall back to the torch ops otherwise).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _lib, _strict

LN_CHANNELS = (128, 256, 512, 1024)


def _stream(t) -> int:
    return _lib.raw_stream(t.device)


def supported(*tensors) -> bool:
    """fp32 CUDA tensors outside autocast: what the epilogue kernels take."""
    if torch.is_autocast_enabled():
        return False
    return all(t is None or (t.is_cuda and t.dtype == torch.float32) for t in tensors)


def column_sum(x2d: torch.Tensor) -> torch.Tensor:
    """sum over rows of a contiguous (rows, C) fp32 matrix, C % 4 == 0."""
    rows, C = x2d.shape
    out = torch.empty(C, dtype=torch.float32, device=x2d.device)
    with _lib.on_device(x2d.device):
        rc = _lib.lib().msda_column_sum_f32(x2d.data_ptr(), rows, C, out.data_ptr(), _stream(x2d))
    _lib.check(rc, "msda_column_sum_f32")
    return out


def new_rng(device) -> torch.Tensor:
    """Two 64-bit words of key material for one layer's dropout masks, in device memory."""
    return torch.randint(-2 ** 63, 2 ** 63 - 1, (2,), dtype=torch.int64, device=device)


def dropout_mask(rng: torch.Tensor, salt: int, p: float, shape) -> torch.Tensor:
    """The keep mask (bool, ``shape``) the kernels derive from ``(rng, salt, p)`` for a contiguous tensor of that shape."""
    n = 1
    for d in shape:
        n *= int(d)
    keep = torch.empty(n, dtype=torch.uint8, device=rng.device)
    with _lib.on_device(rng.device):
        rc = _lib.lib().msda_dropout_mask_u8(rng.data_ptr(), int(salt), float(p), n, keep.data_ptr(), _stream(rng))
    _lib.check(rc, "msda_dropout_mask_u8")
    return keep.view(*shape).bool()


class _BiasResidualLayerNorm(Function):
    @staticmethod
    def forward(ctx, x, bias, residual, gamma, beta, eps, rng=None, salt=0, p=0.0):
        C = x.shape[-1]
        x2, r2 = x.contiguous().view(-1, C), residual.contiguous().view(-1, C)
        rows = x2.shape[0]
        z, y = torch.empty_like(x2), torch.empty_like(x2)
        mean = torch.empty(rows, dtype=torch.float32, device=x.device)
        rstd = torch.empty_like(mean)
        drop = rng is not None and p > 0.0
        with _lib.on_device(x.device):
            if drop:
                rc = _lib.lib().msda_epilogue_ln_dropout_forward_f32(
                    x2.data_ptr(), None if bias is None else bias.data_ptr(), r2.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                    float(eps), rows, C, rng.data_ptr(), int(salt), float(p), z.data_ptr(), y.data_ptr(), mean.data_ptr(),
                    rstd.data_ptr(), _stream(x))
            else:
                rc = _lib.lib().msda_epilogue_ln_forward_f32(
                    x2.data_ptr(), None if bias is None else bias.data_ptr(), r2.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                    float(eps), rows, C, z.data_ptr(), y.data_ptr(), mean.data_ptr(), rstd.data_ptr(), _stream(x))
        _lib.check(rc, "msda_epilogue_ln_forward_f32")
        ctx.save_for_backward(z, mean, rstd, gamma, *([rng] if drop else []))
        ctx.drop = (int(salt), float(p)) if drop else None
        ctx.has_bias = bias is not None
        ctx.shape = x.shape
        return y.view(x.shape)

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        z, mean, rstd, gamma = ctx.saved_tensors[:4]
        rows, C = z.shape
        dy2 = dy.contiguous().view(rows, C)
        dz = torch.empty_like(z)
        dgamma, dbeta = torch.empty_like(gamma), torch.empty_like(gamma)
        dbias = torch.empty_like(gamma) if ctx.has_bias else None
        dx = dz
        with _lib.on_device(z.device):
            if ctx.drop is not None:
                rng, (salt, p) = ctx.saved_tensors[4], ctx.drop
                dx = torch.empty_like(z)
                rc = _lib.lib().msda_epilogue_ln_dropout_backward_f32(
                    dy2.data_ptr(), z.data_ptr(), mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(), rows, C, rng.data_ptr(),
                    salt, p, dz.data_ptr(), dx.data_ptr(), dgamma.data_ptr(), dbeta.data_ptr(),
                    None if dbias is None else dbias.data_ptr(), _stream(z))
            else:
                rc = _lib.lib().msda_epilogue_ln_backward_f32(
                    dy2.data_ptr(), z.data_ptr(), mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(), rows, C, dz.data_ptr(),
                    dgamma.data_ptr(), dbeta.data_ptr(), None if dbias is None else dbias.data_ptr(), _stream(z))
        _lib.check(rc, "msda_epilogue_ln_backward_f32")
        return dx.view(ctx.shape), dbias, dz.view(ctx.shape), dgamma, dbeta, None, None, None, None


def bias_residual_layer_norm(x, bias, residual, gamma, beta, eps=1e-5, rng=None, salt=0, p=0.0):
    """LayerNorm over the last dim of ``residual + dropout(x + bias)``: one kernel forward, one backward (which also
    yields the gradients of bias, gamma and beta).  Dropout is applied when ``rng`` (see ``new_rng``) is given and
    ``p > 0``; ``salt`` names the call site."""
    if supported(x, bias, residual, gamma, beta) and x.shape[-1] in LN_CHANNELS and x.shape == residual.shape:
        return _BiasResidualLayerNorm.apply(x, bias, residual, gamma, beta, eps, rng, salt, p)
    _strict.note_fallback("bias_residual_layer_norm", "needs fp32 CUDA tensors outside autocast, channels in %r, x.shape == residual.shape" % (LN_CHANNELS,))
    y = x if bias is None else x + bias
    if rng is not None and p > 0.0:
        y = F.dropout(y, p, True)
    return F.layer_norm(residual + y, (x.shape[-1],), gamma, beta, eps)


class _Linear(Function):
    """F.linear whose bias gradient is one streaming column-sum kernel (cuBLAS for the three GEMMs)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        ctx.save_for_backward(x, weight)
        return F.linear(x, weight, bias)

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        x, weight = ctx.saved_tensors
        g2 = g.contiguous().view(-1, g.shape[-1])
        gx = (g2 @ weight).view(x.shape) if ctx.needs_input_grad[0] else None
        gw = g2.t() @ x.reshape(-1, x.shape[-1]) if ctx.needs_input_grad[1] else None
        gb = column_sum(g2) if ctx.needs_input_grad[2] else None
        return gx, gw, gb


def linear(x, weight, bias):
    if bias is not None and supported(x, weight, bias) and weight.shape[0] % 4 == 0:
        return _Linear.apply(x, weight, bias)
    _strict.note_fallback("linear", "needs a bias, fp32 CUDA tensors outside autocast and out_features % 4 == 0")
    return F.linear(x, weight, bias)


class _LinearReLU(Function):
    """relu(F.linear(x, W, b)): bias + ReLU in the GEMM epilogue forward (cuBLASLt), ReLU gradient and bias gradient in one
    streaming kernel backward."""

    @staticmethod
    def forward(ctx, x, weight, bias, rng=None, salt=0, p=0.0):
        x2 = x.reshape(-1, x.shape[-1])
        h = torch._addmm_activation(bias, x2, weight.t(), use_gelu=False)
        ctx.p = 0.0
        if rng is not None and p > 0.0:
            # dropout in place on the ReLU output: what is saved is dropout(relu(.)), whose sign pattern is the ReLU
            # mask AND the keep mask, so the backward needs neither the mask nor the generator
            with _lib.on_device(x.device):
                rc = _lib.lib().msda_dropout_inplace_f32(h.data_ptr(), h.numel(), rng.data_ptr(), int(salt), float(p), _stream(x))
            _lib.check(rc, "msda_dropout_inplace_f32")
            ctx.p = float(p)
        ctx.save_for_backward(x2, weight, h)
        ctx.xshape = x.shape
        return h.view(*x.shape[:-1], weight.shape[0])

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        x2, weight, h = ctx.saved_tensors
        g2 = g.contiguous().view(-1, g.shape[-1])
        rows, C = g2.shape
        dpre = torch.empty_like(g2)
        gb = torch.empty(C, dtype=torch.float32, device=g.device)
        with _lib.on_device(g.device):
            if ctx.p > 0.0:
                rc = _lib.lib().msda_relu_dropout_backward_column_sum_f32(g2.data_ptr(), h.data_ptr(), ctx.p, rows, C,
                                                                          dpre.data_ptr(), gb.data_ptr(), _stream(g))
            else:
                rc = _lib.lib().msda_relu_backward_column_sum_f32(g2.data_ptr(), h.data_ptr(), rows, C, dpre.data_ptr(),
                                                                  gb.data_ptr(), _stream(g))
        _lib.check(rc, "msda_relu_backward_column_sum_f32")
        gx = (dpre @ weight).view(ctx.xshape) if ctx.needs_input_grad[0] else None
        gw = dpre.t() @ x2 if ctx.needs_input_grad[1] else None
        return gx, gw, gb, None, None, None


def linear_relu(x, weight, bias, rng=None, salt=0, p=0.0):
    """``dropout(relu(F.linear(x, weight, bias)))`` (dropout when ``rng`` is given and ``p > 0``)."""
    if bias is not None and supported(x, weight, bias) and weight.shape[0] % 4 == 0:
        return _LinearReLU.apply(x, weight, bias, rng, salt, p)
    _strict.note_fallback("linear_relu", "needs a bias, fp32 CUDA tensors outside autocast and out_features % 4 == 0")
    h = F.relu(F.linear(x, weight, bias))
    return F.dropout(h, p, True) if rng is not None and p > 0.0 else h
