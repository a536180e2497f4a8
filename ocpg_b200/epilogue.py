"""Encoder-layer epilogue operators on the C ABI of include/msda_sm100.h.

The reference's DeformableTransformerEncoderLayer (models/deformable_transformer.py:243-260) surrounds the attention and
the FFN with ``norm(src + dropout(linear(x)))``.  In eager PyTorch the parameter gradients of that pattern -- LayerNorm's
gamma/beta kernel and one column-sum kernel per Linear bias -- take a third of a TF32 encoder step
(profiles/r1_encoder_kernel_breakdown.txt).  The autograd Functions below compute the same values with HBM-streaming
kernels; the GEMMs themselves stay with cuBLAS (``torch.matmul`` / ``F.linear``).

    bias_residual_layer_norm(x, bias, residual, gamma, beta, eps)   == F.layer_norm(residual + (x + bias), ...)
    linear(x, weight, bias)                                         == F.linear(x, weight, bias)
    linear_relu(x, weight, bias)                                    == F.relu(F.linear(x, weight, bias))

fp32 CUDA tensors only (the callers fall back to the torch ops otherwise); dropout must be inactive.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _lib

LN_CHANNELS = (128, 256, 512, 1024)


def _stream(t) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def supported(*tensors) -> bool:
    """fp32 CUDA tensors outside autocast: what the epilogue kernels take."""
    if torch.is_autocast_enabled():
        return False
    return all(t is None or (t.is_cuda and t.dtype == torch.float32) for t in tensors)


def column_sum(x2d: torch.Tensor) -> torch.Tensor:
    """sum over rows of a contiguous (rows, C) fp32 matrix, C % 4 == 0."""
    rows, C = x2d.shape
    out = torch.empty(C, dtype=torch.float32, device=x2d.device)
    with torch.cuda.device(x2d.device):
        rc = _lib.lib().msda_column_sum_f32(x2d.data_ptr(), rows, C, out.data_ptr(), _stream(x2d))
    _lib.check(rc, "msda_column_sum_f32")
    return out


class _BiasResidualLayerNorm(Function):
    @staticmethod
    def forward(ctx, x, bias, residual, gamma, beta, eps):
        C = x.shape[-1]
        x2, r2 = x.contiguous().view(-1, C), residual.contiguous().view(-1, C)
        rows = x2.shape[0]
        z, y = torch.empty_like(x2), torch.empty_like(x2)
        mean = torch.empty(rows, dtype=torch.float32, device=x.device)
        rstd = torch.empty_like(mean)
        with torch.cuda.device(x.device):
            rc = _lib.lib().msda_epilogue_ln_forward_f32(
                x2.data_ptr(), None if bias is None else bias.data_ptr(), r2.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                float(eps), rows, C, z.data_ptr(), y.data_ptr(), mean.data_ptr(), rstd.data_ptr(), _stream(x))
        _lib.check(rc, "msda_epilogue_ln_forward_f32")
        ctx.save_for_backward(z, mean, rstd, gamma)
        ctx.has_bias = bias is not None
        ctx.shape = x.shape
        return y.view(x.shape)

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        z, mean, rstd, gamma = ctx.saved_tensors
        rows, C = z.shape
        dy2 = dy.contiguous().view(rows, C)
        dz = torch.empty_like(z)
        dgamma, dbeta = torch.empty_like(gamma), torch.empty_like(gamma)
        dbias = torch.empty_like(gamma) if ctx.has_bias else None
        with torch.cuda.device(z.device):
            rc = _lib.lib().msda_epilogue_ln_backward_f32(
                dy2.data_ptr(), z.data_ptr(), mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(), rows, C, dz.data_ptr(),
                dgamma.data_ptr(), dbeta.data_ptr(), None if dbias is None else dbias.data_ptr(), _stream(z))
        _lib.check(rc, "msda_epilogue_ln_backward_f32")
        dz = dz.view(ctx.shape)
        return dz, dbias, dz, dgamma, dbeta, None


def bias_residual_layer_norm(x, bias, residual, gamma, beta, eps=1e-5):
    """LayerNorm over the last dim of ``residual + (x + bias)``: one kernel forward, one backward (which also yields the
    gradients of bias, gamma and beta)."""
    if supported(x, bias, residual, gamma, beta) and x.shape[-1] in LN_CHANNELS and x.shape == residual.shape:
        return _BiasResidualLayerNorm.apply(x, bias, residual, gamma, beta, eps)
    y = x if bias is None else x + bias
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
    return F.linear(x, weight, bias)


class _LinearReLU(Function):
    """relu(F.linear(x, W, b)): bias + ReLU in the GEMM epilogue forward (cuBLASLt), ReLU gradient and bias gradient in one
    streaming kernel backward."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        x2 = x.reshape(-1, x.shape[-1])
        h = torch._addmm_activation(bias, x2, weight.t(), use_gelu=False)
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
        with torch.cuda.device(g.device):
            rc = _lib.lib().msda_relu_backward_column_sum_f32(g2.data_ptr(), h.data_ptr(), rows, C, dpre.data_ptr(),
                                                              gb.data_ptr(), _stream(g))
        _lib.check(rc, "msda_relu_backward_column_sum_f32")
        gx = (dpre @ weight).view(ctx.xshape) if ctx.needs_input_grad[0] else None
        gw = dpre.t() @ x2 if ctx.needs_input_grad[1] else None
        return gx, gw, gb


def linear_relu(x, weight, bias):
    if bias is not None and supported(x, weight, bias) and weight.shape[0] % 4 == 0:
        return _LinearReLU.apply(x, weight, bias)
    return F.relu(F.linear(x, weight, bias))
