"""First brick of the training step: a linear layer whose forward AND backward run on the pre-split
tcgen05 GEMM -- the autograd boundary the reference's ``loss.backward()`` (``pipeline/train_val_utils.py:277``) will cross for
every ``nn.Linear`` on the path (HF ``BertSelfOutput`` / ``BertIntermediate`` / ``BertOutput``, the head MLPs).

    y = LinearPS.apply(x, weight, bias)          # x [M, K] fp32 CUDA, weight [N, K], bias [N]; N, K multiples of 64

backward:  dX = dY . W          (the forward GEMM over dY planes and the planes of W^T)
           dW = dY^T . X        (vbg_linear_wgrad: both plane operands fed to tcgen05 as MN-major tiles, the row range split
                                 over CTAs with a deterministic finish; N % 128 != 0 falls back to transposed planes + the
                                 forward GEMM)
           db = column sums of dY (fixed order)
All three are bf16x3 products with fp32 accumulation (fp32-class).  Not yet wired into ViBERTgridNet: the training-mode
forward of the whole module is the next milestone.
"""
from __future__ import annotations

import torch

from . import ops


class LinearPS(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias):
        if x.dim() != 2 or weight.shape[1] != x.shape[1] or weight.shape[0] % 64 or weight.shape[1] % 64:
            raise ValueError("LinearPS: x [M, K], weight [N, K] with N and K multiples of 64")
        xs = ops.to_split(x.detach().contiguous())
        w = weight.detach().contiguous()
        y = ops.gemm(xs, w, ep=ops.make_epilogue(None, None if bias is None else bias.detach()), precision=ops.PREC_BF16X3,
                     W_split=ops.split_bf16(w))
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        M, K = x.shape
        N = weight.shape[0]
        dy = dy.contiguous()
        dys = ops.to_split(dy)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            wt = ops.transpose_split(weight.detach().contiguous())                 # [K, N] planes
            dx = ops.gemm(dys, weight.detach().t(), precision=ops.PREC_BF16X3, W_split=wt.t, N=K, K=N, ldw=N)
        if ctx.needs_input_grad[1]:
            if N % 128 == 0:                       # MN-major tcgen05 operands straight from the row-major planes
                dw = ops.linear_wgrad(dys, ops.to_split(x.detach().contiguous()))
            else:                                  # transposed-operand route through the forward GEMM
                Mp = (M + 63) // 64 * 64
                dyt = ops.transpose_split(dys, Mp)                                 # [N, Mp] planes
                xt = ops.transpose_split(x.detach().contiguous(), Mp)              # [K, Mp] planes
                dw = ops.gemm(dyt, x.detach(), precision=ops.PREC_BF16X3, W_split=xt.t, N=K, K=Mp, ldw=Mp)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = ops.colsum(dy)
        return dx, dw, db


class LayerNormPS(torch.autograd.Function):
    """LayerNorm over the last dimension of a [R, hidden] fp32 tensor (vbg_layernorm / vbg_layernorm_bwd)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, eps):
        y = ops.layernorm(x.detach().contiguous(), gamma.detach(), beta.detach(), eps)
        ctx.save_for_backward(x, gamma)
        ctx.eps = eps
        return y

    @staticmethod
    def backward(ctx, dy):
        x, gamma = ctx.saved_tensors
        dx, dg, db = ops.layernorm_bwd(x.detach().contiguous(), dy.contiguous(), gamma.detach(), ctx.eps,
                                       want_params=ctx.needs_input_grad[1] or ctx.needs_input_grad[2])
        return dx, dg, db, None


class Conv2dS1PS(torch.autograd.Function):
    """Stride-1 NHWC convolution: forward and data gradient on the pre-split implicit-GEMM conv, weight gradient on the
    MN-major wgrad kernel (``vbg_conv2d_wgrad``: Cout % 128 == 0, Cin % 64 == 0)."""

    @staticmethod
    def forward(ctx, x_nhwc, w_ohwi, pad):
        xs = ops.to_split(x_nhwc.detach().contiguous())
        y = ops.conv2d(xs, w_ohwi.detach(), 1, pad, precision=ops.PREC_BF16X3, W_split=ops.split_bf16(w_ohwi.detach()))
        ctx.save_for_backward(x_nhwc, w_ohwi)
        ctx.pad = pad
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        Cout, kh, kw, Cin = w.shape
        dys = ops.to_split(dy.contiguous())
        dx = dw = None
        if ctx.needs_input_grad[0]:
            wd = ops.conv_dgrad_weight(w.detach())                          # planes [2, Cin, kh, kw, Cout]
            dx = ops.conv2d(dys, wd[0].float(), 1, kh - 1 - ctx.pad, precision=ops.PREC_BF16X3, W_split=wd)
        if ctx.needs_input_grad[1]:
            dw = ops.conv2d_wgrad(dys, ops.to_split(x.detach().contiguous()), kh, kw, 1, ctx.pad)
        return dx, dw, None
