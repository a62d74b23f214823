"""The autograd boundary of the training step (SURVEY.md 8b "Autograd"): one ``torch.autograd.Function`` per CUDA stage, so
the reference's ``loss.backward()`` (``pipeline/train_val_utils.py:277``) lands gradients in the ``.grad`` of the registered
parameters.  Every forward and backward below is a kernel of libvbg_sm100a (tensors are fp32, channels-last); torch is the
tape, the allocator and the handful of residual adds / concatenations between stages.

The first brick, a linear layer on the pre-split tcgen05 GEMM (HF ``BertSelfOutput`` / ``BertIntermediate`` / ``BertOutput``,
the head MLPs, every 1x1 convolution):

    y = LinearPS.apply(x, weight, bias)          # x [M, K] fp32 CUDA, weight [N, K], bias [N]; N, K multiples of 64

backward:  dX = dY . W          (the forward GEMM over dY planes and the planes of W^T)
           dW = dY^T . X        (vbg_linear_wgrad: both plane operands fed to tcgen05 as MN-major tiles, the row range split
                                 over CTAs with a deterministic finish; N % 128 != 0 falls back to transposed planes + the
                                 forward GEMM)
           db = column sums of dY (fixed order)
All three are bf16x3 products with fp32 accumulation (fp32-class).  ``train_engine.py`` wires these into the training-mode
forward of ViBERTgridNet.
"""
from __future__ import annotations

import os

import torch

from . import ops


def _fp32():
    """VBG_PRECISION=fp32 (the exact CUDA-core mode of the eval engine): forward and data-gradient contractions run on the fp32
    SIMT kernels; weight gradients stay on the tensor cores (their error does not propagate).  Used to separate rounding noise
    from wiring errors when comparing a training step with the reference."""
    return os.environ.get("VBG_PRECISION", "").lower() == "fp32"


# ---------------------------------------------------------------------------------------------------------------------------
# Weight-gradient work on a SIDE STREAM of the captured training step.
#
# The backward's critical path is the data-gradient chain (dY -> dgrad GEMM -> elementwise backward -> next dgrad ...); the
# weight gradients, bias column sums and the stem / embedding-table gradients feed nothing but the gradient arena.  Inside the
# whole-step capture (train_engine.TrainEngine._graphed_loss) they are therefore launched on a second stream right after the
# layer's data gradient, and the main stream joins that stream ONCE, after ``torch.autograd.grad`` has returned: in the CUDA
# graph they become branches that run beside the following layers' bandwidth-bound elementwise kernels (a persistent tcgen05
# CTA leaves the SM's LSU / most of its threads idle).  Safety rules:
#   * only when the gradient's consumer launches no kernel before the join: the parameter is a LEAF of the tape, or a tensor
#     marked by ``pack_rows`` / ``leaf_view`` below (their backward is pure view arithmetic);
#   * every main-stream tensor the side kernels read (dY planes, saved activation planes, dY) is kept referenced until the
#     join, so the caching allocator cannot hand its memory to a later main-stream kernel;
#   * outside a capture (eager steps, DDP, tests of single Functions) ``SIDE`` is None and everything runs in line.
class _SideWork:
    def __init__(self, device, stream=None):
        self.stream = stream if stream is not None else torch.cuda.Stream(device)
        self.keep = []
        self.launched = 0


SIDE = None


def side_begin(device, stream=None):
    global SIDE
    SIDE = _SideWork(device, stream)
    return SIDE


def side_join():
    """Main stream waits for everything deferred so far; releases the kept tensors.  Idempotent."""
    global SIDE
    s, SIDE = SIDE, None
    if s is not None and s.launched:
        torch.cuda.current_stream().wait_stream(s.stream)
    if s is not None:
        s.keep.clear()


def _can_defer(*tensors):
    """Evaluated in a Function's forward (the side stream itself exists only while the captured backward runs)."""
    return all(t is None or t.is_leaf or getattr(t, "_vbg_defer_ok", False) for t in tensors)


def _deferred(fn, ok, *keep):
    """``fn()`` on the side stream (after everything enqueued on the current stream so far) when ``ok``, else in line."""
    s = SIDE
    if s is None or not ok:
        return fn()
    s.stream.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s.stream):
        out = fn()
    s.launched += 1
    s.keep.append((keep, out))
    return out


# Parameter-only preparation (bf16 planes of a weight, its transposed planes for the data gradient, conv weight repacks) on a
# PREP STREAM of the captured training step: these kernels read nothing but parameters, so in the CUDA graph they form a chain
# that depends on no activation and runs ahead of the forward, beside its tensor-core kernels; every consumer waits for its own
# layer's event.  The backward's weight forms are produced in the forward call too, which takes them off the backward's
# critical path.  Results are kept referenced until ``prep_end`` (their memory must not be recycled on the prep stream while a
# main-stream kernel still reads them).  Eligible: leaves of the tape and tensors marked by ``pack_rows`` / ``leaf_view``.
class _PrepWork:
    def __init__(self, device, stream=None):
        self.stream = stream if stream is not None else torch.cuda.Stream(device)
        self.keep = []


PREP = None


def prep_begin(device, stream=None):
    global PREP
    PREP = _PrepWork(device, stream)
    PREP.stream.wait_stream(torch.cuda.current_stream())          # the prep stream joins the capture here, once
    return PREP


def prep_end():
    global PREP
    p, PREP = PREP, None
    if p is not None:
        torch.cuda.current_stream().wait_stream(p.stream)
        p.keep.clear()


def _prepped(fn, ok=True):
    """``fn()`` (kernels over parameters only) on the prep stream; the current stream waits for exactly that work."""
    p = PREP
    if p is None or not ok:
        return fn()
    with torch.cuda.stream(p.stream):
        out = fn()
        ev = torch.cuda.Event()
        ev.record(p.stream)
    torch.cuda.current_stream().wait_event(ev)
    p.keep.append(out)
    return out


class _PackRowsF(torch.autograd.Function):
    """``torch.cat(parts, 0)`` whose backward hands every part a VIEW of the incoming gradient (no kernel): the packed QKV
    weight / bias of a BERT layer stays eligible for the deferred weight gradient."""

    @staticmethod
    def forward(ctx, *parts):
        ctx.rows = [int(p.shape[0]) for p in parts]
        return _prepped(lambda: torch.cat([p.detach() for p in parts], 0), all(p.is_leaf for p in parts))

    @staticmethod
    def backward(ctx, g):
        out, r0 = [], 0
        for r in ctx.rows:
            out.append(g[r0:r0 + r])
            r0 += r
        return tuple(out)


def pack_rows(*parts):
    t = _PackRowsF.apply(*parts)
    t._vbg_defer_ok = all(p.is_leaf for p in parts)
    return t


def leaf_view(p, *shape):
    """``p.view(shape)`` of a leaf parameter, marked: the view's backward is metadata only."""
    t = p.view(*shape)
    t._vbg_defer_ok = bool(p.is_leaf)
    return t


# ---------------------------------------------------------------------------------------------------------------------------
# The PLANES PROTOCOL: between two tensor-core stages an activation (and its gradient) may travel through the tape as the bf16
# ``[2, *shape]`` tensor of a ``Split`` -- hi / lo planes, the operand format of the bf16x3 kernels -- instead of fp32.  A stage
# that accepts it skips its ``to_split`` pass; a stage that emits it writes the planes from its epilogue / elementwise kernel
# and never writes the fp32 tensor.  The gradient of a planes tensor is a planes tensor (same shape and dtype, as autograd
# requires), so the backward chain keeps the format too.  Used for the QKV -> attention and FFN-up -> GELU -> FFN-down chains
# of a BERT layer (the [rows, 2304] / [rows, 3072] tensors, the largest elementwise traffic of the layer).
def _is_planes(t):
    return t is not None and t.dtype == torch.bfloat16 and t.dim() >= 2 and t.shape[0] == 2


class LinearPS(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, out_planes=False):
        ctx.x_planes = _is_planes(x)
        ctx.out_planes = bool(out_planes)
        if (ctx.x_planes or ctx.out_planes) and _fp32():
            raise ValueError("LinearPS: the planes protocol belongs to the tensor-core mode")
        xshape = tuple(x.shape[1:]) if ctx.x_planes else tuple(x.shape)
        if len(xshape) != 2 or weight.shape[1] != xshape[1] or weight.shape[0] % 64 or weight.shape[1] % 64:
            raise ValueError("LinearPS: x [M, K], weight [N, K] with N and K multiples of 64")
        ep = ops.make_epilogue(None, None if bias is None else bias.detach())
        ctx.wt = None
        if _fp32():
            w = weight.detach().contiguous()
            y = ops.gemm(x.detach().contiguous(), w, ep=ep, precision=ops.PREC_FP32)
            ctx.save_for_backward(x, weight)
        else:
            param_only = _can_defer(weight) or (weight._base is not None and _can_defer(weight._base))

            def weight_forms():          # forward planes and, when the input needs a gradient, the [K, N] planes of the data gradient
                w_ = weight.detach().contiguous()
                return w_, ops.split_bf16(w_), (ops.transpose_split(w_) if PREP is not None and x.requires_grad else None)
            w, ws, ctx.wt = _prepped(weight_forms, param_only)
            xs = ops.Split(x.detach()) if ctx.x_planes else ops.to_split(x.detach().contiguous())
            y = ops.gemm(xs, w, ep=ep, precision=ops.PREC_BF16X3, W_split=ws, split_out=ctx.out_planes)
            if ctx.out_planes:
                y = y.t
            ctx.save_for_backward(xs.t, weight)          # the planes are what the weight-gradient kernel reads (same bytes as x)
        ctx.planes = not _fp32()
        ctx.has_bias = bias is not None
        ctx.defer = _can_defer(weight, bias)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        xs = ops.Split(x) if ctx.planes else ops.to_split(x.detach().contiguous())
        M, K = xs.shape
        N = weight.shape[0]
        dy = dy.contiguous()
        dys = ops.Split(dy) if ctx.out_planes else ops.to_split(dy)
        dy_sum = dys if ctx.out_planes else dy                                       # the column sums read either format
        dx = dw = db = None
        if ctx.needs_input_grad[0] and _fp32():
            dx = ops.gemm(dy, weight.detach().t().contiguous(), precision=ops.PREC_FP32)
        elif ctx.needs_input_grad[0]:
            wt = ctx.wt if ctx.wt is not None else ops.transpose_split(weight.detach().contiguous())     # [K, N] planes
            dx = ops.gemm(dys, weight.detach().t(), precision=ops.PREC_BF16X3, W_split=wt.t, N=K, K=N, ldw=N, split_out=ctx.x_planes)
            if ctx.x_planes:
                dx = dx.t
        want_w, want_b = ctx.needs_input_grad[1], ctx.has_bias and ctx.needs_input_grad[2]

        def param_grads():
            # MN-major tcgen05 operands straight from the row-major planes (no transposes)
            return (ops.linear_wgrad(dys, xs) if want_w else None), (ops.colsum(dy_sum) if want_b else None)
        dw, db = _deferred(param_grads, ctx.defer, dys, xs, dy)
        return dx, dw, db, None


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


def _c(t):
    return t if t.is_contiguous() else t.contiguous()


class LinearSmall(torch.autograd.Function):
    """Linear layers the tensor-core kernels do not take (output width not a multiple of 64: the 2 / C-way heads, the packed
    1x1 segmentation heads -- 3 + C columns, 26 for the reference's 23-tag EPHOIE BIO set, the 25 CRF emissions): CUDA-core
    GEMM forward and data gradient; the weight gradient runs ``vbg_small_wgrad`` over column slices of 16 outputs (the
    kernel's register tile), each slice a strided view of dY."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        x = _c(x.detach())
        y = ops.gemm(x, _c(weight.detach()), ep=ops.make_epilogue(None, None if bias is None else bias.detach()),
                     precision=ops.PREC_FP32)
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        dy = _c(dy)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = ops.gemm(dy, _c(weight.detach().t()), precision=ops.PREC_FP32)
        if ctx.needs_input_grad[1]:
            N = dy.shape[1]
            dw = ops.small_wgrad(dy, x) if N <= 16 else torch.cat([ops.small_wgrad(dy[:, i:i + 16], x) for i in range(0, N, 16)], 0)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = ops.colsum(dy)
        return dx, dw, db


def linear(x, weight, bias=None, out_planes=False):
    """[M, K] x [N, K]^T (+ bias) on whichever kernel takes the shape.  ``out_planes``: emit the result in the plane format
    (tensor-core kernel only; see "planes protocol")."""
    if weight.shape[0] % 64 == 0 and weight.shape[1] % 64 == 0:
        return LinearPS.apply(x, weight, bias, out_planes)
    if out_planes or _is_planes(x):
        raise ValueError("linear: the plane format needs output / input widths that are multiples of 64")
    return LinearSmall.apply(x, weight, bias)


class ConvPS(torch.autograd.Function):
    """NHWC convolution (OIHW parameter, optional bias), stride 1 or 2, on the pre-split implicit-GEMM kernels:
    forward ``vbg_conv2d_ps``; data gradient = the stride-1 convolution of dY (zero-inserted for stride 2) with the flipped /
    transposed weight; weight gradient ``vbg_conv2d_wgrad``.  Replaces ``nn.Conv2d`` forward/backward of the ResNet / FPN /
    head convolutions (model/ResNetFPN_ViBERTgrid.py:106-186)."""

    @staticmethod
    def forward(ctx, x, w_oihw, bias, stride, pad):
        w = w_oihw.detach()
        Cout, Cin, kh, kw = w.shape

        def weight_forms():              # OHWI repack, its planes and, when the input needs a gradient, the data-gradient planes
            w_ohwi_ = ops.repack_oihw_to_ohwi(_c(w)) if kh * kw > 1 else _c(w.reshape(Cout, 1, 1, Cin))
            fp = _fp32()
            return (w_ohwi_, None if fp else ops.split_bf16(w_ohwi_),
                    ops.conv_dgrad_weight(w_ohwi_) if (PREP is not None and not fp and x.requires_grad) else None)
        w_ohwi, ws, ctx.wd = _prepped(weight_forms, _can_defer(w_oihw))
        xs = ops.to_split(_c(x.detach()))
        ep = ops.make_epilogue(None, bias.detach()) if bias is not None else None
        if _fp32():
            y = ops.conv2d(_c(x.detach()), w_ohwi, stride, pad, ep=ep, precision=ops.PREC_FP32)
        else:
            y = ops.conv2d(xs, w_ohwi, stride, pad, ep=ep, precision=ops.PREC_BF16X3, W_split=ws)
        ctx.save_for_backward(xs.t, w_ohwi)
        ctx.cfg = (stride, pad, bias is not None)
        ctx.defer = _can_defer(w_oihw, bias)
        return y

    @staticmethod
    def backward(ctx, dy):
        xs_t, w_ohwi = ctx.saved_tensors
        stride, pad, has_bias = ctx.cfg
        xs = ops.Split(xs_t)
        B, H, W, Cin = xs.shape
        Cout, kh, kw, _ = w_ohwi.shape
        dy = _c(dy)
        _, Ho, Wo, _ = dy.shape
        dys = ops.to_split(dy)
        dx = dw = db = None
        if ctx.needs_input_grad[0] and _fp32():
            wf = w_ohwi.flip(1, 2).permute(3, 1, 2, 0).contiguous()         # fp32 twin of conv_dgrad_weight
            src = dy if stride == 1 else ops.expand2x(dy, H - kh + 1 + 2 * pad, W - kw + 1 + 2 * pad, 1.0, zero_insert=True)
            dx = ops.conv2d(src, wf, 1, kh - 1 - pad, precision=ops.PREC_FP32)
        elif ctx.needs_input_grad[0]:
            wd = ctx.wd if ctx.wd is not None else ops.conv_dgrad_weight(w_ohwi)      # planes [2, Cin, kh, kw, Cout]
            if stride == 1:
                dx = ops.conv2d(dys, wd[0], 1, kh - 1 - pad, precision=ops.PREC_BF16X3, W_split=wd)
            elif kh == 1 and kw == 1 and pad == 0:                          # 1x1 / 2: a GEMM on the coarse lattice, then spread
                d = ops.gemm(dys.view(B * Ho * Wo, Cout), wd[0].view(Cin, Cout), precision=ops.PREC_BF16X3,
                             W_split=wd.view(2, Cin, Cout))
                dx = ops.expand2x(d.view(B, Ho, Wo, Cin), H, W, 1.0, zero_insert=True)
            else:                                                           # dY onto the stride-1 lattice, then a stride-1 conv
                dyz = ops.expand2x(dy, H - kh + 1 + 2 * pad, W - kw + 1 + 2 * pad, 1.0, zero_insert=True)
                dx = ops.conv2d(ops.to_split(dyz), wd[0], 1, kh - 1 - pad, precision=ops.PREC_BF16X3, W_split=wd)
        want_w, want_b = ctx.needs_input_grad[1], has_bias and ctx.needs_input_grad[2]

        def param_grads():
            return ((ops.conv2d_wgrad(dys, xs, kh, kw, stride, pad).permute(0, 3, 1, 2).contiguous() if want_w else None),
                    (ops.colsum(dy.view(-1, Cout)) if want_b else None))
        dw, db = _deferred(param_grads, ctx.defer, dys, xs, dy)
        return dx, dw, db, None, None


class StemF(torch.autograd.Function):
    """7x7/2 stem over the zero-bordered NHWC4 batch (no data gradient: the input is the image)."""

    @staticmethod
    def forward(ctx, x4, w_oihw):
        w774, w256 = ops.stem_pack_weights(w_oihw.detach())
        if _fp32():
            y = ops.stem_conv(x4, w774, precision=ops.PREC_FP32)
        else:
            y = ops.stem_conv(x4, w774, precision=ops.PREC_BF16X3, W_split=ops.split_bf16(w256))
        ctx.save_for_backward(x4)
        ctx.defer = _can_defer(w_oihw)
        return y

    @staticmethod
    def backward(ctx, dy):
        (x4,) = ctx.saved_tensors
        dy = _c(dy)
        dw = _deferred(lambda: ops.stem_wgrad(x4, dy)[..., :3].permute(0, 3, 1, 2).contiguous(), ctx.defer, x4, dy)   # via [64, 7, 7, 4]
        return None, dw


def sync_batch_stats(mean, var, rows, eps, group):
    """SyncBatchNorm statistics: this rank's (mean, biased var) over ``rows`` rows -> the statistics of all ranks' rows together
    (count-weighted, combined in float64 on the device: no host synchronisation).  Returns (mean, var, rstd, total rows as a
    0-d float64 device tensor).  What torch's ``batch_norm_gather_stats_with_counts`` does for the reference when
    train_SROIE.py:203-205 converts the model."""
    import torch.distributed as dist
    Cc = mean.numel()
    local = torch.cat([mean, var, torch.full((1,), float(rows), dtype=torch.float32, device=mean.device)])
    gathered = [torch.empty_like(local) for _ in range(dist.get_world_size(group))]
    dist.all_gather(gathered, local, group=group)
    g = torch.stack(gathered).double()
    cnt = g[:, -1:]
    total = cnt.sum()
    m = (g[:, :Cc] * cnt).sum(0) / total
    v = ((g[:, Cc:2 * Cc] + g[:, :Cc] ** 2) * cnt).sum(0) / total - m * m
    v = v.clamp_min(0.0)
    return m.float().contiguous(), v.float().contiguous(), torch.rsqrt(v + eps).float().contiguous(), total


class BatchNormTrainF(torch.autograd.Function):
    """nn.BatchNorm2d in train mode over NHWC, with the residual add and ReLU that follow it in the ResNet blocks fused.
    ``stats`` (a list) receives (mean, biased var, rows) so the caller can update the running statistics.
    ``sync`` = None (per-rank statistics) or a 1-tuple ``(process_group,)`` (None inside = the default group): the
    nn.SyncBatchNorm form -- statistics over all ranks' rows in the forward, the two per-channel sums of the backward
    all-reduced before dx is formed; the parameter gradients stay per-rank (the gradient average adds them up)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, residual, relu, eps, stats, sync=None):
        Cc = x.shape[-1]
        x2 = _c(x.detach()).view(-1, Cc)
        mean, var, rstd = ops.bn_stats(x2, eps)
        rows = x2.shape[0]
        if sync is not None:
            mean, var, rstd, rows = sync_batch_stats(mean, var, rows, eps, sync[0])
        g = _c(gamma.detach())
        y = ops.bn_apply(x2, mean, rstd, g, _c(beta.detach()), None if residual is None else _c(residual.detach()).view(-1, Cc), relu)
        stats.append((mean, var, rows))
        ctx.save_for_backward(x2, y if relu else None, mean, rstd, g, rows if sync is not None else None)
        ctx.has_res = residual is not None
        ctx.sync = sync
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, dy):
        x2, y, mean, rstd, g, total = ctx.saved_tensors
        Cc = x2.shape[1]
        dy2 = _c(dy).view(-1, Cc)
        if ctx.sync is None:
            dx, dres, dg, db = ops.bn_bwd(x2, dy2, y, mean, rstd, g, want_dres=ctx.has_res)
        else:
            import torch.distributed as dist
            dg, db = ops.bn_bwd_reduce(x2, dy2, y, mean, rstd)
            sums = torch.stack([dg, db])
            dist.all_reduce(sums, group=ctx.sync[0])
            sums = (sums.double() / total).float()            # pre-divided by the global row count: inv_count = 1 below
            dx, dres = ops.bn_bwd_dx(x2, dy2, y, mean, rstd, g, sums[0].contiguous(), sums[1].contiguous(), 1,
                                     want_dres=ctx.has_res)
        return dx.view(dy.shape), dg, db, (dres.view(dy.shape) if dres is not None else None), None, None, None, None


class MaxPoolF(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        x = _c(x.detach())
        y = ops.maxpool3x3s2(x)
        ctx.save_for_backward(x, y)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y = ctx.saved_tensors
        return ops.maxpool3x3s2_bwd(x, _c(dy), y)


class AvgPoolF(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        ctx.hw = (x.shape[1], x.shape[2])
        return ops.avgpool2x2(_c(x.detach()))

    @staticmethod
    def backward(ctx, dy):
        return ops.expand2x(_c(dy), ctx.hw[0], ctx.hw[1], 0.25)


class Up2F(torch.autograd.Function):
    """Nearest x2 (the FPN top-down path); backward = 2x2 block sums."""

    @staticmethod
    def forward(ctx, x):
        return ops.expand2x(_c(x.detach()), 2 * x.shape[1], 2 * x.shape[2], 1.0)

    @staticmethod
    def backward(ctx, dy):
        return ops.sumpool2x2(_c(dy), 1.0)


class GeluF(torch.autograd.Function):
    """erf-GELU; an input in the plane format stays in it (output, saved activation and both gradients): nothing fp32 of size
    [rows, 3072] is written in either direction."""

    @staticmethod
    def forward(ctx, x):
        x = _c(x.detach())
        ctx.save_for_backward(x)
        ctx.planes = _is_planes(x)
        return ops.gelu(ops.Split(x), split_out=True).t if ctx.planes else ops.gelu(x)

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        if ctx.planes:
            return ops.gelu(ops.Split(x), ops.Split(_c(dy)), split_out=True).t
        return ops.gelu(x, _c(dy))


class DropoutF(torch.autograd.Function):
    """``step_seed``: optional int64[1] DEVICE tensor folded into the seed at run time (see vbg_dropout_ds): under CUDA-graph
    replay the by-value seed is baked, the device word changes every step."""

    @staticmethod
    def forward(ctx, x, p, seed, step_seed=None):
        ctx.ps = (p, seed, step_seed)
        return ops.dropout(_c(x.detach()), p, seed, step_seed)

    @staticmethod
    def backward(ctx, dy):
        return ops.dropout(_c(dy), *ctx.ps), None, None, None


class AttentionF(torch.autograd.Function):
    """Self-attention over the packed varlen batch, both directions on the tcgen05 kernels: forward
    ``vbg_attention_split_train_fwd`` (dropout on the attention probabilities inside the kernel, row log-sum-exp kept),
    backward ``vbg_attention_bwd_tc`` (probabilities rebuilt from the log-sum-exp, the same counter-based mask regenerated).
    ``VBG_PRECISION=fp32`` selects the CUDA-core pair (``vbg_attention_fwd`` / ``vbg_attention_bwd``), which has no
    attention dropout: that mode exists to prove the wiring against the reference's gradients with dropout off."""

    @staticmethod
    def forward(ctx, qkv, cu, nseq, max_len, heads, p_drop=0.0, seed=0, step_seed=None):
        qkv = _c(qkv.detach())
        hid = qkv.shape[-1] // 3
        if hid // heads != 64:
            raise NotImplementedError("training attention: head dimension 64 only")
        ctx.tc = max_len <= 512 and ops.tc_available() and not _fp32()
        ctx.cfg = (nseq, max_len, heads)
        ctx.planes = _is_planes(qkv)                  # packed QKV straight from the GEMM's epilogue in the plane format
        if ctx.planes and not ctx.tc:
            raise ValueError("AttentionF: a QKV tensor in the plane format needs the tensor-core attention kernels")
        if ctx.tc:
            qs = ops.Split(qkv) if ctx.planes else ops.to_split(qkv)
            out, lse2 = ops.attention_split_train(qs, cu, nseq, max_len, heads, p_drop, seed, step_seed)
            ctx.save_for_backward(qs.t, out, lse2, cu)
            ctx.drop = (float(p_drop), int(seed), step_seed)
        else:
            if p_drop > 0.0 and not getattr(AttentionF, "_warned", False):
                AttentionF._warned = True
                import warnings
                warnings.warn("[vibertgrid_b200] the fp32 CUDA-core attention path applies no attention-probability dropout")
            out = ops.attention(qkv, cu, nseq, max_len, heads, ops.PREC_FP32)
            ctx.save_for_backward(qkv, out, cu)
        return out

    @staticmethod
    def backward(ctx, d_out):
        if ctx.tc:
            qs, out, lse2, cu = ctx.saved_tensors
            dqkv = ops.attention_bwd_tc(ops.Split(qs), out, _c(d_out), lse2, cu, *ctx.cfg, *ctx.drop)
            if ctx.planes:                               # the gradient of a planes tensor is a planes tensor
                dqkv = ops.to_split(dqkv).t
        else:
            qkv, out, cu = ctx.saved_tensors
            dqkv = ops.attention_bwd(qkv, out, _c(d_out), cu, *ctx.cfg)
        return dqkv, None, None, None, None, None, None, None


class EmbedSumF(torch.autograd.Function):
    """word[ids] + position[pos] + token_type[0] (HF BertEmbeddings before its LayerNorm).  The forward is a row gather; the
    backward scatter-adds into the tables (vbg_embed_bwd) and column-sums into token_type row 0."""

    @staticmethod
    def forward(ctx, word, position, type_emb, ids, pos):
        ctx.save_for_backward(ids, pos)
        ctx.shapes = (word.shape[0], position.shape[0], type_emb.shape)
        ctx.defer = _can_defer(word, position, type_emb)
        idl, pol = ids.long(), pos.long()
        return word.detach().index_select(0, idl) + position.detach().index_select(0, pol) + type_emb.detach()[0]

    @staticmethod
    def backward(ctx, dx):
        ids, pos = ctx.saved_tensors
        V, Pm, tshape = ctx.shapes
        dx = _c(dx)

        def tables():
            dword, dpos = ops.embed_bwd(dx, ids, pos, V, Pm)
            dtype = torch.zeros(tshape, dtype=torch.float32, device=dx.device)
            dtype[0] = ops.colsum(dx)
            return dword, dpos, dtype
        dword, dpos, dtype = _deferred(tables, ctx.defer, dx, ids, pos)
        return dword, dpos, dtype, None, None


class SegmentReduceF(torch.autograd.Function):
    @staticmethod
    def forward(ctx, hidden, tok_row, seg_start, K, mode):
        ctx.save_for_backward(tok_row, seg_start)
        ctx.cfg = (hidden.shape[0], mode)
        return ops.segment_reduce(_c(hidden.detach()), tok_row, seg_start, K, mode)

    @staticmethod
    def backward(ctx, dseg):
        tok_row, seg_start = ctx.saved_tensors
        return ops.segment_reduce_bwd(_c(dseg), tok_row, seg_start, *ctx.cfg), None, None, None, None


class GridScatterF(torch.autograd.Function):
    @staticmethod
    def forward(ctx, seg_emb, idx, boxes, seg_off, B, stride):
        ctx.save_for_backward(idx, boxes, seg_off)
        ctx.cfg = (B, seg_emb.shape[0], stride, seg_emb.shape[1])
        return ops.grid_scatter(_c(seg_emb.detach()), idx, seg_off)

    @staticmethod
    def backward(ctx, dgrid):
        idx, boxes, seg_off = ctx.saved_tensors
        B, K, stride, Cc = ctx.cfg
        dgrid = _c(dgrid)
        return ops.grid_scatter_bwd(dgrid.view(-1, Cc), Cc, idx, boxes, seg_off, B, K, stride, Cc), None, None, None, None, None


class RoiAlignF(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat, boxes, seg_off, scale, P):
        ctx.save_for_backward(boxes, seg_off)
        ctx.cfg = (feat.shape[0], feat.shape[1], feat.shape[2], scale)
        return ops.roi_align(_c(feat.detach()), boxes, seg_off, scale, P)

    @staticmethod
    def backward(ctx, dout):
        boxes, seg_off = ctx.saved_tensors
        return ops.roi_align_bwd(_c(dout), boxes, seg_off, *ctx.cfg), None, None, None, None


class SegCEF(torch.autograd.Function):
    """The two mean cross entropies of the auxiliary segmentation head (semantic_segmentation_head.py:343-347, default
    reduction) from the LOW-resolution logits: forward vbg_seg_ce_loss (labels painted in registers), backward
    vbg_seg_ce_bwd against the painted label maps.  Returns a [2] tensor (mask-head CE, class-head CE)."""

    @staticmethod
    def forward(ctx, logits, boxes, seg_off, seg_cls, B, H, W, up, c_split):
        logits = _c(logits.detach())
        out = ops.seg_ce_loss(boxes, seg_off, seg_cls, logits, B, H, W, up, c_split)
        pos_neg, cls = ops.label_paint(boxes, seg_off, seg_cls, B, H, W)
        ctx.save_for_backward(logits, pos_neg, cls)
        ctx.cfg = (H, W, up, c_split)
        return out

    @staticmethod
    def backward(ctx, g):
        logits, pos_neg, cls = ctx.saved_tensors
        return (ops.seg_ce_bwd(logits, pos_neg, cls, *ctx.cfg, _c(g.float())),) + (None,) * 8


class CrfNllF(torch.autograd.Function):
    """Per-sample negative log-likelihood [B] of the linear-chain CRF (model/crf.py:148-152; the training branch of
    CRFFieldTypeClassification.forward, field_type_classification_head.py:686-699) from the emissions [K, T], the transition
    parameter [T, T] and the gold tags: vbg_crf_nll_fwd keeps the forward variables, vbg_crf_nll_bwd turns them into posterior
    marginals minus gold counts."""

    @staticmethod
    def forward(ctx, feats, trans, tags, seg_off, B):
        feats, trans = _c(feats.detach()), _c(trans.detach())
        nll, alpha, logz = ops.crf_nll_fwd(feats, trans, tags, seg_off, B)
        ctx.save_for_backward(feats, trans, tags, seg_off, alpha)
        ctx.B = B
        return nll

    @staticmethod
    def backward(ctx, g):
        feats, trans, tags, seg_off, alpha = ctx.saved_tensors
        dfeats, dtrans = ops.crf_nll_bwd(feats, trans, tags, seg_off, ctx.B, alpha, _c(g.float()))
        return dfeats, dtrans, None, None, None
