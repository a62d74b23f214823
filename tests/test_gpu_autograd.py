"""Linear-layer forward + backward on the pre-split tcgen05 GEMM (vibertgrid_pytorch_b200.autograd.LinearPS) against
torch's float64 autograd of the same layer: y, dX, dW, db within fp32-class tolerance (bf16x3 products, fp32 accumulate)."""
import numpy as np
import pytest
import torch

from conftest import relerr

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M,N,K", [(4128, 768, 768), (516, 3072, 768), (1000, 768, 3072), (130, 64, 128)])
def test_linear_ps_forward_backward(M, N, K):
    from vibertgrid_pytorch_b200 import ops
    from vibertgrid_pytorch_b200.autograd import LinearPS
    assert ops.tc_available()
    g = torch.Generator().manual_seed(M + N + K)
    x = torch.randn(M, K, generator=g); w = torch.randn(N, K, generator=g) / K ** 0.5; b = torch.randn(N, generator=g)
    dy = torch.randn(M, N, generator=g)
    xd, wd, bd = (t.double().requires_grad_(True) for t in (x, w, b))
    yd = torch.nn.functional.linear(xd, wd, bd)
    yd.backward(dy.double())
    xc, wc, bc = (t.cuda().requires_grad_(True) for t in (x, w, b))
    y = LinearPS.apply(xc, wc, bc)
    y.backward(dy.cuda())
    torch.cuda.synchronize()
    tol = 3e-5
    assert relerr(y.detach().cpu().numpy(), yd.detach().numpy()) < tol
    assert relerr(xc.grad.cpu().numpy(), xd.grad.numpy()) < tol
    assert relerr(wc.grad.cpu().numpy(), wd.grad.numpy()) < tol
    assert relerr(bc.grad.cpu().numpy(), bd.grad.numpy()) < 1e-5


def test_transpose_split_and_colsum():
    from vibertgrid_pytorch_b200 import ops
    g = torch.Generator().manual_seed(3)
    x = torch.randn(301, 200, generator=g)
    t = ops.transpose_split(x.cuda(), 320)
    assert t.shape == (200, 320)
    tf = t.float().cpu()
    assert torch.equal(tf[:, 301:], torch.zeros(200, 19))
    assert relerr(tf[:, :301].numpy(), x.t().numpy()) < 2 ** -16
    t2 = ops.transpose_split(ops.to_split(x.cuda()))                      # planes in, planes out
    assert relerr(t2.float().cpu().numpy(), x.t().numpy()) < 2 ** -15
    assert relerr(ops.colsum(x.cuda()).cpu().numpy(), x.double().sum(0).numpy()) < 1e-5


@pytest.mark.parametrize("rows,cols", [(4128, 3072), (301, 200), (77, 26), (5000, 768), (9, 8)])
def test_colsum_both_formats(rows, cols):
    """Column sums over fp32 tensors and over the bf16 hi/lo plane format (what the bias gradients of the planes protocol read);
    widths that are / are not multiples of four take the float4 / the scalar kernel."""
    from vibertgrid_pytorch_b200 import ops
    g = torch.Generator().manual_seed(rows + cols)
    x = torch.randn(rows, cols, generator=g)
    want = x.double().sum(0).numpy()
    scale = float(np.abs(x.numpy()).sum(0).max())
    got = ops.colsum(x.cuda()).cpu().numpy()
    assert np.abs(got - want).max() <= 2e-6 * scale
    if cols % 8 == 0:
        xs = ops.to_split(x.cuda())
        got_s = ops.colsum(xs).cpu().numpy()
        want_s = xs.float().cpu().double().sum(0).numpy()              # the planes hold x to 2^-17
        assert np.abs(got_s - want_s).max() <= 2e-6 * scale
    again = ops.colsum(x.cuda()).cpu().numpy()
    assert np.array_equal(got, again), "fixed-order sums: bitwise reproducible"


@pytest.mark.parametrize("R,H", [(4128, 768), (37, 128), (1000, 1024)])
def test_layernorm_ps_backward(R, H):
    from vibertgrid_pytorch_b200.autograd import LayerNormPS
    g = torch.Generator().manual_seed(R + H)
    x = torch.randn(R, H, generator=g) * 2 + 0.3; gam = torch.randn(H, generator=g); bet = torch.randn(H, generator=g)
    dy = torch.randn(R, H, generator=g)
    xd, gd, bd = (t.double().requires_grad_(True) for t in (x, gam, bet))
    torch.nn.functional.layer_norm(xd, (H,), gd, bd, 1e-12).backward(dy.double())
    xc, gc, bc = (t.cuda().requires_grad_(True) for t in (x, gam, bet))
    y = LayerNormPS.apply(xc, gc, bc, 1e-12)
    y.backward(dy.cuda())
    torch.cuda.synchronize()
    assert relerr(xc.grad.cpu().numpy(), xd.grad.numpy()) < 1e-5
    assert relerr(gc.grad.cpu().numpy(), gd.grad.numpy()) < 1e-5
    assert relerr(bc.grad.cpu().numpy(), bd.grad.numpy()) < 1e-5


@pytest.mark.parametrize("B,H,W,Cin,Cout,k,p", [(2, 32, 48, 64, 64, 3, 1), (8, 16, 16, 512, 512, 3, 1), (2, 20, 12, 128, 64, 1, 0), (1, 128, 128, 64, 256, 3, 1)])
def test_conv2d_ps_data_gradient(B, H, W, Cin, Cout, k, p):
    import torch.nn.functional as F
    from vibertgrid_pytorch_b200 import ops
    from vibertgrid_pytorch_b200.autograd import Conv2dS1PS
    g = torch.Generator().manual_seed(Cin + Cout + H)
    x = torch.randn(B, Cin, H, W, generator=g); w = torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5
    dy = torch.randn(B, Cout, H, W, generator=g)
    xd = x.double().requires_grad_(True)
    yd = F.conv2d(xd, w.double(), None, 1, p)
    yd.backward(dy.double())
    xc = x.permute(0, 2, 3, 1).contiguous().cuda().requires_grad_(True)
    w_ohwi = ops.repack_oihw_to_ohwi(w.cuda())
    y = Conv2dS1PS.apply(xc, w_ohwi, p)
    y.backward(dy.permute(0, 2, 3, 1).contiguous().cuda())
    torch.cuda.synchronize()
    assert relerr(y.detach().permute(0, 3, 1, 2).cpu().numpy(), yd.detach().numpy()) < 3e-5
    assert relerr(xc.grad.permute(0, 3, 1, 2).cpu().numpy(), xd.grad.numpy()) < 3e-5


@pytest.mark.parametrize("M,N,K", [(4128, 768, 768), (516, 3072, 768), (1000, 768, 3072), (130, 128, 64), (64, 128, 128), (8256, 256, 192), (700, 64, 128), (300, 192, 64)])
def test_linear_wgrad_mn_major(M, N, K):
    """dW = dY^T X with both operands fed to tcgen05 as MN-major tiles (no transposes), row range split over CTAs with a
    deterministic finish: equals the float64 product of the plane operands; bit-identical run to run."""
    from vibertgrid_pytorch_b200 import ops
    g = torch.Generator().manual_seed(M + N + K)
    dy = torch.randn(M, N, generator=g); x = torch.randn(M, K, generator=g)
    dys, xs = ops.to_split(dy.cuda()), ops.to_split(x.cuda())
    got = ops.linear_wgrad(dys, xs)
    want = dys.float().double().t() @ xs.float().double()
    assert relerr(got.cpu().numpy(), want.cpu().numpy()) < 2e-5
    assert torch.equal(got, ops.linear_wgrad(dys, xs))


@pytest.mark.parametrize("B,H,W,Cin,Cout,k,stride,pad", [
    (2, 16, 64, 64, 128, 3, 1, 1),      # one image row per 64-pixel block
    (3, 24, 80, 128, 128, 3, 1, 1),     # W tail: the second block of a row is mostly out of bounds
    (5, 8, 8, 64, 256, 3, 1, 1),        # a block is one whole image; B odd
    (9, 7, 7, 128, 128, 3, 1, 1),       # ROI-sized: 49-row boxes, the rest of a stage stays zero
    (2, 20, 42, 64, 128, 1, 1, 0),      # 1x1, short rows (42 < 64)
    (2, 32, 64, 64, 128, 3, 2, 1),      # stride 2 through the traversal stride
    (1, 16, 32, 192, 128, 1, 1, 0),     # Cin = 3 x 64
    (2, 16, 64, 64, 64, 3, 1, 1),       # Cout = 64: the upper half of the 128-row tile is TMA zero fill
    (2, 16, 16, 128, 192, 3, 2, 1),     # Cout = 192 = 128 + 64
])
def test_conv2d_wgrad_mn_major(B, H, W, Cin, Cout, k, stride, pad):
    """dW of an NHWC convolution from the plane operands (vbg_conv2d_wgrad) against torch float64 autograd; two runs agree
    bit for bit (fixed-order finish)."""
    from vibertgrid_pytorch_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(B * 131 + H * 7 + Cin + k + stride)
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    x = torch.randn(B, H, W, Cin, device="cuda", generator=g)
    dy = torch.randn(B, Ho, Wo, Cout, device="cuda", generator=g)
    dw = ops.conv2d_wgrad(ops.to_split(dy), ops.to_split(x), k, k, stride, pad)
    dw2 = ops.conv2d_wgrad(ops.to_split(dy), ops.to_split(x), k, k, stride, pad)
    assert torch.equal(dw, dw2)
    w = torch.zeros(Cout, Cin, k, k, device="cuda", dtype=torch.float64, requires_grad=True)
    y = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2).double(), w, stride=stride, padding=pad)
    (ref,) = torch.autograd.grad(y, w, dy.permute(0, 3, 1, 2).double())
    ref = ref.permute(0, 2, 3, 1)                                              # OHWI
    err = (dw.double() - ref).abs().max().item() / ref.abs().max().item()
    assert err < 2e-5, err


def test_conv2d_s1_ps_weight_grad():
    from vibertgrid_pytorch_b200.autograd import Conv2dS1PS
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn(2, 16, 64, 64, device="cuda", generator=g, requires_grad=True)
    w = (torch.randn(128, 3, 3, 64, device="cuda", generator=g) * 0.05).requires_grad_()
    y = Conv2dS1PS.apply(x, w, 1)
    dy = torch.randn_like(y)
    dx, dw = torch.autograd.grad(y, (x, w), dy)
    xr = x.detach().double().permute(0, 3, 1, 2).requires_grad_()
    wr = w.detach().double().permute(0, 3, 1, 2).requires_grad_()
    yr = torch.nn.functional.conv2d(xr, wr, padding=1)
    dxr, dwr = torch.autograd.grad(yr, (xr, wr), dy.double().permute(0, 3, 1, 2))
    assert (dx.double() - dxr.permute(0, 2, 3, 1)).abs().max().item() / dxr.abs().max().item() < 2e-5
    assert (dw.double() - dwr.permute(0, 2, 3, 1)).abs().max().item() / dwr.abs().max().item() < 2e-5


@pytest.mark.parametrize("M,N,K", [(3000, 26, 256), (257, 23, 1024), (129, 25, 1024), (500, 5, 512), (64, 100, 64)])
def test_linear_small_any_width(M, N, K):
    """Output widths the tensor-core kernels do not take and that exceed the 16-column tile of ``vbg_small_wgrad``:
    the reference's 23-tag EPHOIE BIO set gives a packed segmentation head of 3 + 23 = 26 columns, a 23-way simp head and 25
    CRF emissions (model/field_type_classification_head.py:635-637).  Forward, dX, dW, db against float64 autograd."""
    from vibertgrid_pytorch_b200.autograd import LinearSmall, linear
    g = torch.Generator().manual_seed(M + N + K)
    x = torch.randn(M, K, generator=g); w = torch.randn(N, K, generator=g) / K ** 0.5; b = torch.randn(N, generator=g)
    dy = torch.randn(M, N, generator=g)
    xd, wd, bd = (t.double().requires_grad_(True) for t in (x, w, b))
    torch.nn.functional.linear(xd, wd, bd).backward(dy.double())
    xc, wc, bc = (t.cuda().requires_grad_(True) for t in (x, w, b))
    y = linear(xc, wc, bc)
    assert y.grad_fn is not None and "LinearSmall" in type(y.grad_fn).__name__
    y.backward(dy.cuda())
    torch.cuda.synchronize()
    assert relerr(y.detach().cpu().numpy(), torch.nn.functional.linear(x.double(), w.double(), b.double()).numpy()) < 2e-5
    assert relerr(xc.grad.cpu().numpy(), xd.grad.numpy()) < 2e-5
    assert relerr(wc.grad.cpu().numpy(), wd.grad.numpy()) < 2e-5
    assert relerr(bc.grad.cpu().numpy(), bd.grad.numpy()) < 2e-5
