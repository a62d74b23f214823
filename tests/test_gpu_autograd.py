"""Linear-layer forward + backward on the pre-split tcgen05 GEMM (vibertgrid_pytorch_b200.autograd.LinearPS) against
torch's float64 autograd of the same layer: y, dX, dW, db within fp32-class tolerance (bf16x3 products, fp32 accumulate)."""
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
