"""nn.SyncBatchNorm form of the BatchNorm stage on the GPU: the split backward kernels (vbg_bn_bwd_reduce / vbg_bn_bwd_dx) and the
NCCL plumbing of ``BatchNormTrainF(sync=...)`` in a one-rank process group, where the synchronised statistics must equal the
per-rank ones (float64 torch autograd is the reference).  The count-weighted combination over two ranks is covered on CPU by
tests/test_syncbn_gloo.py.  (File sorts last: it owns a process group for its duration.)"""
import socket

import pytest
import torch
import torch.distributed as dist
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def one_rank_group():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    torch.cuda.set_device(0)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=0, world_size=1)
    yield
    dist.destroy_process_group()


@pytest.mark.parametrize("B,H,W,C,relu,res", [(2, 13, 17, 64, True, True), (3, 8, 8, 128, False, False), (8, 32, 32, 256, True, False)])
def test_syncbn_one_rank_matches_float64(one_rank_group, B, H, W, C, relu, res):
    from vibertgrid_pytorch_b200 import autograd as A
    g = torch.Generator(device="cuda").manual_seed(C + H)
    x = torch.randn(B, H, W, C, device="cuda", generator=g) * 2 + 0.5
    r = torch.randn(B, H, W, C, device="cuda", generator=g) if res else None
    gamma = torch.rand(C, device="cuda", generator=g) + 0.5
    beta = torch.randn(C, device="cuda", generator=g)
    dy = torch.randn(B, H, W, C, device="cuda", generator=g)
    eps = 1e-5
    outs = {}
    for sync in (None, (None,)):
        xs, gs, bs = x.clone().requires_grad_(), gamma.clone().requires_grad_(), beta.clone().requires_grad_()
        rs = None if r is None else r.clone().requires_grad_()
        stats = []
        y = A.BatchNormTrainF.apply(xs, gs, bs, rs, relu, eps, stats, sync)
        y.backward(dy)
        outs[sync is None] = (y.detach(), xs.grad, gs.grad, bs.grad, None if rs is None else rs.grad, stats[0])
    xd = x.double().permute(0, 3, 1, 2).requires_grad_()
    gd, bd = gamma.double().requires_grad_(), beta.double().requires_grad_()
    rd = None if r is None else r.double().permute(0, 3, 1, 2).requires_grad_()
    yd = F.batch_norm(xd, None, None, gd, bd, True, 0.1, eps)
    if rd is not None:
        yd = yd + rd
    if relu:
        yd = F.relu(yd)
    yd.backward(dy.double().permute(0, 3, 1, 2))
    nhwc = lambda t: t.permute(0, 2, 3, 1)
    rel = lambda a, b: float((a.double() - b).abs().max() / b.abs().max().clamp_min(1e-30))
    for plain in (True, False):
        y, dx, dg, db, dres, st = outs[plain]
        assert rel(y, nhwc(yd.detach())) <= 2e-5 and rel(dx, nhwc(xd.grad)) <= 2e-5
        assert rel(dg, gd.grad) <= 2e-5 and rel(db, bd.grad) <= 2e-5
        if rd is not None:
            assert rel(dres, nhwc(rd.grad)) <= 2e-5
    n = outs[False][5][2]
    assert isinstance(n, torch.Tensor) and float(n) == B * H * W            # the global row count stays on the device
    assert rel(outs[False][5][1], outs[True][5][1].double()) <= 1e-5       # same (biased) variance as the per-rank path


def test_training_step_with_converted_syncbn(one_rank_group, tmp_path, monkeypatch):
    """The reference's multi-GPU recipe on the drop-in: ``convert_sync_batchnorm`` (train_SROIE.py:203-205), then a training
    step.  In a one-rank group the result must equal the unconverted module's step."""
    import random
    from conftest import build_case, load_golden
    from vibertgrid_pytorch_b200 import train_engine as te
    fx = load_golden("train_tiny")
    monkeypatch.chdir(tmp_path)
    monkeypatch.setenv("VBG_PRECISION", "fp32")
    losses, grads = [], []
    for convert in (False, True):
        cfg, kw, net, batch = build_case(fx["meta"])
        if convert:
            net = torch.nn.SyncBatchNorm.convert_sync_batchnorm(net)
            assert any(isinstance(m, torch.nn.SyncBatchNorm) for m in net.modules())
        net = net.cuda()
        net.train()
        net.bert_hidden_dropout = net.bert_attn_dropout = 0.0
        monkeypatch.setattr(te, "FORCE_SYNC_BN_SINGLE_RANK", convert)
        random.seed(0)
        img, seg, cls, coors, corpus, mask = batch
        c = lambda ts: tuple(t.cuda() for t in ts)
        loss = net(c(img), c(seg), c(cls), c(coors), corpus.cuda(), mask.cuda())
        loss.backward()
        losses.append(float(loss))
        grads.append({k: p.grad.clone() for k, p in net.named_parameters() if p.grad is not None})
    want = float(fx["loss"][0])
    assert abs(losses[0] - want) <= 1e-3 * max(1.0, abs(want)) and abs(losses[1] - losses[0]) <= 1e-4 * max(1.0, abs(want))
    assert grads[0].keys() == grads[1].keys()
    for k in grads[0]:
        a, b = grads[0][k].double(), grads[1][k].double()
        if k.endswith("attention.self.key.bias"):      # exactly zero in exact arithmetic: rounding noise only
            continue
        # same kernels, statistics combined through float64 in one of the two runs: rounding-level differences, amplified by
        # the tiny BatchNorm populations of this fixture (tests/test_gpu_train_step.py); a wiring error is O(1)
        assert float((a - b).norm()) <= 5e-2 * max(float(a.norm()), 1e-12), k
        assert abs(float(a.norm()) - float(b.norm())) <= 5e-3 * max(float(a.norm()), 1e-12), k
