"""nn.SyncBatchNorm form of the training-mode BatchNorm stage on a world_size-2 gloo group (CPU): the real autograd Function
(vibertgrid_pytorch_b200.autograd.BatchNormTrainF with ``sync``) over torch stand-ins of its kernels must equal plain
BatchNorm over the CONCATENATED rows of both ranks -- outputs, data gradients, parameter gradients (summed over ranks) and the
running-statistics update -- which is what the reference gets from ``convert_sync_batchnorm`` (train_SROIE.py:203-205)."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROWS = (10, 22)         # unequal populations: the combination must be count-weighted
C = 8


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _data(relu, res):
    g = torch.Generator().manual_seed(5)
    n = sum(ROWS)
    x = torch.randn(n, 1, 1, C, generator=g) * 2 + 0.5
    r = torch.randn(n, 1, 1, C, generator=g) if res else None
    dy = torch.randn(n, 1, 1, C, generator=g)
    gamma = torch.rand(C, generator=g) + 0.5
    beta = torch.randn(C, generator=g)
    return x, r, dy, gamma, beta


def _worker(rank, world, port, q):
    sys.path.insert(0, HERE)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import mock_ops
        from vibertgrid_pytorch_b200 import autograd as A
        from vibertgrid_pytorch_b200 import train_engine as te
        A.ops = mock_ops
        out = {}
        lo, hi = sum(ROWS[:rank]), sum(ROWS[:rank + 1])
        for relu, res in ((True, True), (False, False)):
            x, r, dy, gamma, beta = _data(relu, res)
            bn = nn.SyncBatchNorm(C)
            with torch.no_grad():
                bn.weight.copy_(gamma)
                bn.bias.copy_(beta)
            eng = te.TrainEngine(net=None)
            xl = x[lo:hi].clone().requires_grad_()
            rl = None if r is None else r[lo:hi].clone().requires_grad_()
            y = eng._bn(xl, bn, relu=relu, residual=rl)
            y.backward(dy[lo:hi])
            out[(relu, res)] = dict(y=y.detach(), dx=xl.grad, dres=None if rl is None else rl.grad, dg=bn.weight.grad, db=bn.bias.grad,
                                    rm=bn.running_mean.clone(), rv=bn.running_var.clone(), nbt=int(bn.num_batches_tracked))
        # a plain BatchNorm2d is NOT synchronised even inside a process group
        assert te.TrainEngine._sync_group(nn.BatchNorm2d(C)) is None and te.TrainEngine._sync_group(nn.SyncBatchNorm(C)) == (None,)
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


def test_syncbn_world2_equals_batchnorm_over_all_rows():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for relu, rs in ((True, True), (False, False)):
        x, r, dy, gamma, beta = _data(relu, rs)
        xd = x.double().permute(0, 3, 1, 2).requires_grad_()
        gd, bd = gamma.double().requires_grad_(), beta.double().requires_grad_()
        rd = None if r is None else r.double().permute(0, 3, 1, 2).requires_grad_()
        rm, rv = torch.zeros(C, dtype=torch.float64), torch.ones(C, dtype=torch.float64)
        y = F.batch_norm(xd, rm, rv, gd, bd, True, 0.1, 1e-5)
        if rd is not None:
            y = y + rd
        if relu:
            y = F.relu(y)
        y.backward(dy.double().permute(0, 3, 1, 2))
        nhwc = lambda t: t.permute(0, 2, 3, 1)
        close = lambda a, b: float((a.double() - b).abs().max()) <= 2e-5 * max(1.0, float(b.abs().max()))
        for rank in range(world):
            lo, hi = sum(ROWS[:rank]), sum(ROWS[:rank + 1])
            o = res[rank][(relu, rs)]
            assert close(o["y"], nhwc(y.detach())[lo:hi]) and close(o["dx"], nhwc(xd.grad)[lo:hi])
            if rd is not None:
                assert close(o["dres"], nhwc(rd.grad)[lo:hi])
            assert close(o["rm"], rm) and close(o["rv"], rv) and o["nbt"] == 1
        assert close(res[0][(relu, rs)]["dg"] + res[1][(relu, rs)]["dg"], gd.grad)     # per-rank parameter gradients add up
        assert close(res[0][(relu, rs)]["db"] + res[1][(relu, rs)]["db"], bd.grad)
