"""Kernels of the training step (csrc/vbg_train.cu, csrc/vbg_attn_bwd.cu) against torch's float64 autograd of the same op on
the same seeded inputs.  Tolerances are fp32-class (the kernels are plain fp32 with fixed-order reductions)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def gen(seed):
    return torch.Generator(device="cuda").manual_seed(seed)


@pytest.mark.parametrize("B,H,W,C,relu,res", [(2, 13, 17, 64, True, True), (3, 8, 8, 128, True, False), (1, 5, 7, 512, False, False),
                                             (8, 64, 64, 64, True, True), (5, 7, 7, 256, True, False)])
def test_batchnorm_train_forward_backward(B, H, W, C, relu, res):
    from vibertgrid_pytorch_b200 import ops
    g = gen(C + H)
    x = torch.randn(B, H, W, C, device="cuda", generator=g) * 2 + 0.5
    r = torch.randn(B, H, W, C, device="cuda", generator=g) if res else None
    gamma = torch.rand(C, device="cuda", generator=g) + 0.5
    beta = torch.randn(C, device="cuda", generator=g)
    dy = torch.randn(B, H, W, C, device="cuda", generator=g)
    eps = 1e-5
    mean, var, rstd = ops.bn_stats(x.view(-1, C), eps)
    y = ops.bn_apply(x.view(-1, C), mean, rstd, gamma, beta, None if r is None else r.view(-1, C), relu)
    dx, dres, dg, db = ops.bn_bwd(x.view(-1, C), dy.view(-1, C), y if relu else None, mean, rstd, gamma, want_dres=res)

    xd = x.double().permute(0, 3, 1, 2).requires_grad_()
    gd, bd = gamma.double().requires_grad_(), beta.double().requires_grad_()
    rd = None if r is None else r.double().permute(0, 3, 1, 2).requires_grad_()
    yd = F.batch_norm(xd, None, None, gd, bd, True, 0.1, eps)
    if rd is not None:
        yd = yd + rd
    if relu:
        yd = F.relu(yd)
    grads = torch.autograd.grad(yd, [xd, gd, bd] + ([rd] if rd is not None else []), dy.double().permute(0, 3, 1, 2))
    nhwc = lambda t: t.permute(0, 2, 3, 1).reshape(-1, C)
    assert rel(mean, xd.mean((0, 2, 3))) < 1e-5 and rel(var, xd.var((0, 2, 3), unbiased=False)) < 1e-5
    assert rel(y, nhwc(yd.detach())) < 1e-5
    assert rel(dx, nhwc(grads[0])) < 2e-5
    assert rel(dg, grads[1]) < 2e-5 and rel(db, grads[2]) < 2e-5
    if res:
        assert rel(dres, nhwc(grads[3])) < 1e-6


@pytest.mark.parametrize("B,H,W,C", [(2, 16, 20, 64), (1, 9, 11, 8), (3, 32, 32, 64)])
def test_maxpool_bwd_first_max_rule(B, H, W, C):
    from vibertgrid_pytorch_b200 import ops
    g = gen(H * W)
    x = torch.relu(torch.randint(-3, 4, (B, H, W, C), device="cuda", generator=g).float())     # many ties (zeros and small integers)
    y = ops.maxpool3x3s2(x)
    dy = torch.randn(y.shape, device="cuda", generator=g)
    dx = ops.maxpool3x3s2_bwd(x, dy)
    xr = x.permute(0, 3, 1, 2).contiguous().requires_grad_()
    yr = F.max_pool2d(xr, 3, 2, 1)
    assert torch.equal(yr.permute(0, 2, 3, 1), y)
    (ref,) = torch.autograd.grad(yr, xr, dy.permute(0, 3, 1, 2).contiguous())
    assert rel(dx, ref.permute(0, 2, 3, 1)) < 1e-6
    assert torch.equal(ops.maxpool3x3s2_bwd(x, dy, y), dx)            # the y-assisted gather: same rule, same bits


def test_sumpool_expand_gelu_dropout():
    from vibertgrid_pytorch_b200 import ops
    g = gen(3)
    x = torch.randn(2, 12, 10, 32, device="cuda", generator=g)
    ref = F.avg_pool2d(x.permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1) * 4
    assert rel(ops.sumpool2x2(x), ref) < 1e-6
    up = ops.expand2x(x, 24, 20, 0.25)
    assert rel(up, 0.25 * x.repeat_interleave(2, 1).repeat_interleave(2, 2)) < 1e-7
    z = ops.expand2x(x, 23, 20, 1.0, zero_insert=True)
    zr = torch.zeros(2, 23, 20, 32, device="cuda")
    zr[:, ::2, ::2] = x[:, :12, :10]
    assert torch.equal(z, zr)
    u = torch.randn(1000, 64, device="cuda", generator=g) * 3
    ud = u.double().requires_grad_()
    yd = F.gelu(ud)
    dy = torch.randn(1000, 64, device="cuda", generator=g)
    (gd,) = torch.autograd.grad(yd, ud, dy.double())
    assert rel(ops.gelu(u), yd.detach()) < 1e-6 and rel(ops.gelu(u, dy), gd) < 2e-6
    d1, d2 = ops.dropout(u, 0.1, 1234), ops.dropout(u, 0.1, 1234)
    assert torch.equal(d1, d2)
    kept = (d1 != 0).float().mean().item()
    assert abs(kept - 0.9) < 0.01
    m = d1 != 0
    assert rel(d1[m], u[m] / 0.9) < 1e-6
    assert not torch.equal(ops.dropout(u, 0.1, 99), d1)


def test_grid_scatter_and_segment_mean_backward():
    from vibertgrid_pytorch_b200 import ops
    g = gen(11)
    B, Hg, Wg, C, stride = 2, 12, 16, 64, 8
    seg_counts = [7, 5]
    K = sum(seg_counts)
    seg_off = torch.tensor([0, 7, 12], dtype=torch.int32, device="cuda")
    boxes = torch.zeros(K, 4, dtype=torch.int32, device="cuda")
    cpu = torch.Generator().manual_seed(5)
    for k in range(K):
        x1, y1 = int(torch.randint(0, 100, (1,), generator=cpu)), int(torch.randint(0, 70, (1,), generator=cpu))
        boxes[k] = torch.tensor([x1, y1, x1 + int(torch.randint(4, 60, (1,), generator=cpu)), y1 + int(torch.randint(4, 40, (1,), generator=cpu))])
    idx = ops.box_index_map(boxes, seg_off, B, stride, Hg, Wg)
    ld = C + 32
    dwide = torch.randn(B * Hg * Wg, ld, device="cuda", generator=g)
    dgrid = dwide[:, 32:]                                   # column slice of a wider gradient
    demb = ops.grid_scatter_bwd(dgrid, ld, idx, boxes, seg_off, B, K, stride, C)
    ref = torch.zeros(K, C, dtype=torch.float64, device="cuda")
    flat = idx.view(B, -1)
    for b in range(B):
        m = flat[b] >= 0
        ref.index_add_(0, (flat[b][m] + int(seg_off[b])).long(), dgrid.view(B, Hg * Wg, C)[b][m].double())
    assert rel(demb, ref) < 1e-6
    # segment mean backward
    n_tok = [20, 9]
    seg_ids = torch.cat([torch.sort(torch.randint(0, s, (n,), generator=cpu))[0] for s, n in zip(seg_counts, n_tok)])
    # runs: consecutive equal ids; build seg_start / tok_row by hand (token t of sample b lives at row 1 + t + 40 b)
    starts, rows, base = [], [], 0
    for b, n in enumerate(n_tok):
        ids = seg_ids[base:base + n].tolist()
        for t in range(n):
            if t == 0 or ids[t] != ids[t - 1]:
                starts.append(base + t)
            rows.append(1 + t + 40 * b)
        base += n
    Kr = len(starts)
    seg_start = torch.tensor(starts + [sum(n_tok)], dtype=torch.int32, device="cuda")
    tok_row = torch.tensor(rows, dtype=torch.int32, device="cuda")
    R = 80
    hid = torch.randn(R, C, device="cuda", generator=g)
    dseg = torch.randn(Kr, C, device="cuda", generator=g)
    for mode in (ops.AGG_MEAN, ops.AGG_FIRST):
        out = ops.segment_reduce(hid, tok_row, seg_start, Kr, mode)
        hd = hid.double().requires_grad_()
        refs = []
        for k in range(Kr):
            r = tok_row[starts[k]:int(seg_start[k + 1])].long()
            refs.append(hd[r].mean(0) if mode == ops.AGG_MEAN else hd[r[0]])
        refo = torch.stack(refs)
        assert rel(out, refo.detach()) < 1e-6
        (gref,) = torch.autograd.grad(refo, hd, dseg.double())
        assert rel(ops.segment_reduce_bwd(dseg, tok_row, seg_start, R, mode), gref) < 1e-6


def test_embed_bwd_scatter_add():
    from vibertgrid_pytorch_b200 import ops
    g = gen(2)
    R, Hd, V, Pm = 300, 128, 50, 64
    dx = torch.randn(R, Hd, device="cuda", generator=g)
    ids = torch.randint(0, V, (R,), device="cuda", generator=g).int()
    pos = torch.randint(0, Pm, (R,), device="cuda", generator=g).int()
    dw, dp = ops.embed_bwd(dx, ids, pos, V, Pm)
    rw = torch.zeros(V, Hd, dtype=torch.float64, device="cuda").index_add_(0, ids.long(), dx.double())
    rp = torch.zeros(Pm, Hd, dtype=torch.float64, device="cuda").index_add_(0, pos.long(), dx.double())
    assert rel(dw, rw) < 1e-5 and rel(dp, rp) < 1e-5
    dw2, dp2 = ops.embed_bwd(dx, ids, pos, V, Pm)
    assert torch.equal(dw, dw2) and torch.equal(dp, dp2)                 # ascending-row sums, no atomics: bitwise reproducible
    # BERT-sized: 4128 packed rows, vocabulary 30522, a token repeated a thousand times
    R, Hd, V, Pm = 4128, 768, 30522, 512
    dx = torch.randn(R, Hd, device="cuda", generator=g)
    ids = torch.randint(1000, V, (R,), device="cuda", generator=g).int()
    ids[::4] = 1012
    pos = (torch.arange(R, device="cuda") % 514).clamp_max(511).int()
    dw, dp = ops.embed_bwd(dx, ids, pos, V, Pm)
    rw = torch.zeros(V, Hd, dtype=torch.float64, device="cuda").index_add_(0, ids.long(), dx.double())
    rp = torch.zeros(Pm, Hd, dtype=torch.float64, device="cuda").index_add_(0, pos.long(), dx.double())
    assert rel(dw, rw) < 1e-5 and rel(dp, rp) < 1e-5


def test_roi_align_bwd_matches_torchvision():
    torchvision = pytest.importorskip("torchvision")
    from vibertgrid_pytorch_b200 import ops
    g = gen(8)
    B, Hf, Wf, C, P = 2, 24, 32, 64, 7
    feat = torch.randn(B, Hf, Wf, C, device="cuda", generator=g)
    seg_off = torch.tensor([0, 5, 9], dtype=torch.int32, device="cuda")
    boxes = torch.tensor([[4, 4, 60, 20], [0, 0, 127, 95], [30, 50, 34, 52], [100, 10, 140, 90], [-8, -4, 20, 30],
                          [10, 10, 11, 11], [64, 32, 120, 40], [5, 80, 100, 96], [0, 0, 8, 8]], dtype=torch.int32, device="cuda")
    out = ops.roi_align(feat, boxes, seg_off, 0.25, P)
    dout = torch.randn(out.shape, device="cuda", generator=g)
    dfeat = ops.roi_align_bwd(dout, boxes, seg_off, B, Hf, Wf, 0.25)
    fr = feat.double().permute(0, 3, 1, 2).contiguous().requires_grad_()
    lists = [boxes[int(seg_off[b]):int(seg_off[b + 1])].double() for b in range(B)]
    ref = torchvision.ops.roi_align(fr, lists, output_size=P, spatial_scale=0.25, sampling_ratio=-1, aligned=False)   # [K,C,P,P]
    assert rel(out, ref.detach().permute(0, 2, 3, 1)) < 1e-5
    (gref,) = torch.autograd.grad(ref, fr, dout.double().permute(0, 3, 1, 2).contiguous())
    assert rel(dfeat, gref.permute(0, 2, 3, 1)) < 1e-5
    assert torch.equal(dfeat, ops.roi_align_bwd(dout, boxes, seg_off, B, Hf, Wf, 0.25))      # gather form: no atomics, bitwise reproducible


def test_roi_align_bwd_many_rois_per_tile_and_wide_channels():
    """More ROIs over one 8 x 8 tile than a round of the gather kernel lists (32), page-sized ROIs (bins of > 32 pixels), a
    document without ROIs, C = 512 (two channel slabs), map sizes that are not multiples of the tile."""
    torchvision = pytest.importorskip("torchvision")
    from vibertgrid_pytorch_b200 import ops
    g = gen(9)
    B, Hf, Wf, C, P = 3, 75, 83, 512, 7
    counts = [70, 0, 9]
    per = []
    for n in counts:
        l = torch.randint(0, 60, (n,), generator=torch.Generator().manual_seed(n)); t = torch.randint(0, 60, (n,), generator=torch.Generator().manual_seed(n + 1))
        per.append(torch.stack([l, t, l + torch.randint(1, 270, (n,), generator=torch.Generator().manual_seed(n + 2)),
                                t + torch.randint(1, 240, (n,), generator=torch.Generator().manual_seed(n + 3))], 1))
    per[0][0] = torch.tensor([0, 0, Wf * 4 - 1, Hf * 4 - 1]); per[0][1] = torch.tensor([0, 0, 2000, 2000]); per[0][2] = torch.tensor([10, 10, 10, 10])
    boxes = torch.cat(per).int().cuda()
    seg_off = torch.tensor([0, 70, 70, 79], dtype=torch.int32, device="cuda")
    dout = torch.randn(79, P, P, C, device="cuda", generator=g)
    dfeat = ops.roi_align_bwd(dout, boxes, seg_off, B, Hf, Wf, 0.25)
    fr = torch.zeros(B, C, Hf, Wf, dtype=torch.float64, device="cuda", requires_grad=True)
    lists = [boxes[int(seg_off[b]):int(seg_off[b + 1])].double() for b in range(B)]
    ref = torchvision.ops.roi_align(fr, lists, output_size=P, spatial_scale=0.25, sampling_ratio=-1, aligned=False)
    (gref,) = torch.autograd.grad(ref, fr, dout.double().permute(0, 3, 1, 2).contiguous())
    assert rel(dfeat, gref.permute(0, 2, 3, 1)) < 1e-5
    assert float(dfeat[1].abs().max()) == 0.0
    assert torch.equal(dfeat, ops.roi_align_bwd(dout, boxes, seg_off, B, Hf, Wf, 0.25))


def test_seg_ce_and_upsample_backward():
    from vibertgrid_pytorch_b200 import ops
    g = gen(4)
    B, H, W, up, C = 2, 32, 48, 4, 5
    Ct = 3 + C
    lg = torch.randn(B, H // up, W // up, Ct, device="cuda", generator=g)
    pn = torch.randint(0, 3, (B, H, W), device="cuda", generator=g)
    cl = torch.randint(0, C, (B, H, W), device="cuda", generator=g)
    gs = torch.tensor([0.7, 1.3], device="cuda")
    dl = ops.seg_ce_bwd(lg, pn, cl, H, W, up, 3, gs)
    lr = lg.double().requires_grad_()
    full = lr.permute(0, 3, 1, 2).repeat_interleave(up, 2).repeat_interleave(up, 3)
    loss = 0.7 * F.cross_entropy(full[:, :3], pn) + 1.3 * F.cross_entropy(full[:, 3:], cl)
    (ref,) = torch.autograd.grad(loss, lr)
    assert rel(dl, ref) < 2e-5
    d1 = torch.randn(B, 3, H, W, device="cuda", generator=g)
    d2 = torch.randn(B, C, H, W, device="cuda", generator=g)
    (ref2,) = torch.autograd.grad(full, lr, torch.cat([d1, d2], 1).double())
    assert rel(ops.upsample_split_bwd(d1, d2, up), ref2) < 1e-6


@pytest.mark.parametrize("M,N,K", [(5000, 8, 256), (131, 2, 512), (70000, 5, 300), (64, 16, 64)])
def test_small_wgrad(M, N, K):
    from vibertgrid_pytorch_b200 import ops
    g = gen(M)
    wide = torch.randn(M, N + 3, device="cuda", generator=g)
    dy = wide[:, 1:1 + N]
    x = torch.randn(M, K, device="cuda", generator=g)
    dw = ops.small_wgrad(dy, x)
    assert rel(dw, dy.double().t() @ x.double()) < 2e-5
    assert torch.equal(dw, ops.small_wgrad(dy, x))


@pytest.mark.parametrize("B,H,W", [(2, 64, 96), (1, 50, 38), (3, 128, 128)])
def test_stem_wgrad(B, H, W):
    from vibertgrid_pytorch_b200 import ops
    g = gen(H)
    x = torch.randn(B, 3, H, W, device="cuda", generator=g)
    x4 = torch.zeros(B, H + 6, W + 6, 4, device="cuda")
    x4[:, 3:-3, 3:-3, :3] = x.permute(0, 2, 3, 1)
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    dy = torch.randn(B, Ho, Wo, 64, device="cuda", generator=g)
    dw = ops.stem_wgrad(x4, dy)
    w = torch.zeros(64, 3, 7, 7, device="cuda", dtype=torch.float64, requires_grad=True)
    y = F.conv2d(x.double(), w, stride=2, padding=3)
    (ref,) = torch.autograd.grad(y, w, dy.double().permute(0, 3, 1, 2))
    assert rel(dw[..., :3].permute(0, 3, 1, 2), ref) < 2e-5
    assert torch.equal(dw, ops.stem_wgrad(x4, dy))


@pytest.mark.parametrize("lens,heads", [([130, 64, 2, 200], 3), ([512, 2, 511], 2), ([1], 1)])
def test_attention_bwd(lens, heads):
    from vibertgrid_pytorch_b200 import ops
    g = gen(sum(lens))
    hid = heads * 64
    R = sum(lens)
    qkv = torch.randn(R, 3 * hid, device="cuda", generator=g)
    d_o = torch.randn(R, hid, device="cuda", generator=g)
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device="cuda")
    qd = qkv.double().requires_grad_()
    outs = []
    for i, n in enumerate(lens):
        s = qd[int(cu[i]):int(cu[i + 1])].view(n, 3, heads, 64)
        q, k, v = (s[:, j].transpose(0, 1) for j in range(3))             # [heads, n, 64]
        p = torch.softmax(q @ k.transpose(1, 2) / 8.0, -1)
        outs.append((p @ v).transpose(0, 1).reshape(n, hid))
    o = torch.cat(outs)
    (ref,) = torch.autograd.grad(o, qd, d_o.double())
    dqkv = ops.attention_bwd(qkv, o.detach().float().contiguous(), d_o, cu, len(lens), max(lens), heads)
    assert rel(dqkv, ref) < 2e-5
    assert torch.equal(dqkv, ops.attention_bwd(qkv, o.detach().float().contiguous(), d_o, cu, len(lens), max(lens), heads))


@pytest.mark.parametrize("seed,lens,C", [(0, [7, 5], 4), (1, [64, 64], 4), (2, [1, 2, 130], 12), (3, [200], 30),
                                         (4, [65, 63, 128, 64, 1, 300, 17, 90], 6), (5, [1024] * 4, 12)])
def test_crf_nll_forward_backward(seed, lens, C):
    """vbg_crf_nll_{fwd,bwd} through the autograd Function against the oracle's float64 restatement of model/crf.py
    (cfg5: 2 documents x 64 segments, T = 6; the last case is a cfg4-sized character-level batch)."""
    from oracle import oracle_ops
    from vibertgrid_pytorch_b200 import autograd as A
    g = torch.Generator().manual_seed(seed)
    T, K, B = C + 2, sum(lens), len(lens)
    feats = torch.randn(K, T, generator=g) * 1.5
    trans = torch.randn(T, T, generator=g)
    trans[T - 2, :] = -10000.0
    trans[:, T - 1] = -10000.0
    tags = torch.randint(0, C, (K,), generator=g).to(torch.int32)
    off = [0]
    for n in lens:
        off.append(off[-1] + n)
    w = (torch.arange(B, dtype=torch.float32) + 1.0) / B
    fd, td = feats.clone().cuda().requires_grad_(), trans.clone().cuda().requires_grad_()
    nll = A.CrfNllF.apply(fd, td, tags.cuda(), torch.tensor(off, dtype=torch.int32).cuda(), B)
    (nll * w.cuda()).sum().backward()
    f64, t64 = feats.double().requires_grad_(), trans.double().requires_grad_()
    want = torch.stack([oracle_ops.crf_nll_torch(f64[off[b]:off[b + 1]], tags[off[b]:off[b + 1]], t64, T - 2, T - 1)
                        for b in range(B)])
    (want * w.double()).sum().backward()
    assert (nll.cpu().double() - want.detach()).abs().max() <= 2e-5 * max(1.0, float(want.abs().max()))
    assert (fd.grad.cpu().double() - f64.grad).abs().max() <= 2e-5
    assert (td.grad.cpu().double() - t64.grad).abs().max() <= 2e-5 * max(1.0, float(t64.grad.abs().max()))
