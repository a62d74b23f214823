"""Per-kernel parity: each C-ABI entry point against the oracle (numpy) or a plain
torch fp32 CPU computation of the same op, on seeded inputs.  Integer outputs are
compared bit-exactly; fp32 outputs to the tolerance written in each test."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import relerr

from oracle import oracle_ops

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from vibertgrid_pytorch_b200 import ops as o
    assert torch.cuda.is_available()
    return o


def _boxes(rng, S, H, W, tail=True):
    l = rng.integers(0, W - 2, S); t = rng.integers(0, H - 2, S)
    r = np.minimum(l + rng.integers(1, W // 2, S), W - 1); b = np.minimum(t + rng.integers(1, H // 3, S), H - 1)
    bx = np.stack([l, t, r, b], 1).astype(np.int32)
    if tail and S > 4:
        bx[1] = [5, 5, 5, 9]          # empty slice
        bx[2] = [0, 0, W + 40, H + 9]  # clipped by the array end
        bx[3] = [9, 9, 3, 3]          # inverted -> empty
    return bx


def _dev_off(counts):
    off = np.zeros(len(counts) + 1, np.int32); off[1:] = np.cumsum(counts)
    return off, torch.from_numpy(off).cuda()


@pytest.mark.parametrize("stride,H,W,counts", [(8, 512, 512, [128, 97]), (8, 96, 160, [9, 0, 300]), (4, 64, 64, [5]), (1, 96, 128, [40, 1])])
def test_box_index_map_bit_exact(ops, stride, H, W, counts):
    rng = np.random.default_rng(1)
    per = [_boxes(rng, c, H, W) for c in counts]
    off, doff = _dev_off(counts)
    boxes = torch.from_numpy(np.concatenate(per + [np.zeros((0, 4), np.int32)], 0)).cuda()
    if boxes.shape[0] == 0:
        boxes = torch.zeros((1, 4), dtype=torch.int32).cuda()
    idx = ops.box_index_map(boxes, doff, len(counts), stride, H // stride, W // stride)
    assert np.array_equal(idx.cpu().numpy(), oracle_ops.box_index_map(per, H, W, stride))


def test_scatter_and_label_paint(ops):
    rng = np.random.default_rng(2)
    H, W, counts, C = 128, 192, [33, 20], 768
    per = [_boxes(rng, c, H, W) for c in counts]
    off, doff = _dev_off(counts)
    boxes = torch.from_numpy(np.concatenate(per, 0)).cuda()
    emb = torch.randn(sum(counts), C)
    idx = ops.box_index_map(boxes, doff, 2, 8, H // 8, W // 8)
    grid = ops.grid_scatter(emb.cuda(), idx, doff)
    want = oracle_ops.scatter_grid([emb[off[b]:off[b + 1]].numpy() for b in range(2)], idx.cpu().numpy())
    assert np.array_equal(grid.permute(0, 3, 1, 2).cpu().numpy(), want)            # pure copy: bit-exact
    cls = [rng.integers(0, 5, c).astype(np.int32) for c in counts]
    pn, cl = ops.label_paint(boxes, doff, torch.from_numpy(np.concatenate(cls)).cuda(), 2, H, W)
    wpn, wcl = oracle_ops.paint_labels(oracle_ops.box_index_map(per, H, W, 1), cls)
    assert np.array_equal(pn.cpu().numpy(), wpn) and np.array_equal(cl.cpu().numpy(), wcl)


def test_fused_seg_ce_loss(ops):
    """vbg_seg_ce_loss == F.cross_entropy of the x4-upsampled logits against the painted label maps (mean over pixels)."""
    rng = np.random.default_rng(6)
    g = torch.Generator().manual_seed(6)
    H, W, counts, Cn, up = 64, 96, [23, 7], 5, 4
    per = [_boxes(rng, c, H, W) for c in counts]
    off, doff = _dev_off(counts)
    boxes = torch.from_numpy(np.concatenate(per, 0)).cuda()
    cls = torch.from_numpy(np.concatenate([rng.integers(0, Cn, c).astype(np.int32) for c in counts])).cuda()
    lg = (torch.randn(2, H // up, W // up, 3 + Cn, generator=g) * 3).cuda()
    got = ops.seg_ce_loss(boxes, doff, cls, lg, 2, H, W, up, 3).cpu()
    pn, cl = ops.label_paint(boxes, doff, cls, 2, H, W)
    full = lg.permute(0, 3, 1, 2).repeat_interleave(up, 2).repeat_interleave(up, 3).double()
    want = torch.stack([F.cross_entropy(full[:, :3], pn), F.cross_entropy(full[:, 3:], cl)]).cpu()
    assert relerr(got.numpy(), want.numpy()) < 2e-6


@pytest.mark.parametrize("mode", ["mean", "first"])
def test_segment_aggregate_bit_exact(ops, mode):
    rng = np.random.default_rng(3)
    ntoks, C = [37, 1, 300], 768
    seg, runs = [], []
    for n in ntoks:
        ids = np.sort(rng.integers(0, max(2, n // 3), n)).astype(np.int32)
        seg.append(ids); runs.append(len(oracle_ops.segment_runs(ids)) - 1)
    hidden = torch.randn(sum(ntoks) + 11, C)
    tok_row = torch.from_numpy(rng.permutation(sum(ntoks) + 11)[:sum(ntoks)].astype(np.int32))
    toff, dtoff = _dev_off(ntoks)
    K = sum(runs)
    status = torch.zeros(1, dtype=torch.int32).cuda()
    starts = ops.segment_starts(torch.from_numpy(np.concatenate(seg)).cuda(), dtoff, 3, K, status)
    out = ops.segment_reduce(hidden.cuda(), tok_row.cuda(), starts, K, ops.AGG_MEAN if mode == "mean" else ops.AGG_FIRST)
    assert int(status.item()) == 0
    want = np.concatenate([oracle_ops.segment_aggregate(hidden[tok_row[toff[b]:toff[b + 1]].long()].numpy(), seg[b], mode) for b in range(3)], 0)
    assert np.array_equal(out.cpu().numpy(), want)          # sequential fp32 sum + one divide: bit-exact
    # run-count mismatch is flagged like the reference's assert (BERTgrid_generator.py:233)
    ops.segment_starts(torch.from_numpy(np.concatenate(seg)).cuda(), dtoff, 3, K + 1, status)
    assert int(status.item()) == 1


def test_roi_align_matches_oracle(ops):
    rng = np.random.default_rng(4)
    B, Hf, Wf, C = 2, 32, 48, 256
    feat = torch.randn(B, C, Hf, Wf)
    counts = [9, 6]
    per = [_boxes(rng, c, Hf * 4, Wf * 4, tail=False) for c in counts]
    per[0][0] = [0, 0, Wf * 4 - 1, Hf * 4 - 1]     # full page -> large adaptive grid
    per[0][1] = [10, 10, 10, 10]                   # zero size -> clamped to 1
    per[1][0] = [Wf * 4 - 3, Hf * 4 - 3, Wf * 4 + 30, Hf * 4 + 30]   # runs off the map
    off, doff = _dev_off(counts)
    boxes = np.concatenate(per, 0)
    out, sg = ops.roi_align(feat.permute(0, 2, 3, 1).contiguous().cuda(), torch.from_numpy(boxes).cuda(), doff, 0.25, 7, want_grid=True)
    bidx = np.concatenate([np.full(c, b, np.int32) for b, c in enumerate(counts)])
    want, grids = oracle_ops.roi_align(feat.numpy(), boxes.astype(np.float32), bidx, 0.25, 7)
    assert np.array_equal(sg.cpu().numpy(), grids)                                   # sample grid: bit-exact
    assert relerr(out.permute(0, 3, 1, 2).cpu().numpy(), want) < 1e-5               # values: <= 1e-5 rel


def test_transform_kernels(ops):
    g = torch.Generator().manual_seed(5)
    mean, std = [0.9248, 0.9224, 0.9215], [0.1532, 0.1545, 0.1536]
    for (h, w), (oh, ow) in [((64, 96), (64, 96)), ((85, 107), (96, 120)), ((333, 777), (342, 800))]:
        img = torch.rand(3, h, w, generator=g)
        H, W = ((oh + 31) // 32) * 32, ((ow + 31) // 32) * 32
        batch = torch.zeros(1, H + 6, W + 6, 4).cuda()
        ops.normalize_resize_pad(img.cuda(), batch, 0, oh, ow, mean, std)
        assert float(batch[0, :3].abs().sum() + batch[0, -3:].abs().sum() + batch[0, :, :3].abs().sum() + batch[0, :, -3:].abs().sum()
                     + batch[..., 3].abs().sum()) == 0.0                      # border and 4th channel stay zero
        batch = batch[:, 3:-3, 3:-3, :3]
        x = (img - torch.tensor(mean)[:, None, None]) / torch.tensor(std)[:, None, None]
        x = F.interpolate(x[None], size=(oh, ow), mode="bilinear", align_corners=False)[0]
        want = torch.zeros(3, H, W); want[:, :oh, :ow] = x
        assert relerr(batch[0].permute(2, 0, 1).cpu().numpy(), want.numpy()) < 2e-6
    coors = torch.tensor([[100, 50, 300, 150], [0, 0, 776, 332], [5, 7, 9, 11]], dtype=torch.int64)
    nh, nw = oracle_ops.resized_shape(333, 777, oracle_ops.resize_scale(333, 777, 512, 800))
    ratios = torch.tensor([[nh / 333, nw / 777]], dtype=torch.float32)
    got = ops.resize_coords(coors.cuda(), torch.tensor([0, 3], dtype=torch.int32).cuda(), ratios.cuda(), 1)
    assert np.array_equal(got.cpu().numpy(), oracle_ops.resize_coords(coors.numpy(), (333, 777), (nh, nw)))
    assert got.cpu().numpy()[0].tolist() == [102, 51, 308, 154]                       # SURVEY A.1 known answer


@pytest.mark.parametrize("M,N,K,K1", [(300, 768, 768, 768), (4100, 3072, 768, 768), (128, 1024, 1792, 1024), (257, 5, 512, 512), (64, 2, 1024, 1024), (1000, 130, 36, 36)])
def test_gemm_fp32(ops, M, N, K, K1):
    g = torch.Generator().manual_seed(M + N)
    A = torch.randn(M, K, generator=g); Wt = torch.randn(N, K, generator=g) / K ** 0.5
    bias = torch.randn(N, generator=g); res = torch.randn(M, N, generator=g)
    want = F.gelu(A @ Wt.t() + bias + res)
    a1, a2 = A[:, :K1].contiguous().cuda(), (A[:, K1:].contiguous().cuda() if K1 < K else None)
    ep = ops.make_epilogue(None, bias.cuda(), res.cuda(), ops.RES_SAME, ldr=N, act=ops.ACT_GELU)
    got = ops.gemm(a1, Wt.cuda(), A2=a2, ep=ep, precision=ops.PREC_FP32)
    assert relerr(got.cpu().numpy(), want.numpy()) < 2e-6


def _trunc_tf32(x):
    """What kind::tf32 sees: the low 13 mantissa bits of each fp32 operand are ignored."""
    return (x.contiguous().view(torch.int32) & ~0x1FFF).view(torch.float32)


@pytest.mark.parametrize("M,N,K,K1", [(128, 64, 32, 32), (300, 768, 768, 768), (4100, 3072, 768, 768), (128, 1024, 1792, 1024),
                                      (1000, 136, 64, 64), (77, 64, 96, 32), (4096, 128, 896, 128), (257, 5, 512, 512)])
def test_gemm_tf32_tensor_core(ops, M, N, K, K1):
    """tcgen05 path: exact against a float64 product of TF32-truncated operands (proves the tiling /
    descriptors / epilogue), and within 2e-3 of the fp32 product (the precision the mode trades)."""
    if not ops.tc_available():
        pytest.fail("tcgen05 path unavailable on this GPU box: " + __import__("vibertgrid_pytorch_b200")._lib.last_error())
    g = torch.Generator().manual_seed(M * 7 + N)
    A = torch.randn(M, K, generator=g); Wt = torch.randn(N, K, generator=g) / K ** 0.5
    bias = torch.randn(N, generator=g); res = torch.randn(M, N, generator=g)
    a1, a2 = A[:, :K1].contiguous().cuda(), (A[:, K1:].contiguous().cuda() if K1 < K else None)
    ep = ops.make_epilogue(None, bias.cuda(), res.cuda(), ops.RES_SAME, ldr=N, act=ops.ACT_GELU)
    got = ops.gemm(a1, Wt.cuda(), A2=a2, ep=ep, precision=ops.PREC_TF32).cpu()
    want_t = F.gelu(_trunc_tf32(A).double() @ _trunc_tf32(Wt).double().t() + bias + res)
    want = F.gelu(A.double() @ Wt.double().t() + bias + res)
    if N >= 64:      # smaller N is served by the fp32 CUDA-core kernel
        assert relerr(got.numpy(), want_t.numpy()) < 5e-6
    assert relerr(got.numpy(), want.numpy()) < 2e-3


def _split_ref(x):
    """bf16x3 operand model: x ~ hi + lo with hi = bf16(x), lo = bf16(x - hi)."""
    hi = x.to(torch.bfloat16).float()
    return hi, (x - hi).to(torch.bfloat16).float()


def _bf16x3_product(A, Wt):
    a1, a2 = _split_ref(A); w1, w2 = _split_ref(Wt)
    return a1.double() @ w1.double().t() + a2.double() @ w1.double().t() + a1.double() @ w2.double().t()


@pytest.mark.parametrize("M,N,K,K1", [(128, 64, 64, 64), (300, 768, 768, 768), (4100, 3072, 768, 768), (128, 1024, 1792, 1024),
                                      (1000, 136, 64, 64), (77, 64, 192, 64), (4096, 128, 896, 128), (513, 256, 12544, 12544)])
def test_gemm_bf16x3_tensor_core(ops, M, N, K, K1):
    """Parity-grade tensor-core mode: three bf16 tcgen05 products on hi/lo splits.  (1) matches a float64 evaluation of
    exactly those three products (proves TMA / converter / descriptors / epilogue); (2) within 2e-5 of the fp32 product."""
    assert ops.tc_available(), "tcgen05 path unavailable on this GPU box"
    g = torch.Generator().manual_seed(M * 7 + N)
    A = torch.randn(M, K, generator=g); Wt = torch.randn(N, K, generator=g) / K ** 0.5
    bias = torch.randn(N, generator=g); res = torch.randn(M, N, generator=g)
    a1, a2 = A[:, :K1].contiguous().cuda(), (A[:, K1:].contiguous().cuda() if K1 < K else None)
    ep = ops.make_epilogue(None, bias.cuda(), res.cuda(), ops.RES_SAME, ldr=N, act=ops.ACT_GELU)
    Wd = Wt.cuda()
    got = ops.gemm(a1, Wd, A2=a2, ep=ep, precision=ops.PREC_BF16X3, W_split=ops.split_bf16(Wd)).cpu()
    want3 = F.gelu(_bf16x3_product(A, Wt) + bias + res)
    want = F.gelu(A.double() @ Wt.double().t() + bias + res)
    # fp32 accumulation in TMEM: the rounding of 3K partial products grows with K
    assert relerr(got.numpy(), want3.numpy()) < (5e-6 if K <= 3072 else 6e-5)
    assert relerr(got.numpy(), want.numpy()) < (2e-5 if K <= 3072 else 8e-5)


@pytest.mark.parametrize("prec", ["bf16x3", "tf32"])
@pytest.mark.parametrize("B,H,W,Cin,Cout,k,s,p", [(2, 32, 48, 64, 64, 3, 1, 1), (3, 16, 16, 512, 512, 3, 1, 1), (9, 7, 7, 256, 256, 3, 1, 1),
                                                 (2, 20, 12, 64, 128, 3, 1, 1), (1, 128, 128, 64, 256, 3, 1, 1), (2, 8, 8, 128, 64, 1, 1, 0),
                                                 (2, 32, 32, 64, 128, 3, 2, 1), (8, 64, 64, 128, 256, 3, 2, 1), (2, 16, 24, 64, 128, 1, 2, 0),
                                                 (1, 256, 256, 64, 128, 3, 2, 1)])
def test_conv2d_tensor_core(ops, prec, B, H, W, Cin, Cout, k, s, p):
    """Implicit-GEMM conv through 4-D TMA maps: padding by out-of-bounds zero fill, stride 2 by traversal stride."""
    assert ops.tc_available()
    g = torch.Generator().manual_seed(Cin + Cout + H + s)
    x = torch.randn(B, Cin, H, W, generator=g); w = torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5
    scale = torch.rand(Cout, generator=g) + 0.5; shift = torch.randn(Cout, generator=g)
    Ho, Wo = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
    res = torch.randn(B, Cout, Ho, Wo, generator=g)
    want = F.relu(F.conv2d(x.double(), w.double(), None, s, p) * scale[None, :, None, None] + shift[None, :, None, None] + res)
    ep = ops.make_epilogue(scale.cuda(), shift.cuda(), res.permute(0, 2, 3, 1).contiguous().cuda(), ops.RES_SAME, ldr=Cout, act=ops.ACT_RELU)
    w_ohwi = ops.repack_oihw_to_ohwi(w.cuda())
    kw = dict(precision=ops.PREC_BF16X3, W_split=ops.split_bf16(w_ohwi)) if prec == "bf16x3" else dict(precision=ops.PREC_TF32)
    got = ops.conv2d(x.permute(0, 2, 3, 1).contiguous().cuda(), w_ohwi, s, p, ep=ep, **kw)
    assert relerr(got.permute(0, 3, 1, 2).cpu().numpy(), want.numpy()) < (2e-5 if prec == "bf16x3" else 2e-3)


@pytest.mark.parametrize("prec", ["bf16x3", "fp32"])
@pytest.mark.parametrize("B,H,W", [(2, 64, 96), (1, 512, 512), (3, 32, 32)])
def test_stem_conv(ops, prec, B, H, W):
    """7x7/2 pad-3 stem over the zero-bordered NHWC4 batch: tensor-core GEMM over overlapping 32-float TMA windows."""
    g = torch.Generator().manual_seed(H + W)
    x = torch.randn(B, 3, H, W, generator=g); w = torch.randn(64, 3, 7, 7, generator=g) / 147 ** 0.5
    scale = torch.rand(64, generator=g) + 0.5; shift = torch.randn(64, generator=g)
    want = F.relu(F.conv2d(x.double(), w.double(), None, 2, 3) * scale[None, :, None, None] + shift[None, :, None, None])
    x4 = torch.zeros(B, H + 6, W + 6, 4); x4[:, 3:-3, 3:-3, :3] = x.permute(0, 2, 3, 1)
    w774, w256 = ops.stem_pack_weights(w.cuda())
    assert torch.equal(w774.cpu()[..., :3], w.permute(0, 2, 3, 1)) and float(w774[..., 3].abs().sum()) == 0
    ep = ops.make_epilogue(scale.cuda(), shift.cuda(), act=ops.ACT_RELU)
    kw = dict(precision=ops.PREC_BF16X3, W_split=ops.split_bf16(w256)) if prec == "bf16x3" else dict(precision=ops.PREC_FP32)
    got = ops.stem_conv(x4.cuda(), w774, ep=ep, **kw)
    assert relerr(got.permute(0, 3, 1, 2).cpu().numpy(), want.numpy()) < (2e-5 if prec == "bf16x3" else 2e-6)


@pytest.mark.parametrize("B,H,W,Cin,Cout,k,s,p", [(2, 32, 48, 64, 64, 3, 1, 1), (1, 64, 64, 3, 64, 7, 2, 3), (2, 16, 16, 128, 256, 3, 2, 1),
                                                 (2, 16, 24, 64, 128, 1, 2, 0), (9, 7, 7, 256, 256, 3, 1, 1)])
def test_conv2d_fp32(ops, B, H, W, Cin, Cout, k, s, p):
    g = torch.Generator().manual_seed(Cin + Cout + k)
    x = torch.randn(B, Cin, H, W, generator=g); w = torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5
    scale = torch.rand(Cout, generator=g) + 0.5; shift = torch.randn(Cout, generator=g)
    want = F.relu(F.conv2d(x, w, None, s, p) * scale[None, :, None, None] + shift[None, :, None, None])
    w_ohwi = ops.repack_oihw_to_ohwi(w.cuda())
    assert torch.equal(w_ohwi.cpu(), w.permute(0, 2, 3, 1).contiguous())
    ep = ops.make_epilogue(scale.cuda(), shift.cuda(), act=ops.ACT_RELU)
    got = ops.conv2d(x.permute(0, 2, 3, 1).contiguous().cuda(), w_ohwi, s, p, ep=ep, precision=ops.PREC_FP32)
    assert relerr(got.permute(0, 3, 1, 2).cpu().numpy(), want.numpy()) < 2e-6


def test_upsample_residual_epilogue_and_pools(ops):
    g = torch.Generator().manual_seed(7)
    B, H, W, Cin, N = 2, 8, 12, 64, 256
    x = torch.randn(B, H, W, Cin, generator=g); Wt = torch.randn(N, Cin, generator=g) / 8
    small = torch.randn(B, H // 2, W // 2, N, generator=g)
    want = x.reshape(-1, Cin) @ Wt.t() + small.repeat_interleave(2, 1).repeat_interleave(2, 2).reshape(-1, N)
    ep = ops.make_epilogue(residual=small.cuda(), res_mode=ops.RES_UP2, out_h=H, out_w=W)
    got = ops.gemm(x.reshape(-1, Cin).cuda(), Wt.cuda(), ep=ep)
    assert relerr(got.cpu().numpy(), want.numpy()) < 2e-6
    y = torch.randn(2, 17, 23, 64, generator=g)
    assert torch.equal(ops.maxpool3x3s2(y.cuda()).cpu(), F.max_pool2d(y.permute(0, 3, 1, 2), 3, 2, 1).permute(0, 2, 3, 1))
    z = torch.randn(2, 16, 24, 64, generator=g)
    assert relerr(ops.avgpool2x2(z.cuda()).cpu().numpy(), F.avg_pool2d(z.permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1).numpy()) < 1e-6
    lg = torch.randn(2, 6, 5, 8, generator=g)
    o1, o2 = ops.upsample_split_nchw(lg.cuda(), 4, 3)
    up = lg.permute(0, 3, 1, 2).repeat_interleave(4, 2).repeat_interleave(4, 3)
    assert torch.equal(o1.cpu(), up[:, :3]) and torch.equal(o2.cpu(), up[:, 3:])
    assert torch.equal(ops.nhwc_to_nchw(z.cuda()).cpu(), z.permute(0, 3, 1, 2))


def test_bert_kernels(ops):
    g = torch.Generator().manual_seed(8)
    hid, heads = 768, 12
    lens = [512, 4, 77, 130]
    cu = np.zeros(len(lens) + 1, np.int32); cu[1:] = np.cumsum(lens)
    R = int(cu[-1])
    qkv = torch.randn(R, 3 * hid, generator=g)
    got = ops.attention(qkv.cuda(), torch.from_numpy(cu).cuda(), len(lens), max(lens), heads)
    want = torch.empty(R, hid)
    for q in range(len(lens)):
        a, b = cu[q], cu[q + 1]
        Q, K, V = [qkv[a:b, i * hid:(i + 1) * hid].reshape(b - a, heads, 64).transpose(0, 1) for i in range(3)]
        want[a:b] = ((Q @ K.transpose(-1, -2) / 8).softmax(-1) @ V).transpose(0, 1).reshape(b - a, hid)
    assert relerr(got.cpu().numpy(), want.numpy()) < 5e-6
    x = torch.randn(333, hid, generator=g) * 3 + 1
    gam, bet = torch.rand(hid, generator=g) + 0.5, torch.randn(hid, generator=g)
    assert relerr(ops.layernorm(x.cuda(), gam.cuda(), bet.cuda(), 1e-12).cpu().numpy(), F.layer_norm(x, (hid,), gam, bet, 1e-12).numpy()) < 2e-6
    word, posw, typ = torch.randn(500, hid, generator=g), torch.randn(512, hid, generator=g), torch.randn(2, hid, generator=g)
    ids = torch.randint(0, 500, (97,), generator=g).int(); pos = torch.randint(0, 512, (97,), generator=g).int()
    got = ops.embed_ln(ids.cuda(), pos.cuda(), word.cuda(), posw.cuda(), typ[0].contiguous().cuda(), gam.cuda(), bet.cuda(), 1e-12)
    want = F.layer_norm(word[ids.long()] + typ[0] + posw[pos.long()], (hid,), gam, bet, 1e-12)
    assert relerr(got.cpu().numpy(), want.numpy()) < 2e-6
    lg = torch.randn(130, 5, generator=g) * 4
    assert relerr(ops.softmax_rows(lg.cuda()).cpu().numpy(), lg.softmax(1).numpy()) < 2e-6


@pytest.mark.parametrize("lens", [[512, 4, 77, 130], [200, 64, 65, 1, 128, 129, 300]])
def test_attention_tensor_core(ops, lens):
    """tcgen05 attention (S in TMEM, exact two-pass softmax, 3-term bf16 split) against a float64 softmax(QK^T/8)V."""
    assert ops.tc_available()
    g = torch.Generator().manual_seed(len(lens))
    hid, heads = 768, 12
    cu = np.zeros(len(lens) + 1, np.int32); cu[1:] = np.cumsum(lens)
    R = int(cu[-1])
    qkv = torch.randn(R, 3 * hid, generator=g) * 1.5
    got = ops.attention(qkv.cuda(), torch.from_numpy(cu).cuda(), len(lens), max(lens), heads, ops.PREC_BF16X3).cpu()
    want = torch.empty(R, hid, dtype=torch.float64)
    for q in range(len(lens)):
        a, b = cu[q], cu[q + 1]
        Q, K, V = [qkv[a:b, i * hid:(i + 1) * hid].double().reshape(b - a, heads, 64).transpose(0, 1) for i in range(3)]
        want[a:b] = ((Q @ K.transpose(-1, -2) / 8).softmax(-1) @ V).transpose(0, 1).reshape(b - a, hid)
    assert relerr(got.numpy(), want.numpy()) < 3e-5


@pytest.mark.parametrize("lens", [[512, 4, 77, 130], [200, 64, 65, 1, 128, 129, 300]])
def test_split_qkv_gemm_and_tma_attention(ops, lens):
    """QKV projection writing bf16 hi/lo planes (VBG_OUT_SPLIT_BF16) feeding the TMA-fed attention kernel (V consumed as an
    MN-major operand, no transpose) against float64."""
    assert ops.tc_available()
    g = torch.Generator().manual_seed(len(lens) + 11)
    hid, heads = 768, 12
    cu = np.zeros(len(lens) + 1, np.int32); cu[1:] = np.cumsum(lens)
    R = int(cu[-1])
    x = torch.randn(R, hid, generator=g); Wq = torch.randn(3 * hid, hid, generator=g) / hid ** 0.5 * 1.5
    bq = torch.randn(3 * hid, generator=g) * 0.1
    Wd = Wq.cuda()
    planes = ops.gemm(x.cuda(), Wd, ep=ops.make_epilogue(None, bq.cuda()), precision=ops.PREC_BF16X3, W_split=ops.split_bf16(Wd),
                      split_out=True)
    qkv = x.double() @ Wq.double().t() + bq
    rec = planes.t[0].float().double().cpu() + planes.t[1].float().double().cpu()
    assert planes.shape == (R, 3 * hid) and relerr(rec.numpy(), qkv.numpy()) < 2e-5
    got = ops.attention_split(planes, torch.from_numpy(cu).cuda(), len(lens), max(lens), heads).cpu()
    got_s = ops.attention_split(planes, torch.from_numpy(cu).cuda(), len(lens), max(lens), heads, split_out=True)
    assert relerr(got_s.float().cpu().numpy(), got.numpy()) < 2 ** -15          # same kernel, Split output
    want = torch.empty(R, hid, dtype=torch.float64)
    for q in range(len(lens)):
        a, b = cu[q], cu[q + 1]
        Q, K, V = [qkv[a:b, i * hid:(i + 1) * hid].reshape(b - a, heads, 64).transpose(0, 1) for i in range(3)]
        want[a:b] = ((Q @ K.transpose(-1, -2) / 8).softmax(-1) @ V).transpose(0, 1).reshape(b - a, hid)
    assert relerr(got.numpy(), want.numpy()) < 5e-5


def test_crf_viterbi_matches_oracle(ops):
    g = torch.Generator().manual_seed(9)
    T, counts = 7, [40, 1, 13]
    feats = torch.randn(sum(counts), T, generator=g)
    trans = torch.randn(T, T, generator=g); trans[T - 2, :] = -10000; trans[:, T - 1] = -10000
    off, doff = _dev_off(counts)
    tags, scores = ops.crf_viterbi(feats.cuda(), trans.cuda(), doff, 3)
    for b in range(3):
        s, path = oracle_ops.crf_viterbi(feats[off[b]:off[b + 1]].numpy(), trans.numpy(), T - 2, T - 1)
        assert tags[off[b]:off[b + 1]].cpu().numpy().astype(int).tolist() == path
        assert abs(float(scores[b]) - s) < 1e-3


# ------------------------------------------------------------------ pre-split activations (bf16 hi/lo planes end to end)
def _merged(x):
    """What a Split of x holds: hi + lo."""
    hi, lo = _split_ref(x)
    return hi + lo


@pytest.fixture(params=["64", "32"])
def ps_kb(request, monkeypatch):
    """Both ring-stage variants of the pre-split kernel: K=64 (SWIZZLE_128B) and K=32 (SWIZZLE_64B)."""
    from vibertgrid_pytorch_b200 import ops as _ops
    monkeypatch.setattr(_ops, "TUNE", (_ops.TUNE & ~_ops.TUNE_KB32) | (_ops.TUNE_KB32 if request.param == "32" else 0))
    return int(request.param)


@pytest.fixture(params=["0", "1"])
def ps_cg2(request, monkeypatch):
    """Single-CTA tiles (128 x BN) and CTA-pair tiles (tcgen05 cta_group::2, 256 x BN over a cluster of two)."""
    from vibertgrid_pytorch_b200 import ops as _ops
    keep = _ops.TUNE & ~(_ops.TUNE_PAIRS_OFF | _ops.TUNE_PAIRS_ON)
    monkeypatch.setattr(_ops, "TUNE", keep | (_ops.TUNE_PAIRS_ON if request.param == "1" else _ops.TUNE_PAIRS_OFF))
    return int(request.param)


def test_split_merge_roundtrip(ops):
    g = torch.Generator().manual_seed(5)
    x = torch.randn(1000, 64, generator=g) * 3
    s = ops.to_split(x.cuda())
    assert torch.equal(s.t[0].float().cpu(), _split_ref(x)[0]) and torch.equal(s.t[1].float().cpu(), _split_ref(x)[1])
    assert torch.equal(s.float().cpu(), _merged(x))
    assert relerr(s.float().cpu().numpy(), x.numpy()) < 2 ** -16


@pytest.mark.parametrize("M,N,K,K1", [(128, 64, 64, 64), (300, 768, 768, 768), (4128, 3072, 768, 768), (4128, 768, 3072, 3072),
                                      (128, 1024, 1792, 1024), (1000, 136, 64, 64), (77, 64, 192, 64), (4096, 128, 896, 128),
                                      (513, 256, 12544, 12544), (20000, 256, 256, 256),
                                      # M tails that leave a warp with both fully-valid and partially-valid lanes
                                      (94, 384, 128, 128), (222, 192, 64, 64)])
@pytest.mark.parametrize("out_split", [False, True])
def test_gemm_presplit(ops, ps_kb, ps_cg2, M, N, K, K1, out_split):
    """TMA-fed bf16x3 GEMM over Split operands (no in-kernel conversion): equals a float64 evaluation of the three
    products it issues; residual read from bf16 planes; fp32 or Split output."""
    assert ops.tc_available()
    g = torch.Generator().manual_seed(M * 7 + N + K)
    A = torch.randn(M, K, generator=g); Wt = torch.randn(N, K, generator=g) / K ** 0.5
    bias = torch.randn(N, generator=g); res = torch.randn(M, N, generator=g)
    a1 = ops.to_split(A[:, :K1].contiguous().cuda())
    a2 = ops.to_split(A[:, K1:].contiguous().cuda()) if K1 < K else None
    ep = ops.make_epilogue(None, bias.cuda(), ops.to_split(res.cuda()), ops.RES_SAME, ldr=N, act=ops.ACT_GELU)
    Wd = Wt.cuda()
    got = ops.gemm(a1, Wd, A2=a2, ep=ep, precision=ops.PREC_BF16X3, W_split=ops.split_bf16(Wd), split_out=out_split)
    want3 = F.gelu(_bf16x3_product(A, Wt) + bias + _merged(res))
    if out_split:
        assert isinstance(got, ops.Split) and got.shape == (M, N)
        assert relerr(got.float().cpu().numpy(), want3.numpy()) < (2e-5 if K <= 3072 else 7e-5)    # + 2^-17 storage rounding
    else:
        assert relerr(got.cpu().numpy(), want3.numpy()) < (5e-6 if K <= 768 else (1e-5 if K <= 3072 else 6e-5))
        assert relerr(got.cpu().numpy(), F.gelu(A.double() @ Wt.double().t() + bias + res).numpy()) < (3e-5 if K <= 3072 else 9e-5)


@pytest.mark.parametrize("B,H,W,Cin,Cout,k,s,p", [(2, 32, 48, 64, 64, 3, 1, 1), (3, 16, 16, 512, 512, 3, 1, 1), (9, 7, 7, 256, 256, 3, 1, 1),
                                                 (2, 20, 12, 64, 128, 3, 1, 1), (1, 128, 128, 64, 256, 3, 1, 1), (2, 32, 32, 64, 128, 3, 2, 1),
                                                 (8, 64, 64, 128, 256, 3, 2, 1), (2, 16, 24, 64, 128, 1, 2, 0), (1, 256, 256, 64, 128, 3, 2, 1)])
@pytest.mark.parametrize("res_mode", ["same", "up2"])
def test_conv2d_presplit(ops, ps_kb, ps_cg2, B, H, W, Cin, Cout, k, s, p, res_mode):
    """Implicit-GEMM conv over a Split NHWC activation through a rank-5 TMA map (c, w, h, b, plane)."""
    assert ops.tc_available()
    g = torch.Generator().manual_seed(Cin + Cout + H + s)
    x = torch.randn(B, Cin, H, W, generator=g); w = torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5
    scale = torch.rand(Cout, generator=g) + 0.5; shift = torch.randn(Cout, generator=g)
    Ho, Wo = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
    if res_mode == "up2" and (Ho % 2 or Wo % 2):
        pytest.skip("nearest-x2 residual needs even output dims")
    rs = (B, Cout, Ho, Wo) if res_mode == "same" else (B, Cout, Ho // 2, Wo // 2)
    res = torch.randn(*rs, generator=g)
    rfull = _merged(res) if res_mode == "same" else _merged(res).repeat_interleave(2, 2).repeat_interleave(2, 3)
    want = F.relu(F.conv2d(_merged(x).double(), _merged(w).double(), None, s, p) * scale[None, :, None, None]
                  + shift[None, :, None, None] + rfull)
    ep = ops.make_epilogue(scale.cuda(), shift.cuda(), ops.to_split(res.permute(0, 2, 3, 1).contiguous().cuda()),
                           ops.RES_SAME if res_mode == "same" else ops.RES_UP2, ldr=Cout, act=ops.ACT_RELU)
    w_ohwi = ops.repack_oihw_to_ohwi(w.cuda())
    xs = ops.to_split(x.permute(0, 2, 3, 1).contiguous().cuda())
    got = ops.conv2d(xs, w_ohwi, s, p, ep=ep, precision=ops.PREC_BF16X3, W_split=ops.split_bf16(w_ohwi), split_out=True)
    assert isinstance(got, ops.Split) and got.shape == (B, Ho, Wo, Cout)
    assert relerr(got.float().permute(0, 3, 1, 2).cpu().numpy(), want.numpy()) < 3e-5


def test_format_aware_memory_kernels(ops):
    """LayerNorm / pools / scatter / ROI-align writing (and reading) bf16 planes agree with their fp32 forms to the storage
    rounding of the Split format (2^-16 relative per element)."""
    g = torch.Generator().manual_seed(11)
    tol = 2 ** -15
    x = torch.randn(300, 768, generator=g).cuda(); gam = torch.randn(768, generator=g).cuda(); bet = torch.randn(768, generator=g).cuda()
    assert relerr(ops.layernorm(x, gam, bet, 1e-12, split=True).float().cpu().numpy(), ops.layernorm(x, gam, bet, 1e-12).cpu().numpy()) < tol
    a = torch.randn(2, 30, 44, 64, generator=g).cuda()
    mp = ops.maxpool3x3s2(a)
    assert torch.equal(ops.maxpool3x3s2(a, split_out=True).float(), ops.to_split(mp).float())
    sa = ops.to_split(a)
    assert relerr(ops.maxpool3x3s2(sa, split_out=True).float().cpu().numpy(), mp.cpu().numpy()) < tol
    assert relerr(ops.avgpool2x2(sa).float().cpu().numpy(), ops.avgpool2x2(a).cpu().numpy()) < tol
    # scatter
    rng = np.random.default_rng(3)
    H, W, counts = 128, 192, [33, 20]
    per = [_boxes(rng, c, H, W) for c in counts]
    off, doff = _dev_off(counts)
    boxes = torch.from_numpy(np.concatenate(per, 0)).cuda()
    emb = torch.randn(sum(counts), 768, generator=g).cuda()
    idx = ops.box_index_map(boxes, doff, 2, 8, H // 8, W // 8)
    assert torch.equal(ops.grid_scatter(emb, idx, doff, split=True).float(), ops.to_split(ops.grid_scatter(emb, idx, doff)).float())
    assert torch.equal(ops.grid_scatter(ops.to_split(emb), idx, doff).t, ops.to_split(ops.grid_scatter(emb, idx, doff)).t)   # plane copy
    # ROI align from / to planes: same sample grid, values to storage rounding
    feat = torch.randn(2, H // 4, W // 4, 256, generator=g).cuda()
    r32, g32 = ops.roi_align(feat, boxes, doff, 0.25, 7, want_grid=True)
    rs, gs = ops.roi_align(ops.to_split(feat), boxes, doff, 0.25, 7, want_grid=True, split_out=True)
    assert torch.equal(g32, gs)
    assert relerr(rs.float().cpu().numpy(), r32.cpu().numpy()) < 2 * tol


@pytest.mark.parametrize("M,N,K", [(2048, 512, 4608), (1024, 1024, 12544), (640, 256, 2304), (300, 512, 1024)])
def test_gemm_presplit_split_k(ops, monkeypatch, M, N, K):
    """Under-filled shapes run as (tile, K-range) work units + a deterministic finishing kernel: same result as the unsplit
    kernel up to fp32 re-association of the K sum, residual / activation / plane output applied once at the end."""
    from vibertgrid_pytorch_b200 import _lib
    assert _lib.load().vbg_gemm_ps_workspace(M, N, K, ops.TUNE_SPLITK) > 0, "shape expected to split"
    assert _lib.load().vbg_gemm_ps_workspace(M, N, K, 0) == 0, "split-K is opt-in"
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g); Wt = torch.randn(N, K, generator=g) / K ** 0.5
    bias = torch.randn(N, generator=g); res = torch.randn(M, N, generator=g)
    Wd = Wt.cuda(); Ws = ops.split_bf16(Wd); As = ops.to_split(A.cuda()); rs = ops.to_split(res.cuda())
    outs = {}
    for flag in ("0", "1"):
        monkeypatch.setattr(ops, "TUNE", ops.TUNE_SPLITK if flag == "1" else 0)
        mk = lambda: ops.make_epilogue(None, bias.cuda(), rs, ops.RES_SAME, ldr=N, act=ops.ACT_RELU)   # gemm(split_out) edits its ep
        outs[flag] = (ops.gemm(As, Wd, ep=mk(), precision=ops.PREC_BF16X3, W_split=Ws, split_out=True).float().cpu(),
                      ops.gemm(As, Wd, ep=mk(), precision=ops.PREC_BF16X3, W_split=Ws).cpu())
    want3 = F.relu(_bf16x3_product(A, Wt) + bias + _merged(res))
    for flag in ("0", "1"):
        assert relerr(outs[flag][1].numpy(), want3.numpy()) < (1e-5 if K <= 3072 else 6e-5), flag
        assert relerr(outs[flag][0].numpy(), want3.numpy()) < (3e-5 if K <= 3072 else 7e-5), flag
    again = ops.gemm(As, Wd, ep=ops.make_epilogue(None, bias.cuda(), rs, ops.RES_SAME, ldr=N, act=ops.ACT_RELU),
                     precision=ops.PREC_BF16X3, W_split=Ws).cpu()
    assert torch.equal(again, outs["1"][1]), "split-K must be deterministic"


def test_conv2d_presplit_split_k(ops, monkeypatch):
    B, H, W, Cin, Cout = 8, 16, 16, 512, 512
    from vibertgrid_pytorch_b200 import _lib
    assert _lib.load().vbg_conv2d_ps_workspace(B, H, W, Cin, Cout, 3, 3, 1, 1, ops.TUNE_SPLITK) > 0
    g = torch.Generator().manual_seed(99)
    x = torch.randn(B, Cin, H, W, generator=g); w = torch.randn(Cout, Cin, 3, 3, generator=g) / (Cin * 9) ** 0.5
    scale = torch.rand(Cout, generator=g) + 0.5; shift = torch.randn(Cout, generator=g); res = torch.randn(B, Cout, H, W, generator=g)
    want = F.relu(F.conv2d(_merged(x).double(), _merged(w).double(), None, 1, 1) * scale[None, :, None, None]
                  + shift[None, :, None, None] + _merged(res))
    w_ohwi = ops.repack_oihw_to_ohwi(w.cuda()); ws = ops.split_bf16(w_ohwi)
    xs = ops.to_split(x.permute(0, 2, 3, 1).contiguous().cuda()); rs = ops.to_split(res.permute(0, 2, 3, 1).contiguous().cuda())
    got = {}
    for flag in ("0", "1"):
        monkeypatch.setattr(ops, "TUNE", ops.TUNE_SPLITK if flag == "1" else 0)
        ep = ops.make_epilogue(scale.cuda(), shift.cuda(), rs, ops.RES_SAME, ldr=Cout, act=ops.ACT_RELU)
        got[flag] = ops.conv2d(xs, w_ohwi, 1, 1, ep=ep, precision=ops.PREC_BF16X3, W_split=ws, split_out=True).float().permute(0, 3, 1, 2).cpu()
        assert relerr(got[flag].numpy(), want.numpy()) < 3e-5
