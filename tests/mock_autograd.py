"""TEST-ONLY stand-ins for ``vibertgrid_pytorch_b200.autograd`` built from plain differentiable torch CPU ops with the same
call signatures and NHWC layouts (the counterpart of mock_ops.py for the training engine).

Purpose: run ``TrainEngine.loss`` + ``loss.backward()`` on a box without a GPU and hold the HOST wiring of the training step
(stage order, residual / upsample wiring, weight views, head variants, loss configurations) to the unmodified reference's
loss and gradients (tests/golden/train_*.npz).  Never imported by the product; the real kernels behind each Function are
verified one by one by the ``-m gpu`` tests.
"""
import numpy as np
import torch
import torch.nn.functional as F

from oracle import oracle_ops


def _nchw(x):
    return x.permute(0, 3, 1, 2)


def _nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def linear(x, weight, bias=None, out_planes=False):
    assert not out_planes, "the stand-ins have no plane format"
    return F.linear(x, weight, bias)


def pack_rows(*parts):
    return torch.cat(parts, 0)


def leaf_view(p, *shape):
    return p.view(*shape)


def _deferred(fn, ok, *keep):
    return fn()


class _Apply:
    def __init__(self, fn):
        self.apply = fn


def _conv(x, w, bias, stride, pad):
    return _nhwc(F.conv2d(_nchw(x), w, bias, stride=stride, padding=pad))


def _stem(x4, w):
    # zero-bordered NHWC4 batch [B, H+6, W+6, 4]: a plain 7x7 / stride-2 / pad-0 conv over channels 0..2
    return _nhwc(F.conv2d(_nchw(x4[..., :3]), w, None, stride=2, padding=0))


def _bn(x, gamma, beta, residual, relu, eps, stats, sync=None):
    Cc = x.shape[-1]
    x2 = x.reshape(-1, Cc)
    mean = x2.mean(0)
    var = x2.var(0, unbiased=False)
    y = (x2 - mean) * torch.rsqrt(var + eps) * gamma + beta
    y = y.view(x.shape)
    if residual is not None:
        y = y + residual
    if relu:
        y = torch.relu(y)
    stats.append((mean.detach(), var.detach(), x2.shape[0]))
    return y


def _attention(qkv, cu, nseq, max_len, heads, p_drop=0.0, seed=0, step_seed=None):
    assert p_drop == 0.0, "the stand-in has no attention dropout (tests zero it)"
    hid = qkv.shape[1] // 3
    d = hid // heads
    outs = []
    for q in range(nseq):
        a, b = int(cu[q]), int(cu[q + 1])
        Q, K, V = [qkv[a:b, i * hid:(i + 1) * hid].reshape(b - a, heads, d).transpose(0, 1) for i in range(3)]
        s = (Q @ K.transpose(-1, -2)) / (d ** 0.5)
        outs.append((s.softmax(-1) @ V).transpose(0, 1).reshape(b - a, hid))
    return torch.cat(outs, 0)


def _segment_reduce(hidden, tok_row, seg_start, K, mode):
    rows = hidden[tok_row.long()]
    out = []
    for k in range(K):
        a, b = int(seg_start[k]), int(seg_start[k + 1])
        out.append(rows[a] if mode == 1 else rows[a:b].sum(0) / (b - a))
    return torch.stack(out, 0)


def _grid_scatter(seg_emb, idx, boxes, seg_off, B, stride):
    so = seg_off.numpy()
    grids = []
    for b in range(B):
        rows = seg_emb[so[b]:so[b + 1]]
        i = idx[b].long()
        g = rows[i.clamp(min=0)] * (i >= 0).unsqueeze(-1).to(rows.dtype)
        grids.append(g)
    return torch.stack(grids, 0)


def _roi_align(feat, boxes, seg_off, scale, P):
    from torchvision.ops import roi_align
    so = seg_off.numpy()
    bidx = torch.cat([torch.full((int(so[b + 1] - so[b]),), float(b)) for b in range(feat.shape[0])])
    rois = torch.cat([bidx[:, None], boxes.float()], 1)
    return _nhwc(roi_align(_nchw(feat), rois, output_size=P, spatial_scale=scale, sampling_ratio=-1, aligned=False))


def _seg_ce(logits, boxes, seg_off, seg_cls, B, H, W, up, c_split):
    so = seg_off.numpy()
    split = lambda t: [t[so[b]:so[b + 1]].numpy() for b in range(B)]
    idx = oracle_ops.box_index_map(split(boxes), H, W, 1)
    pn, cl = oracle_ops.paint_labels(idx, split(seg_cls))
    full = _nchw(logits).repeat_interleave(up, 2).repeat_interleave(up, 3)
    return torch.stack([F.cross_entropy(full[:, :c_split], torch.from_numpy(pn)),
                        F.cross_entropy(full[:, c_split:], torch.from_numpy(cl))])


def _crf_nll(feats, trans, tags, seg_off, B):
    so = seg_off.numpy()
    T = trans.shape[0]
    return torch.stack([oracle_ops.crf_nll_torch(feats[so[b]:so[b + 1]], tags[so[b]:so[b + 1]], trans, T - 2, T - 1).float()
                        for b in range(B)])


ConvPS = _Apply(_conv)
StemF = _Apply(_stem)
BatchNormTrainF = _Apply(_bn)
MaxPoolF = _Apply(lambda x: _nhwc(F.max_pool2d(_nchw(x), 3, 2, 1)))
AvgPoolF = _Apply(lambda x: _nhwc(F.avg_pool2d(_nchw(x), 2, 2)))
Up2F = _Apply(lambda x: x.repeat_interleave(2, 1).repeat_interleave(2, 2))
GeluF = _Apply(F.gelu)
DropoutF = _Apply(lambda t, p, seed, step_seed=None: t if p == 0.0 else F.dropout(t, p, True))
LayerNormPS = _Apply(lambda x, g, b, eps: F.layer_norm(x, (x.shape[-1],), g, b, eps))
AttentionF = _Apply(_attention)
EmbedSumF = _Apply(lambda word, position, type_emb, ids, pos: word[ids.long()] + position[pos.long()] + type_emb[0])
SegmentReduceF = _Apply(_segment_reduce)
GridScatterF = _Apply(_grid_scatter)
RoiAlignF = _Apply(_roi_align)
SegCEF = _Apply(_seg_ce)
CrfNllF = _Apply(_crf_nll)
