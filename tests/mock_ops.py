"""TEST-ONLY stand-ins for ``vibertgrid_pytorch_b200.ops`` built from the oracle
(numpy) and plain torch CPU ops, with the same signatures and NHWC layouts.

Purpose: exercise the HOST logic of the forward engine (batch planning, weight
preparation, K-slicing of the fuse conv, residual/upsample wiring, head
variants) on a box without a GPU, against the golden fixtures.  Never imported
by the product; the real kernels are verified by the ``-m gpu`` tests.
"""
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn.functional as F

from oracle import oracle_ops

ACT_NONE, ACT_RELU, ACT_GELU = 0, 1, 2
RES_NONE, RES_SAME, RES_UP2 = 0, 1, 2
AGG_MEAN, AGG_FIRST = 0, 1
PREC_FP32, PREC_TF32, PREC_BF16X3 = 0, 1, 2


def tc_available():
    return False


def make_epilogue(scale=None, shift=None, residual=None, res_mode=RES_NONE, ldr=0, out_h=0, out_w=0, act=ACT_NONE):
    return SimpleNamespace(scale=scale, shift=shift, residual=residual, res_mode=res_mode, ldr=ldr, out_h=out_h,
                           out_w=out_w, act=act)


def _epilogue(y, ep, nhwc_shape=None):
    if ep is None:
        return y
    if ep.scale is not None:
        y = y * ep.scale
    if ep.shift is not None:
        y = y + ep.shift
    if ep.residual is not None:
        if ep.res_mode == RES_UP2:
            N = y.shape[-1]
            B = y.numel() // (ep.out_h * ep.out_w * N)
            r = ep.residual.reshape(B, ep.out_h // 2, ep.out_w // 2, N)
            r = r.repeat_interleave(2, 1).repeat_interleave(2, 2)
            y = y + r.reshape(y.shape)
        else:
            y = y + ep.residual.reshape(y.shape)
    if ep.act == ACT_RELU:
        y = F.relu(y)
    elif ep.act == ACT_GELU:
        y = F.gelu(y)
    return y


def normalize_resize_pad(img, batch, b, oh, ow, mean, std):
    m = torch.tensor(mean)[:, None, None]
    s = torch.tensor(std)[:, None, None]
    x = (img - m) / s
    if (oh, ow) != tuple(x.shape[-2:]):
        x = F.interpolate(x[None], size=(oh, ow), mode="bilinear", align_corners=False)[0]
    batch[b, 3:3 + oh, 3:3 + ow, :3] = x.permute(1, 2, 0)       # zero-bordered NHWC4 layout


def resize_coords(coors, seg_off, ratios, B):
    out = torch.empty((coors.shape[0], 4), dtype=torch.int32)
    so = seg_off.numpy()
    for b in range(B):
        c = coors[so[b]:so[b + 1]].numpy().astype(np.float32)
        c[:, [0, 2]] *= ratios[2 * b].numpy()
        c[:, [1, 3]] *= ratios[2 * b + 1].numpy()
        out[so[b]:so[b + 1]] = torch.from_numpy(np.trunc(c).astype(np.int32))
    return out


def bert_assemble(corpus, seq_tab, cu, nseq, R):
    ids = torch.zeros(R, dtype=torch.int32)
    pos = torch.zeros(R, dtype=torch.int32)
    st = seq_tab.reshape(-1, 4).numpy()
    for q in range(nseq):
        b, col0, n, sep = st[q]
        r0 = int(cu[q])
        ids[r0], pos[r0] = 101, 0
        ids[r0 + 1:r0 + 1 + n] = corpus[b, col0:col0 + n].int()
        pos[r0 + 1:r0 + 1 + n] = torch.arange(1, n + 1, dtype=torch.int32)
        ids[r0 + n + 1], pos[r0 + n + 1] = 102, int(sep)
    return ids, pos


def embed_ln(ids, pos, word, position, type0, gamma, beta, eps, split=False):
    x = word[ids.long()] + type0 + position[pos.long()]
    return F.layer_norm(x, x.shape[-1:], gamma, beta, eps)


def layernorm(x, gamma, beta, eps, out=None, split=False):
    return F.layer_norm(x, x.shape[-1:], gamma, beta, eps)


def attention(qkv, cu, nseq, max_len, heads, precision=0):
    R, th = qkv.shape
    hid = th // 3
    d = hid // heads
    out = torch.empty(R, hid)
    for q in range(nseq):
        a, b = int(cu[q]), int(cu[q + 1])
        Q, K, V = [qkv[a:b, i * hid:(i + 1) * hid].reshape(b - a, heads, d).transpose(0, 1) for i in range(3)]
        s = (Q @ K.transpose(-1, -2)) / (d ** 0.5)
        out[a:b] = (s.softmax(-1) @ V).transpose(0, 1).reshape(b - a, hid)
    return out


def mask_check(mask, tok_off, status):
    to = tok_off.numpy()
    for b in range(mask.shape[0]):
        n = int(to[b + 1] - to[b])
        want = torch.zeros(mask.shape[1], dtype=mask.dtype)
        want[:n] = 1
        if not torch.equal(mask[b].ne(0), want.ne(0)):
            status |= 4


def segment_starts(seg_ids, tok_off, B, K, status):
    starts = []
    to = tok_off.numpy()
    for b in range(B):
        r = oracle_ops.segment_runs(seg_ids[to[b]:to[b + 1]].numpy())
        starts += [int(v) + int(to[b]) for v in r[:-1]]
    starts.append(int(to[B]))
    if len(starts) != K + 1:
        status |= 1
    return torch.tensor(starts, dtype=torch.int32)


def segment_reduce(hidden, tok_row, seg_start, K, mode=AGG_MEAN):
    out = torch.empty(K, hidden.shape[1])
    rows = hidden[tok_row.long()]
    for k in range(K):
        a, b = int(seg_start[k]), int(seg_start[k + 1])
        if mode == AGG_FIRST:
            out[k] = rows[a]
        else:
            acc = rows[a].clone()
            for t in range(a + 1, b):
                acc += rows[t]
            out[k] = acc / (b - a)
    return out


def _split(boxes, seg_off, B):
    so = seg_off.numpy()
    return [boxes[so[b]:so[b + 1]].numpy() for b in range(B)]


def box_index_map(boxes, seg_off, B, stride, Hg, Wg):
    return torch.from_numpy(oracle_ops.box_index_map(_split(boxes, seg_off, B), Hg * stride, Wg * stride, stride))


def grid_scatter(seg_emb, idx, seg_off, split=False):
    B = idx.shape[0]
    so = seg_off.numpy()
    g = oracle_ops.scatter_grid([seg_emb[so[b]:so[b + 1]].numpy() for b in range(B)], idx.numpy())
    return torch.from_numpy(g).permute(0, 2, 3, 1).contiguous()


def label_paint(boxes, seg_off, seg_cls, B, H, W):
    idx = oracle_ops.box_index_map(_split(boxes, seg_off, B), H, W, 1)
    pn, cl = oracle_ops.paint_labels(idx, _split(seg_cls, seg_off, B))
    return torch.from_numpy(pn), torch.from_numpy(cl)


def seg_ce_loss(boxes, seg_off, seg_cls, lg, B, H, W, up, c_split):
    pn, cl = label_paint(boxes, seg_off, seg_cls, B, H, W)
    full = lg.permute(0, 3, 1, 2).repeat_interleave(up, 2).repeat_interleave(up, 3)
    return torch.stack([F.cross_entropy(full[:, :c_split], pn), F.cross_entropy(full[:, c_split:], cl)])


def gemm(A, W, *, A2=None, ep=None, precision=0, N=None, K=None, ldw=None, out=None, w_offset=0, W_split=None, split_out=False):
    X = A if A2 is None else torch.cat([A, A2], 1)
    Kt = X.shape[1] if K is None else K
    Nn = W.shape[0] if N is None else N
    ld = W.stride(0) if ldw is None else ldw
    Wsub = torch.as_strided(W.reshape(-1), (Nn, Kt), (ld, 1), w_offset)
    return _epilogue(X @ Wsub.t(), ep)


def conv2d(x, w_ohwi, stride, pad, *, ep=None, precision=0, W_split=None, split_out=False):
    y = F.conv2d(x.permute(0, 3, 1, 2), w_ohwi.permute(0, 3, 1, 2), None, stride, pad).permute(0, 2, 3, 1).contiguous()
    if ep is not None and ep.res_mode == RES_UP2:
        ep.out_h, ep.out_w = y.shape[1], y.shape[2]
    return _epilogue(y, ep)


def split_bf16(w):
    hi = w.detach().to(torch.bfloat16)
    return torch.stack([hi, (w.detach() - hi.float()).to(torch.bfloat16)], 0)


def stem_pack_weights(w):
    O = w.shape[0]
    w774 = torch.zeros(O, 7, 7, 4)
    w774[..., :3] = w.detach().permute(0, 2, 3, 1)
    w884 = torch.zeros(O, 8, 8, 4)
    w884[:, :7, :7] = w774
    return w774, w884.reshape(O, 256)


def stem_conv(x4, w774, *, ep=None, precision=0, W_split=None):
    # the 3-pixel border is already in x4: plain 7x7 / stride 2 / pad 0
    return conv2d(x4, w774, 2, 0, ep=ep)


def maxpool3x3s2(x, split_out=False):
    return F.max_pool2d(x.permute(0, 3, 1, 2), 3, 2, 1).permute(0, 2, 3, 1).contiguous()


def avgpool2x2(x, split_out=None):
    return F.avg_pool2d(x.permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1).contiguous()


def bn_fold(bn):
    a = bn.weight.detach() / torch.sqrt(bn.running_var + bn.eps)
    return a, bn.bias.detach() - bn.running_mean * a


def repack_oihw_to_ohwi(w):
    return w.detach().permute(0, 2, 3, 1).contiguous()


def roi_align(feat, boxes, seg_off, spatial_scale, P, want_grid=False, split_out=False):
    B = feat.shape[0]
    so = seg_off.numpy()
    bidx = np.concatenate([np.full(so[b + 1] - so[b], b, np.int32) for b in range(B)])
    out, grids = oracle_ops.roi_align(feat.permute(0, 3, 1, 2).numpy(), boxes.numpy().astype(np.float32), bidx,
                                      spatial_scale, P)
    out = torch.from_numpy(out).permute(0, 2, 3, 1).contiguous()
    return (out, torch.from_numpy(grids)) if want_grid else out


def softmax_rows(x):
    return x.softmax(1)


def full_head_scores(pn, cls):
    p = pn.sigmoid()
    return torch.cat([p[:, None], torch.where((p >= 0.5)[:, None], cls.sigmoid(), torch.zeros_like(cls))], 1)


def upsample_split_nchw(x, up, c_split):
    y = x.permute(0, 3, 1, 2).repeat_interleave(up, 2).repeat_interleave(up, 3)
    return y[:, :c_split].contiguous(), y[:, c_split:].contiguous()


def nhwc_to_nchw(x):
    return x.permute(0, 3, 1, 2).contiguous()


def crf_viterbi(feats, trans, seg_off, B):
    T = trans.shape[0]
    so = seg_off.numpy()
    tags, scores = [], []
    for b in range(B):
        s, path = oracle_ops.crf_viterbi(feats[so[b]:so[b + 1]].numpy(), trans.numpy(), T - 2, T - 1)
        tags += path
        scores.append(s)
    return torch.tensor(tags, dtype=torch.float32), torch.tensor(scores, dtype=torch.float32)


# ---- training-mode BatchNorm pieces (used by tests/test_syncbn_gloo.py to run the real autograd Function on CPU)
def bn_stats(x2d, eps):
    mean = x2d.mean(0)
    var = x2d.var(0, unbiased=False)
    return mean, var, torch.rsqrt(var + eps)


def bn_apply(x2d, mean, rstd, gamma, beta, residual=None, relu=False):
    y = (x2d - mean) * rstd * gamma + beta
    if residual is not None:
        y = y + residual
    return F.relu(y) if relu else y


def _masked(dy2d, y_relu):
    return dy2d if y_relu is None else dy2d * (y_relu > 0).to(dy2d.dtype)


def bn_bwd_reduce(x2d, dy2d, y_relu, mean, rstd):
    g = _masked(dy2d, y_relu)
    return (g * (x2d - mean) * rstd).sum(0), g.sum(0)


def bn_bwd_dx(x2d, dy2d, y_relu, mean, rstd, gamma, s_xhat, s_dy, count, want_dres=False):
    g = _masked(dy2d, y_relu)
    inv = 1.0 / float(count)
    dx = gamma * rstd * (g - s_dy * inv - (x2d - mean) * rstd * s_xhat * inv)
    return dx, (g.clone() if want_dres else None)


def bn_bwd(x2d, dy2d, y_relu, mean, rstd, gamma, want_dres=False):
    s_xhat, s_dy = bn_bwd_reduce(x2d, dy2d, y_relu, mean, rstd)
    dx, dres = bn_bwd_dx(x2d, dy2d, y_relu, mean, rstd, gamma, s_xhat, s_dy, x2d.shape[0], want_dres)
    return dx, dres, s_xhat, s_dy
