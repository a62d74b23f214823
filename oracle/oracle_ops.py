"""ORACLE (test infrastructure only) -- integer / index stages of the ViBERTgrid
joint forward restated on the CPU with numpy and plain loops.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package.  The product
(``vibertgrid_pytorch_b200``) never does.

Pinning status: PINNED against the live reference -- ``oracle/make_golden.py``
imports the unmodified reference from /root/reference in the build container,
runs it on seeded inputs/weights and commits its outputs under
``tests/golden/``; ``tests/test_oracle_golden.py`` checks this restatement
against those fixtures.  (The reference itself ships no tests or golden
vectors)

Every function cites the reference lines it restates (paths relative to
/root/reference).
"""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple

import numpy as np


# ----------------------------------------------------------------------------- a1
def resize_scale(h: int, w: int, min_size: float, max_size: float) -> float:
    """pipeline/transform.py:137-144 -- scale so the short side hits ``min_size``
    unless that pushes the long side past ``max_size``."""
    lo, hi = float(min(h, w)), float(max(h, w))
    scale = float(min_size) / lo
    if hi * scale > float(max_size):
        scale = float(max_size) / hi
    return scale


def resized_shape(h: int, w: int, scale: float) -> Tuple[int, int]:
    """F.interpolate(recompute_scale_factor=True): out = floor(in * scale)
    (pipeline/transform.py:149-155)."""
    return int(math.floor(h * scale)), int(math.floor(w * scale))


def resize_coords(coor: np.ndarray, orig_hw, new_hw) -> np.ndarray:
    """pipeline/transform.py:163-169.  NOTE the axis swap that is part of the
    reference's behaviour: x columns (0,2) are scaled by the HEIGHT ratio and
    y columns (1,3) by the WIDTH ratio; then truncation to int32.
    Arithmetic is float32 like the reference's ``coor.float()``."""
    rh = new_hw[0] / orig_hw[0]
    rw = new_hw[1] / orig_hw[1]
    c = coor.astype(np.float32).copy()
    c[:, [0, 2]] *= np.float32(rh)
    c[:, [1, 3]] *= np.float32(rw)
    return np.trunc(c).astype(np.int32)


def padded_shape(shapes: Sequence[Tuple[int, int]], div: int = 32) -> Tuple[int, int]:
    """pipeline/transform.py:248-255 -- batch max, rounded up to ``div``."""
    H = max(s[0] for s in shapes)
    W = max(s[1] for s in shapes)
    return int(math.ceil(H / div) * div), int(math.ceil(W / div) * div)


# ----------------------------------------------------------------------------- a2
def bert_windows(corpus: np.ndarray, mask: np.ndarray):
    """model/BERTgrid_generator.py:81-146 -- 510-token windows, each framed by
    [CLS]=101 / [SEP]=102 and zero padded to 512.  Returns a list of
    (ids[B,512], attn_mask[B,512], slice_len)."""
    B, L = corpus.shape
    out = []
    start = 0
    for w in range(L // 510 + 1):
        end = (w + 1) * 510
        if end > L:
            seq, m = corpus[:, start:], mask[:, start:]
            pad = np.zeros((B, end - L), dtype=np.int64)
            ids = np.concatenate([np.full((B, 1), 101), seq, np.full((B, 1), 102), pad], 1)
            am = np.concatenate([np.ones((B, 1)), m, np.ones((B, 1)), pad], 1)
        else:
            seq, m = corpus[:, start:end], mask[:, start:end]
            ids = np.concatenate([np.full((B, 1), 101), seq, np.full((B, 1), 102)], 1)
            am = np.concatenate([np.ones((B, 1)), m, np.ones((B, 1))], 1)
        out.append((ids.astype(np.int64), am.astype(np.int64), seq.shape[1]))
        start = end
    return out


# ----------------------------------------------------------------------------- a3
def segment_runs(seg_indices: np.ndarray) -> np.ndarray:
    """model/BERTgrid_generator.py:162-182 -- a new segment starts whenever the id
    differs from the previous token's.  Returns run start offsets [S+1]."""
    n = seg_indices.shape[0]
    starts = [0] + [t for t in range(1, n) if seg_indices[t] != seg_indices[t - 1]] + [n]
    return np.asarray(starts, dtype=np.int32)


def segment_aggregate(tok: np.ndarray, seg_indices: np.ndarray, mode: str = "mean") -> np.ndarray:
    """model/BERTgrid_generator.py:156-188 -- ``mean``: sequential in-place fp32 sum
    then one divide; ``first``: first token of the run."""
    runs = segment_runs(seg_indices)
    out = np.zeros((len(runs) - 1, tok.shape[1]), dtype=np.float32)
    for s in range(len(runs) - 1):
        a, b = int(runs[s]), int(runs[s + 1])
        if mode == "first":
            out[s] = tok[a]
        else:
            acc = tok[a].astype(np.float32).copy()
            for t in range(a + 1, b):
                acc += tok[t]
            out[s] = acc / np.float32(b - a)
    return out


# ----------------------------------------------------------------------------- a4 / a6
def box_index_map(coors: List[np.ndarray], H: int, W: int, stride: int) -> np.ndarray:
    """model/BERTgrid_generator.py:230-243 (stride 8) and
    model/semantic_segmentation_head.py:199-214 (stride 1).

    Cell (y, x) of sample b holds the index (within the sample) of the LAST
    segment whose slice [y1:y2, x1:x2] covers it, or -1.  Slice bounds are
    ``int(c / stride)`` on int32 coords; numpy slicing reproduces Python slice
    clipping (and wrap-around for negatives) exactly as the reference's tensor
    slicing does."""
    B = len(coors)
    Hg, Wg = int(H / stride), int(W / stride)
    idx = np.full((B, Hg, Wg), -1, dtype=np.int32)
    for b in range(B):
        for s in range(coors[b].shape[0]):
            c = coors[b][s]
            if stride == 1:
                x1, y1, x2, y2 = int(c[0]), int(c[1]), int(c[2]), int(c[3])
            else:
                x1, y1, x2, y2 = (int(np.float32(c[0]) / np.float32(stride)), int(np.float32(c[1]) / np.float32(stride)),
                                  int(np.float32(c[2]) / np.float32(stride)), int(np.float32(c[3]) / np.float32(stride)))
            idx[b, y1:y2, x1:x2] = s
    return idx


def scatter_grid(seg_emb: List[np.ndarray], idx: np.ndarray) -> np.ndarray:
    """BERTgrid [B,C,Hg,Wg] (NCHW, zeros where idx == -1) -- BERTgrid_generator.py:220-243."""
    B, Hg, Wg = idx.shape
    C = seg_emb[0].shape[1]
    grid = np.zeros((B, C, Hg, Wg), dtype=np.float32)
    for b in range(B):
        m = idx[b] >= 0
        grid[b][:, m] = seg_emb[b][idx[b][m]].T
    return grid


def paint_labels(idx: np.ndarray, seg_classes: List[np.ndarray]):
    """semantic_segmentation_head.py:199-214 -- pos_neg: 1 if class>0 else 2 (0 = background);
    class map = class id.  int64 like the reference."""
    B = idx.shape[0]
    pos_neg = np.zeros(idx.shape, dtype=np.int64)
    cls = np.zeros(idx.shape, dtype=np.int64)
    for b in range(B):
        m = idx[b] >= 0
        c = seg_classes[b].astype(np.int64)[idx[b][m]]
        cls[b][m] = c
        pos_neg[b][m] = np.where(c > 0, 1, 2)
    return pos_neg, cls


# ----------------------------------------------------------------------------- a7
def roi_geometry(box: np.ndarray, scale: float, P: int):
    """torchvision roi_align (aligned=False, sampling_ratio=-1) geometry for one box,
    as called at model/grid_roi_align.py:37-41,81.  float32 arithmetic.
    Returns (start_w, start_h, bin_w, bin_h, grid_w, grid_h)."""
    f = np.float32
    sw, sh, ew, eh = f(box[0]) * f(scale), f(box[1]) * f(scale), f(box[2]) * f(scale), f(box[3]) * f(scale)
    rw = max(f(ew - sw), f(1.0))
    rh = max(f(eh - sh), f(1.0))
    bw, bh = f(rw / f(P)), f(rh / f(P))
    gh = int(math.ceil(f(rh / f(P))))
    gw = int(math.ceil(f(rw / f(P))))
    return sw, sh, bw, bh, gw, gh


def roi_align(feat: np.ndarray, boxes: np.ndarray, batch_idx: np.ndarray, scale: float, P: int = 7):
    """feat [B,C,H,W] NCHW fp32; boxes [K,4] float (x1,y1,x2,y2) image pixels.
    Returns out [K,C,P,P] and the integer sample-grid table [K,2] = (grid_h, grid_w).
    Restates torchvision's legacy ROIAlign."""
    f = np.float32
    B, C, H, W = feat.shape
    K = boxes.shape[0]
    out = np.zeros((K, C, P, P), dtype=np.float32)
    grids = np.zeros((K, 2), dtype=np.int32)
    for k in range(K):
        sw, sh, bw, bh, gw, gh = roi_geometry(boxes[k], scale, P)
        grids[k] = (gh, gw)
        fm = feat[int(batch_idx[k])]
        count = f(max(gh * gw, 1))
        for ph in range(P):
            for pw in range(P):
                acc = np.zeros(C, dtype=np.float32)
                for iy in range(gh):
                    y = f(sh + f(ph) * bh + f(f(iy) + f(0.5)) * bh / f(gh))
                    for ix in range(gw):
                        x = f(sw + f(pw) * bw + f(f(ix) + f(0.5)) * bw / f(gw))
                        if y < -1.0 or y > H or x < -1.0 or x > W:
                            continue
                        yy, xx = max(y, f(0)), max(x, f(0))
                        yl, xl = int(yy), int(xx)
                        if yl >= H - 1:
                            yh = yl = H - 1
                            yy = f(yl)
                        else:
                            yh = yl + 1
                        if xl >= W - 1:
                            xh = xl = W - 1
                            xx = f(xl)
                        else:
                            xh = xl + 1
                        ly, lx = f(yy - f(yl)), f(xx - f(xl))
                        hy, hx = f(f(1) - ly), f(f(1) - lx)
                        acc += (f(hy * hx) * fm[:, yl, xl] + f(hy * lx) * fm[:, yl, xh]
                                + f(ly * hx) * fm[:, yh, xl] + f(ly * lx) * fm[:, yh, xh])
                out[k, :, ph, pw] = acc / count
    return out, grids


# ----------------------------------------------------------------------------- a10
def crf_viterbi(feats: np.ndarray, trans: np.ndarray, start: int, stop: int):
    """model/crf.py:96-146 -- first-max tie-breaking like torch.max."""
    T = trans.shape[0]
    fv = np.full(T, -10000.0, dtype=np.float32)
    fv[start] = 0.0
    back = []
    for feat in feats:
        nv = fv[None, :] + trans            # [next, prev]
        bp = nv.argmax(1)
        fv = (nv[np.arange(T), bp] + feat).astype(np.float32)
        back.append(bp)
    term = fv + trans[stop]
    best = int(term.argmax())
    score = float(term[best])
    path = [best]
    for bp in reversed(back):
        best = int(bp[best])
        path.append(best)
    assert path.pop() == start
    path.reverse()
    return score, path


def crf_nll(feats: np.ndarray, tags: np.ndarray, trans: np.ndarray, start: int, stop: int) -> float:
    """model/crf.py:47-94,148-152 -- (log Z - gold score) / len."""
    T = trans.shape[0]
    fv = np.full(T, -10000.0, dtype=np.float64)
    fv[start] = 0.0
    for feat in feats:
        nv = fv[None, :] + trans.astype(np.float64) + feat.astype(np.float64)[:, None]
        m = nv.max(1)
        fv = m + np.log(np.exp(nv - m[:, None]).sum(1))
    term = fv + trans[stop]
    m = term.max()
    logz = m + math.log(np.exp(term - m).sum())
    gold, prev = 0.0, start
    for feat, t in zip(feats, tags):
        gold += float(trans[int(t), prev]) + float(feat[int(t)])
        prev = int(t)
    gold += float(trans[stop, prev])
    return (logz - gold) / max(len(feats), 1)


def crf_nll_torch(feats, tags, trans, start: int, stop: int):
    """Differentiable float64 restatement of model/crf.py:47-93,148-152 for ONE sequence (torch autograd supplies the
    gradients the reference's own backward produces).  feats [n, T], tags int [n], trans [T, T] (to <- from)."""
    import torch
    feats, trans = feats.double(), trans.double()
    T = trans.shape[0]
    fv = torch.full((T,), -10000.0, dtype=torch.float64)
    fv[start] = 0.0
    for feat in feats:
        fv = torch.logsumexp(fv[None, :] + trans, dim=1) + feat
    logz = torch.logsumexp(fv + trans[stop], dim=0)
    gold = feats.new_zeros(())
    prev = start
    for feat, t in zip(feats, tags.tolist()):
        gold = gold + trans[t, prev] + feat[t]
        prev = t
    gold = gold + trans[stop, prev]
    return (logz - gold) / feats.shape[0]
