"""Python-side operator wrappers over the C-ABI (one function per ``vbg_*`` entry).

torch is used for device memory and the current stream only; every wrapper
passes raw device pointers and sizes to ``libvbg_sm100a.so``.  Activations are
fp32 NHWC.  Wrappers raise ``VbgError`` with the library's message on failure
and ``TypeError`` on a tensor that is not a contiguous CUDA tensor of the
expected dtype (no silent copies, no eager fallback).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib as L
from ._lib import (ACT_GELU, ACT_NONE, ACT_RELU, AGG_FIRST, AGG_MEAN, PREC_BF16X3, PREC_FP32, PREC_TF32, RES_NONE,
                   RES_SAME, RES_UP2, Epilogue, VbgError)


def _p(t: Optional[torch.Tensor], dtype=None, name="tensor"):
    if t is None:
        return None
    if not t.is_cuda:
        raise TypeError(f"{name} must be a CUDA tensor (the hot path has no CPU fallback)")
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise TypeError(f"{name} must be contiguous")
    return t.data_ptr()


# torch.cuda.current_stream() walks several Python layers (device-index resolution, availability checks, a Stream object):
# ~7 us per call, once per kernel launch -- a fifth of the host time of the launch-bound training step
# (profiles/r1_train_host_profile.log).  The raw accessors return the same cudaStream_t (they honour torch.cuda.stream(...)
# contexts and graph capture) in one C call each.
_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)
_raw_device = getattr(torch._C, "_cuda_getDevice", None)


def _stream():
    if _raw_stream is not None and _raw_device is not None:
        return _raw_stream(_raw_device())
    return torch.cuda.current_stream().cuda_stream


class Split:
    """An activation stored as two bf16 planes (include/vbg.h "storage formats"): ``t`` is a contiguous bf16
    ``[2, *shape]`` tensor, ``t[0] = bf16_rn(x)``, ``t[1] = bf16_rn(x - t[0])``.  The operand format the pre-split
    bf16x3 tensor-core kernels consume without conversion; ``shape`` / ``view`` / ``float`` mirror a tensor."""
    __slots__ = ("t",)

    def __init__(self, t: torch.Tensor):
        if t.dtype != torch.bfloat16 or t.dim() < 2 or t.shape[0] != 2 or not t.is_contiguous():
            raise TypeError("Split expects a contiguous bf16 [2, ...] tensor")
        if t[0].numel() % 8:
            raise ValueError("Split: elements per plane must be a multiple of 8 (TMA stride alignment)")
        self.t = t

    @staticmethod
    def empty(shape, device):
        return Split(torch.empty((2,) + tuple(shape), dtype=torch.bfloat16, device=device))

    @property
    def shape(self):
        return tuple(self.t.shape[1:])

    @property
    def device(self):
        return self.t.device

    @property
    def plane(self):
        return self.t[0].numel()

    def dim(self):
        return self.t.dim() - 1

    def view(self, *shape):
        return Split(self.t.view(2, *shape))

    def float(self):
        """fp32 copy (hi + lo) -- inspection and tests only."""
        out = torch.empty(self.shape, dtype=torch.float32, device=self.t.device)
        L.check(L.load().vbg_merge_bf16(_p(self.t[0]), _p(self.t[1]), self.plane, _f32(out), _stream()), "vbg_merge_bf16")
        return out


def to_split(x: torch.Tensor) -> Split:
    """fp32 tensor -> Split (one extra pass; producers normally write the planes themselves)."""
    return Split(split_bf16(x))


def as_f32(x):
    return x.float() if isinstance(x, Split) else x


def _act(x, name="tensor"):
    """(device pointer, plane) of an activation in either storage format; plane == 0 means fp32."""
    if isinstance(x, Split):
        return _p(x.t), x.plane
    return _f32(x, name), 0


def _workspace(nbytes, device):
    """Caller-owned scratch for the split-K form of an under-filled GEMM (the library never allocates); None when unused."""
    return torch.empty(int(nbytes), dtype=torch.uint8, device=device) if nbytes > 0 else None


def _new_act(shape, device, split):
    return Split.empty(shape, device) if split else torch.empty(tuple(shape), dtype=torch.float32, device=device)


def _f32(t, name="tensor"):
    return _p(t, torch.float32, name)


def _i32(t, name="tensor"):
    return _p(t, torch.int32, name)


def tc_available() -> bool:
    return bool(L.load().vbg_tc_available())


def make_epilogue(scale=None, shift=None, residual=None, res_mode=RES_NONE, ldr=0, out_h=0, out_w=0, act=ACT_NONE):
    """``residual`` may be an fp32 tensor or a Split (read from its bf16 planes by the tensor-core epilogues)."""
    rp, rplane = _act(residual, "residual") if residual is not None else (None, 0)
    ep = Epilogue(_f32(scale, "scale"), _f32(shift, "shift"), rp, res_mode, ldr, out_h, out_w, act)
    ep.res_plane = rplane
    ep._keep = (scale, shift, residual)
    return ep


# Variant overrides of the pre-split tensor-core kernels (include/vbg.h VBG_TUNE_*), passed with every call in the epilogue
# descriptor: 0 (what ships) = the library's own heuristics.  Sweeps and parity tests set ops.TUNE; the library itself
# reads no environment variable.
TUNE_KB32, TUNE_PAIRS_OFF, TUNE_PAIRS_ON, TUNE_SPLITK, TUNE_NO_PDL = 1, 2, 4, 8, 16
TUNE = 0


def _tuned(ep):
    if TUNE == 0 and (ep is None or ep.tune == 0):
        return ep
    if ep is None:
        ep = make_epilogue()
    ep.tune = TUNE
    return ep


# ------------------------------------------------------------------ a1
def image_hw(img):
    """(h, w) of one input image: ToTensor's float32 [3, h, w] (the reference's dataset output) or the decoded uint8 [h, w, 3]
    pixels of the shard loader (shards.py)."""
    if img.dtype == torch.uint8:
        if img.dim() != 3 or img.shape[2] != 3:
            raise TypeError(f"uint8 images are decoded pixels [h, w, 3]; got {tuple(img.shape)}")
        return int(img.shape[0]), int(img.shape[1])
    if img.dim() != 3 or img.shape[0] != 3:
        raise TypeError(f"float images are [3, h, w]; got {tuple(img.shape)}")
    return int(img.shape[1]), int(img.shape[2])


def _u8(t, what="image"):
    if t.dtype != torch.uint8 or not t.is_contiguous() or t.device.type != "cuda":
        raise TypeError(f"{what}: contiguous uint8 CUDA tensor required (no CPU fallback)")
    return t.data_ptr()


def normalize_resize_pad(img_chw, batch_nhwc4, b, oh, ow, mean, std):
    """``batch_nhwc4`` is the zero-initialised [B, H+6, W+6, 4] stem input (3-pixel border, 4th channel 0).  ``img_chw`` may
    also be a decoded uint8 [h, w, 3] image (the kernel then applies ToTensor's ``/ 255`` itself)."""
    B, Hp, Wp, c4 = batch_nhwc4.shape
    assert c4 == 4
    H, W = Hp - 6, Wp - 6
    m = (C.c_float * 3)(*mean)
    s = (C.c_float * 3)(*std)
    if img_chw.dtype == torch.uint8:
        h, w = image_hw(img_chw)
        L.check(L.load().vbg_normalize_resize_pad_u8(_u8(img_chw), 1, h, w, _f32(batch_nhwc4), b, H, W, oh, ow, m, s, _stream()),
                "vbg_normalize_resize_pad_u8")
        return
    _, h, w = img_chw.shape
    batch_nhwc = batch_nhwc4
    L.check(L.load().vbg_normalize_resize_pad(_f32(img_chw, "image"), h, w, _f32(batch_nhwc), b, H, W, oh, ow, m, s,
                                              _stream()), "vbg_normalize_resize_pad")


def image_table(images, sizes):
    """Device table [B, 6] int32 for ``decode_batch_u8``: per image the signed 64-bit byte offset of its pixels from the FIRST
    image's (low word, high word -- any two device addresses work, the images need not share an allocation), h, w and the
    resize target (oh, ow).  Staged through pinned memory with an asynchronous copy; build it OUTSIDE a graph capture."""
    import numpy as np
    base = _u8(images[0])
    rows = []
    for im, (oh, ow) in zip(images, sizes):
        h, w = image_hw(im)
        off = (_u8(im) - base) & 0xFFFFFFFFFFFFFFFF
        rows.append([off & 0xFFFFFFFF, off >> 32, h, w, int(oh), int(ow)])
    host = torch.from_numpy(np.asarray(rows, dtype=np.int64).astype(np.uint32).view(np.int32)).pin_memory()
    return host.to(images[0].device, non_blocking=True)


def decode_batch_u8(images, tab, batch_nhwc4, sizes, mean, std):
    """All uint8 [h, w, 3] images of a batch (any sizes) -> samples 0..B-1 of the padded batch in ONE launch (``tab`` from
    ``image_table`` over the same tensors)."""
    B, Hp, Wp, c4 = batch_nhwc4.shape
    assert c4 == 4 and len(images) == B == len(sizes) == tab.shape[0]
    m = (C.c_float * 3)(*mean)
    s = (C.c_float * 3)(*std)
    L.check(L.load().vbg_decode_batch_u8(_u8(images[0]), _i32(tab), B, max(int(o[0]) for o in sizes), max(int(o[1]) for o in sizes),
                                         _f32(batch_nhwc4), Hp - 6, Wp - 6, m, s, _stream()), "vbg_decode_batch_u8")


def normalize_resize_pad_batch(imgs_nchw, batch_nhwc4, b0, oh, ow, mean, std):
    """n same-shape images [n,3,h,w] (or uint8 [n,h,w,3]) -> samples b0..b0+n-1 of the padded batch, one launch."""
    B, Hp, Wp, c4 = batch_nhwc4.shape
    m = (C.c_float * 3)(*mean)
    s = (C.c_float * 3)(*std)
    if imgs_nchw.dtype == torch.uint8:
        n, h, w, _ = imgs_nchw.shape
        assert c4 == 4 and b0 + n <= B
        L.check(L.load().vbg_normalize_resize_pad_u8(_u8(imgs_nchw, "images"), n, h, w, _f32(batch_nhwc4), b0, Hp - 6, Wp - 6, oh, ow,
                                                     m, s, _stream()), "vbg_normalize_resize_pad_u8")
        return
    n, _, h, w = imgs_nchw.shape
    assert c4 == 4 and b0 + n <= B
    L.check(L.load().vbg_normalize_resize_pad_batch(_f32(imgs_nchw, "images"), n, h, w, _f32(batch_nhwc4), b0, Hp - 6, Wp - 6, oh, ow,
                                                    m, s, _stream()), "vbg_normalize_resize_pad_batch")


def resize_coords(coors_i64, seg_off, ratios, B):
    K = coors_i64.shape[0]
    out = torch.empty((K, 4), dtype=torch.int32, device=coors_i64.device)
    L.check(L.load().vbg_resize_coords(_p(coors_i64, torch.int64, "coors"), _i32(seg_off), _f32(ratios), B, K,
                                       _i32(out), _stream()), "vbg_resize_coords")
    return out


# ------------------------------------------------------------------ a2
def bert_assemble(corpus, seq_tab, cu, nseq, R):
    ids = torch.empty(R, dtype=torch.int32, device=corpus.device)
    pos = torch.empty(R, dtype=torch.int32, device=corpus.device)
    L.check(L.load().vbg_bert_assemble(_p(corpus, torch.int64, "corpus"), corpus.shape[1], _i32(seq_tab), _i32(cu),
                                       nseq, R, _i32(ids), _i32(pos), _stream()), "vbg_bert_assemble")
    return ids, pos


def embed_ln(ids, pos, word, position, type_emb, gamma, beta, eps, split=False):
    R, hidden = ids.shape[0], word.shape[1]
    out = _new_act((R, hidden), word.device, split)
    op, oplane = _act(out)
    L.check(L.load().vbg_embed_ln_x(_i32(ids), _i32(pos), _f32(word), _f32(position), _f32(type_emb), _f32(gamma),
                                    _f32(beta), eps, R, hidden, word.shape[0], position.shape[0], op, oplane, _stream()),
            "vbg_embed_ln")
    return out


def layernorm(x, gamma, beta, eps, out=None, split=False):
    """``split``: write bf16 hi/lo planes (a Split) instead of fp32; ``out`` may be the fp32 input (in place)."""
    R, hidden = x.shape
    if out is None:
        out = _new_act((R, hidden), x.device, split)
    op, oplane = _act(out)
    L.check(L.load().vbg_layernorm_x(_f32(x), _f32(gamma), _f32(beta), eps, R, hidden, op, oplane, _stream()), "vbg_layernorm")
    return out


def attention(qkv, cu, nseq, max_len, heads, precision=PREC_FP32):
    R, three_h = qkv.shape
    hidden = three_h // 3
    out = torch.empty((R, hidden), dtype=torch.float32, device=qkv.device)
    L.check(L.load().vbg_attention_fwd(_f32(qkv), _i32(cu), nseq, max_len, heads, hidden // heads, _f32(out), precision,
                                       _stream()), "vbg_attention_fwd")
    return out


def attention_split(qkv_split, cu, nseq, max_len, heads, split_out=False):
    """Attention over the Split [R, 3*hidden] of ``gemm(..., split_out=True)`` -> [R, hidden] (fp32, or a Split)."""
    if not isinstance(qkv_split, Split):
        qkv_split = Split(qkv_split)
    if qkv_split.dim() != 2:
        raise TypeError("attention_split expects the Split [R, 3*hidden] of gemm(..., split_out=True)")
    R, three_h = qkv_split.shape
    hidden = three_h // 3
    out = _new_act((R, hidden), qkv_split.device, split_out)
    op, oplane = _act(out)
    L.check(L.load().vbg_attention_split_fwd(_p(qkv_split.t), R * three_h, _i32(cu), nseq, R, max_len, heads, hidden // heads,
                                             op, oplane, _stream()), "vbg_attention_split_fwd")
    return out


def _u64ptr(t):
    """Device pointer of an optional int64[1] step-seed tensor (None -> NULL)."""
    return None if t is None else _p(t, torch.int64)


def attention_split_train(qkv_split, cu, nseq, max_len, heads, p_drop=0.0, seed=0, step_seed=None):
    """Training-mode attention: like ``attention_split`` (fp32 result) plus dropout on the attention probabilities (mask =
    pure function of ``seed``) and the base-2 row log-sum-exp [R, heads] the backward rebuilds the probabilities from."""
    if not isinstance(qkv_split, Split):
        qkv_split = Split(qkv_split)
    R, three_h = qkv_split.shape
    hidden = three_h // 3
    out = torch.empty((R, hidden), dtype=torch.float32, device=qkv_split.device)
    lse2 = torch.empty((R, heads), dtype=torch.float32, device=qkv_split.device)
    L.check(L.load().vbg_attention_split_train_fwd(_p(qkv_split.t), R * three_h, _i32(cu), nseq, R, max_len, heads, hidden // heads,
                                                   _f32(out), 0, _f32(lse2), float(p_drop), int(seed), _u64ptr(step_seed), _stream()),
            "vbg_attention_split_train_fwd")
    return out, lse2


def attention_bwd_tc(qkv_split, out, d_out, lse2, cu, nseq, max_len, heads, p_drop=0.0, seed=0, step_seed=None):
    """dQKV fp32 [R, 3*hidden] on the tensor cores from the QKV planes, fp32 O / dO and the forward's ``lse2`` (same seed)."""
    if not isinstance(qkv_split, Split):
        qkv_split = Split(qkv_split)
    R, three_h = qkv_split.shape
    hidden = three_h // 3
    dos = to_split(d_out)
    # rows past a sequence's end are never written by the kernels (every packed row belongs to a sequence: all rows are)
    dqkv = torch.empty((R, three_h), dtype=torch.float32, device=out.device)
    ws = torch.empty(R * heads, dtype=torch.float32, device=out.device)
    L.check(L.load().vbg_attention_bwd_tc(_p(qkv_split.t), R * three_h, _p(dos.t), R * hidden, _f32(out), _f32(d_out), _f32(lse2),
                                          _i32(cu), nseq, R, max_len, heads, hidden // heads, float(p_drop), int(seed), _u64ptr(step_seed),
                                          _f32(dqkv), _f32(ws), ws.numel() * 4, _stream()), "vbg_attention_bwd_tc")
    return dqkv


def attention_dropout_mask(seed, p_drop, row0, length, head, device, step_seed=None):
    """(keep mask [len, len] fp32, 1 / (1 - p_effective)) of the attention dropout for one (sequence, head) -- tests."""
    import ctypes
    mask = torch.empty((length, length), dtype=torch.float32, device=device)
    ik = ctypes.c_float(0.0)
    L.check(L.load().vbg_attention_dropout_mask(int(seed), int(step_seed or 0), int(step_seed is not None), float(p_drop), int(row0),
                                                int(length), int(head), _f32(mask), ctypes.byref(ik), _stream()),
            "vbg_attention_dropout_mask")
    return mask, float(ik.value)


# ------------------------------------------------------------------ a3
def mask_check(mask, tok_off, status):
    """Raises bit 2 of ``status`` when ``mask`` [B, L] is not the prefix mask the packed layout assumes (no host sync)."""
    B, Lm = mask.shape
    L.check(L.load().vbg_mask_check(_i32(mask), B, Lm, _i32(tok_off), _i32(status), _stream()), "vbg_mask_check")


def segment_starts(seg_ids, tok_off, B, K, status):
    n_tok = seg_ids.shape[0]
    out = torch.empty(K + 1, dtype=torch.int32, device=seg_ids.device)
    L.check(L.load().vbg_segment_starts(_i32(seg_ids), _i32(tok_off), B, n_tok, K, _i32(out), _i32(status), _stream()),
            "vbg_segment_starts")
    return out


def segment_reduce(hidden, tok_row, seg_start, K, mode=AGG_MEAN):
    Cc = hidden.shape[1]
    out = torch.empty((K, Cc), dtype=torch.float32, device=hidden.device)
    L.check(L.load().vbg_segment_reduce(_f32(hidden), _i32(tok_row), _i32(seg_start), K, Cc, mode, _f32(out), _stream()),
            "vbg_segment_reduce")
    return out


# ------------------------------------------------------------------ a4 / a6
def box_index_map(boxes, seg_off, B, stride, Hg, Wg):
    idx = torch.empty((B, Hg, Wg), dtype=torch.int32, device=boxes.device)
    L.check(L.load().vbg_box_index_map(_i32(boxes), _i32(seg_off), B, stride, Hg, Wg, _i32(idx), _stream()),
            "vbg_box_index_map")
    return idx


def grid_scatter(seg_emb, idx, seg_off, split=False, out=None):
    B, Hg, Wg = idx.shape
    Cc = seg_emb.shape[1]
    split = split or isinstance(seg_emb, Split)         # a Split source is copied plane-wise into a Split grid
    grid = out if out is not None else _new_act((B, Hg, Wg, Cc), seg_emb.device, split)
    (sp, splane), (gp, gplane) = _act(seg_emb), _act(grid)
    L.check(L.load().vbg_grid_scatter_x(sp, splane, _i32(idx), _i32(seg_off), B, Hg * Wg, Cc, gp, gplane, _stream()),
            "vbg_grid_scatter")
    return grid


def label_paint(boxes, seg_off, seg_cls, B, H, W):
    pn = torch.empty((B, H, W), dtype=torch.int64, device=boxes.device)
    cl = torch.empty((B, H, W), dtype=torch.int64, device=boxes.device)
    L.check(L.load().vbg_label_paint(_i32(boxes), _i32(seg_off), _i32(seg_cls), B, H, W, _p(pn), _p(cl), _stream()),
            "vbg_label_paint")
    return pn, cl


def seg_ce_loss(boxes, seg_off, seg_cls, logits_lowres, B, H, W, up, c_split):
    """(mean CE of the 3-way mask head, mean CE of the class head) over all pixels, from the low-res NHWC logits; fp32 [2]."""
    Ct = logits_lowres.shape[-1]
    nblk = ((W + 31) // 32) * ((H + 7) // 8) * B
    ws = torch.empty(2 * nblk, dtype=torch.float32, device=boxes.device)
    out = torch.empty(2, dtype=torch.float32, device=boxes.device)
    L.check(L.load().vbg_seg_ce_loss(_i32(boxes), _i32(seg_off), _i32(seg_cls), _f32(logits_lowres), B, H, W, up, Ct, c_split,
                                     _f32(ws), ws.numel() * 4, _f32(out), _stream()), "vbg_seg_ce_loss")
    return out


# ------------------------------------------------------------------ dense contractions
def split_bf16(w):
    """fp32 tensor -> bf16 [2, *w.shape]: plane 0 = bf16_rn(w), plane 1 = bf16_rn(w - plane 0)  (VBG_PREC_BF16X3 weights)."""
    w = w.detach().contiguous()
    out = torch.empty((2,) + tuple(w.shape), dtype=torch.bfloat16, device=w.device)
    L.check(L.load().vbg_split_bf16(_f32(w, "w"), w.numel(), _p(out[0]), _p(out[1]), _stream()), "vbg_split_bf16")
    return out


def _split_args(W_split, w_offset):
    if W_split is None:
        return None, 0
    if W_split.dtype != torch.bfloat16 or not W_split.is_contiguous() or W_split.shape[0] != 2:
        raise TypeError("W_split must be the contiguous bf16 [2, ...] tensor returned by split_bf16")
    return _p(W_split) + 2 * w_offset, W_split[0].numel()


def gemm(A, W, *, A2=None, ep: Optional[Epilogue] = None, precision=PREC_FP32, N=None, K=None, ldw=None, out=None,
         w_offset=0, W_split=None, split_out=False):
    """C[M,N] = epilogue([A | A2] @ W[N,K]^T).  ``w_offset``/``ldw``/``N``/``K`` select a sub-block of W (and W_split).
    ``split_out``: return bf16 [2, M, N] hi/lo planes instead of fp32 (tensor-core paths only)."""
    M, K1 = A.shape
    K2 = 0 if A2 is None else A2.shape[1]
    Kt = K1 + K2 if K is None else K
    Nn = W.shape[0] if N is None else N
    ldw_ = W.stride(0) if ldw is None else ldw
    if split_out:
        if ep is None:
            ep = make_epilogue()
        out = Split.empty((M, Nn), A.device)
        ep.out_mode, ep.out_plane = L.OUT_SPLIT_BF16, out.plane
        optr, ldc = _p(out.t), Nn
    else:
        if out is None:
            out = torch.empty((M, Nn), dtype=torch.float32, device=A.device)
        optr, ldc = _f32(out, "out"), out.stride(0)
    sp, plane = _split_args(W_split, w_offset)
    if isinstance(A, Split):      # pre-split activations: TMA-fed bf16x3 kernel, no in-kernel conversion
        if sp is None:
            raise TypeError("gemm over a Split activation needs W_split (the bf16 planes of the weight)")
        if A2 is not None and not isinstance(A2, Split):
            raise TypeError("gemm: A and A2 must use the same storage format")
        ap, aplane = _act(A)
        a2p, a2plane = _act(A2) if A2 is not None else (None, 0)
        ep = _tuned(ep)
        ws = _workspace(L.load().vbg_gemm_ps_workspace(M, Nn, Kt, TUNE), A.device)
        L.check(L.load().vbg_gemm_ps(ap, aplane, K1, a2p, a2plane, K2, K1, sp, plane, ldw_, optr, ldc, M, Nn, Kt,
                                     C.byref(ep) if ep is not None else None, _p(ws), 0 if ws is None else ws.numel(), _stream()),
                "vbg_gemm_ps")
        return out
    wp = _f32(W, "W") + 4 * w_offset
    L.check(L.load().vbg_gemm(_f32(A, "A"), A.stride(0), _f32(A2, "A2"), 0 if A2 is None else A2.stride(0), K1, wp, ldw_,
                              sp, plane, optr, ldc, M, Nn, Kt,
                              C.byref(ep) if ep is not None else None, precision, _stream()), "vbg_gemm")
    return out


def conv2d(x, w_ohwi, stride, pad, *, ep: Optional[Epilogue] = None, precision=PREC_FP32, W_split=None, split_out=False):
    """``x`` fp32 NHWC tensor or Split; ``split_out`` returns a Split (tensor-core paths only)."""
    B, H, W, Cin = x.shape
    Cout, kh, kw, Cin2 = w_ohwi.shape
    if Cin2 != Cin:
        raise ValueError(f"conv2d: weight expects Cin={Cin2}, input has {Cin}")
    Ho, Wo = (H + 2 * pad - kh) // stride + 1, (W + 2 * pad - kw) // stride + 1
    sp, plane = _split_args(W_split, 0)
    if split_out:
        if ep is None:
            ep = make_epilogue()
        y = Split.empty((B, Ho, Wo, Cout), x.device)
        ep.out_mode, ep.out_plane = L.OUT_SPLIT_BF16, y.plane
        yp = _p(y.t)
    else:
        y = torch.empty((B, Ho, Wo, Cout), dtype=torch.float32, device=x.device)
        yp = _f32(y)
    if isinstance(x, Split):
        if sp is None:
            raise TypeError("conv2d over a Split activation needs W_split")
        xp, xplane = _act(x)
        ep = _tuned(ep)
        ws = _workspace(L.load().vbg_conv2d_ps_workspace(B, H, W, Cin, Cout, kh, kw, stride, pad, TUNE), x.device)
        L.check(L.load().vbg_conv2d_ps(xp, xplane, B, H, W, Cin, sp, plane, Cout, kh, kw, stride, pad, yp,
                                       C.byref(ep) if ep is not None else None, _p(ws), 0 if ws is None else ws.numel(), _stream()),
                "vbg_conv2d_ps")
        return y
    L.check(L.load().vbg_conv2d(_f32(x, "x"), B, H, W, Cin, _f32(w_ohwi, "w"), sp, plane, Cout, kh, kw, stride, pad, yp,
                                C.byref(ep) if ep is not None else None, precision, _stream()), "vbg_conv2d")
    return y


def stem_pack_weights(w_oihw):
    """[O,3,7,7] -> ([O,7,7,4] for the CUDA-core path, [O,256] zero-padded K operand of the tensor-core stem)."""
    O = w_oihw.shape[0]
    assert tuple(w_oihw.shape[1:]) == (3, 7, 7)
    w774 = torch.empty((O, 7, 7, 4), dtype=torch.float32, device=w_oihw.device)
    w256 = torch.empty((O, 256), dtype=torch.float32, device=w_oihw.device)
    L.check(L.load().vbg_stem_pack_weights(_f32(w_oihw.detach().contiguous()), O, _f32(w774), _f32(w256), _stream()),
            "vbg_stem_pack_weights")
    return w774, w256


def stem_conv(x4, w774, *, ep: Optional[Epilogue] = None, precision=PREC_FP32, W_split=None):
    """7x7/2 pad-3 stem over the padded NHWC4 batch [B, H+6, W+6, 4] -> [B, H/2, W/2, Cout]."""
    B, Hp, Wp, c4 = x4.shape
    H, W = Hp - 6, Wp - 6
    Cout = w774.shape[0]
    y = torch.empty((B, (H - 1) // 2 + 1, (W - 1) // 2 + 1, Cout), dtype=torch.float32, device=x4.device)
    sp, plane = _split_args(W_split, 0)
    L.check(L.load().vbg_stem_conv(_f32(x4, "x4"), B, H, W, _f32(w774, "w"), sp, plane, Cout, _f32(y),
                                   C.byref(ep) if ep is not None else None, precision, _stream()), "vbg_stem_conv")
    return y


def maxpool3x3s2(x, split_out=False):
    B, H, W, Cc = x.shape
    y = _new_act((B, (H - 1) // 2 + 1, (W - 1) // 2 + 1, Cc), x.device, split_out)
    (xp, xpl), (yp, ypl) = _act(x), _act(y)
    L.check(L.load().vbg_maxpool3x3s2_x(xp, xpl, B, H, W, Cc, yp, ypl, _stream()), "vbg_maxpool3x3s2")
    return y


def avgpool2x2(x, split_out=None):
    B, H, W, Cc = x.shape
    y = _new_act((B, H // 2, W // 2, Cc), x.device, isinstance(x, Split) if split_out is None else split_out)
    (xp, xpl), (yp, ypl) = _act(x), _act(y)
    L.check(L.load().vbg_avgpool2x2_x(xp, xpl, B, H, W, Cc, yp, ypl, _stream()), "vbg_avgpool2x2")
    return y


def bn_fold(bn: torch.nn.modules.batchnorm._BatchNorm):
    Cc = bn.num_features
    out = torch.empty((2, Cc), dtype=torch.float32, device=bn.weight.device)
    L.check(L.load().vbg_bn_fold(_f32(bn.weight.detach()), _f32(bn.bias.detach()), _f32(bn.running_mean),
                                 _f32(bn.running_var), bn.eps, Cc, _f32(out[0]), _f32(out[1]), _stream()), "vbg_bn_fold")
    return out[0], out[1]


def repack_oihw_to_ohwi(w):
    O, I, H, W = w.shape
    out = torch.empty((O, H, W, I), dtype=torch.float32, device=w.device)
    L.check(L.load().vbg_repack_oihw_to_ohwi(_f32(w.detach().contiguous()), O, I, H, W, _f32(out), _stream()),
            "vbg_repack_oihw_to_ohwi")
    return out


# ------------------------------------------------------------------ a7
ROI_AUTO, ROI_STREAM, ROI_ROW, ROI_DIRECT = 0, 1, 2, 3
ROI_KERNEL_NAMES = {ROI_STREAM: "roi_align_stream_kernel", ROI_ROW: "roi_align_row_kernel<128>", ROI_DIRECT: "roi_align_kernel"}


def roi_variant(P, C, variant=ROI_AUTO):
    """The kernel ``vbg_roi_align_sel`` launches for this shape (mirrors its AUTO rule; bench.py keys ncu traffic by it)."""
    if variant != ROI_AUTO:
        return variant
    return ROI_STREAM if (P == 7 and C in (128, 256)) else (ROI_ROW if (P == 7 and C % 128 == 0) else ROI_DIRECT)


def roi_align(feat, boxes, seg_off, spatial_scale, P, want_grid=False, split_out=False, variant=ROI_AUTO, out=None):
    """``feat`` fp32 NHWC tensor or Split; ``split_out`` writes the [K,P,P,C] result as a Split.  ``variant`` selects the
    kernel explicitly (tests / scripts); the default picks by shape.  ``out``: a caller-owned result buffer to reuse."""
    B, Hf, Wf, Cc = feat.shape
    K = boxes.shape[0]
    if out is None:
        out = _new_act((K, P, P, Cc), feat.device, split_out)
    sg = torch.empty((K, 2), dtype=torch.int32, device=feat.device) if want_grid else None
    (fp, fpl), (op, opl) = _act(feat), _act(out)
    L.check(L.load().vbg_roi_align_sel(fp, fpl, B, Hf, Wf, Cc, _i32(boxes), _i32(seg_off), K, spatial_scale, P,
                                       op, opl, _i32(sg), int(variant), _stream()), "vbg_roi_align_fwd")
    return (out, sg) if want_grid else out


# ------------------------------------------------------------------ heads / outputs
def softmax_rows(x):
    y = torch.empty_like(x)
    L.check(L.load().vbg_softmax_rows(_f32(x), x.shape[0], x.shape[1], _f32(y), _stream()), "vbg_softmax_rows")
    return y


def full_head_scores(pos_neg, cls):
    R, Cm1 = cls.shape
    out = torch.empty((R, Cm1 + 1), dtype=torch.float32, device=cls.device)
    L.check(L.load().vbg_full_head_scores(_f32(pos_neg), _f32(cls), R, Cm1 + 1, _f32(out), _stream()), "vbg_full_head_scores")
    return out


def upsample_split_nchw(x, up, c_split):
    B, h, w, Ct = x.shape
    o1 = torch.empty((B, c_split, h * up, w * up), dtype=torch.float32, device=x.device)
    o2 = torch.empty((B, Ct - c_split, h * up, w * up), dtype=torch.float32, device=x.device)
    L.check(L.load().vbg_upsample_split_nchw(_f32(x), B, h, w, Ct, up, c_split, _f32(o1), _f32(o2), _stream()),
            "vbg_upsample_split_nchw")
    return o1, o2


def nhwc_to_nchw(x):
    B, H, W, Cc = x.shape
    y = torch.empty((B, Cc, H, W), dtype=torch.float32, device=x.device)
    L.check(L.load().vbg_nhwc_to_nchw(_f32(x), B, H, W, Cc, _f32(y), _stream()), "vbg_nhwc_to_nchw")
    return y


def crf_viterbi(feats, trans, seg_off, B):
    K, T = feats.shape
    tags = torch.empty(K, dtype=torch.float32, device=feats.device)
    scores = torch.empty(B, dtype=torch.float32, device=feats.device)
    ws = torch.empty(max(K * T, 1), dtype=torch.uint8, device=feats.device)
    L.check(L.load().vbg_crf_viterbi(_f32(feats), _f32(trans), _i32(seg_off), B, K, T, _f32(tags), _f32(scores),
                                     _p(ws), ws.numel(), _stream()), "vbg_crf_viterbi")
    return tags, scores


def crf_nll_fwd(feats, trans, tags, seg_off, B):
    """Per-sample CRF negative log-likelihood (model/crf.py:148-152) -> (nll [B], alpha [K,T], logz [B]); tags int32 [K]."""
    K, T = feats.shape
    dev = feats.device
    alpha = torch.empty((K, T), dtype=torch.float32, device=dev)
    logz = torch.empty(B, dtype=torch.float32, device=dev)
    nll = torch.empty(B, dtype=torch.float32, device=dev)
    L.check(L.load().vbg_crf_nll_fwd(_f32(feats), _f32(trans), _i32(tags), _i32(seg_off), B, K, T, _f32(alpha), _f32(logz),
                                     _f32(nll), _stream()), "vbg_crf_nll_fwd")
    return nll, alpha, logz


def crf_nll_bwd(feats, trans, tags, seg_off, B, alpha, dnll):
    """Gradient of sum_b dnll[b] * nll[b] -> (dfeats [K,T], dtrans [T,T])."""
    K, T = feats.shape
    dev = feats.device
    dfeats = torch.empty((K, T), dtype=torch.float32, device=dev)
    part = torch.empty((B, T, T), dtype=torch.float32, device=dev)
    L.check(L.load().vbg_crf_nll_bwd(_f32(feats), _f32(trans), _i32(tags), _i32(seg_off), B, K, T, _f32(alpha), _f32(dnll),
                                     _f32(dfeats), _f32(part), _stream()), "vbg_crf_nll_bwd")
    return dfeats, part.sum(0)


# ------------------------------------------------------------------ backward building blocks
def transpose_split(x, ld_out=None):
    """[rows, cols] (fp32 tensor or Split) -> Split [cols, ld_out] holding x^T, columns rows..ld_out-1 zero."""
    rows, cols = x.shape
    ld = rows if ld_out is None else ld_out
    out = Split.empty((cols, ld), x.device)
    xp, xpl = _act(x)
    L.check(L.load().vbg_transpose_split(xp, xpl, rows, cols, _p(out.t), out.plane, ld, _stream()), "vbg_transpose_split")
    return out


def colsum(x):
    rows, cols = x.shape
    out = torch.empty(cols, dtype=torch.float32, device=x.device)
    xp, xpl = _act(x)
    nb = L.load().vbg_colsum_workspace(rows, cols)
    ws = torch.empty(nb // 4, dtype=torch.float32, device=out.device) if nb else None
    L.check(L.load().vbg_colsum(xp, xpl, rows, cols, _f32(out), None if ws is None else _f32(ws), nb, _stream()), "vbg_colsum")
    return out


def conv_dgrad_weight(w_ohwi):
    """[Cout,kh,kw,Cin] fp32 -> bf16 planes [2, Cin,kh,kw,Cout] of the flipped / transposed weight of the data-gradient conv."""
    Cout, kh, kw, Cin = w_ohwi.shape
    out = torch.empty((2, Cin, kh, kw, Cout), dtype=torch.bfloat16, device=w_ohwi.device)
    L.check(L.load().vbg_conv_dgrad_weight(_f32(w_ohwi.contiguous()), Cout, kh, kw, Cin, _p(out), out[0].numel(), _stream()),
            "vbg_conv_dgrad_weight")
    return out


def layernorm_bwd(x, dy, gamma, eps, want_params=True):
    R, hidden = x.shape
    dx = torch.empty_like(x)
    if not want_params:
        L.check(L.load().vbg_layernorm_bwd(_f32(x), _f32(dy), _f32(gamma), eps, R, hidden, _f32(dx), None, None, None, 0, _stream()),
                "vbg_layernorm_bwd")
        return dx, None, None
    dg = torch.empty(hidden, dtype=torch.float32, device=x.device)
    db = torch.empty(hidden, dtype=torch.float32, device=x.device)
    ws = torch.empty(((R + 63) // 64) * 2 * hidden + 2 * R, dtype=torch.float32, device=x.device)
    L.check(L.load().vbg_layernorm_bwd(_f32(x), _f32(dy), _f32(gamma), eps, R, hidden, _f32(dx), _f32(dg), _f32(db), _f32(ws),
                                       ws.numel() * 4, _stream()), "vbg_layernorm_bwd")
    return dx, dg, db


def linear_wgrad(dy_split, x_split):
    """dW [N, K] = dY^T X over Split operands [M, N] and [M, K] (MN-major tcgen05 operands, no transposes)."""
    M, N = dy_split.shape
    M2, K = x_split.shape
    assert M == M2
    dw = torch.empty((N, K), dtype=torch.float32, device=dy_split.device)
    ws = _workspace(L.load().vbg_linear_wgrad_workspace(M, N, K), dy_split.device)
    L.check(L.load().vbg_linear_wgrad(_p(dy_split.t), dy_split.plane, _p(x_split.t), x_split.plane, M, N, K, _f32(dw), _p(ws),
                                      0 if ws is None else ws.numel(), _stream()), "vbg_linear_wgrad")
    return dw


def conv2d_wgrad(dy_split, x_split, kh, kw, stride, pad):
    """dW [Cout, kh, kw, Cin] from Split dY [B,Ho,Wo,Cout] and Split X [B,H,W,Cin]."""
    B, H, W, Cin = x_split.shape
    Cout = dy_split.shape[-1]
    dw = torch.empty((Cout, kh, kw, Cin), dtype=torch.float32, device=x_split.device)
    ws = _workspace(L.load().vbg_conv2d_wgrad_workspace(B, H, W, Cin, Cout, kh, kw, stride, pad), x_split.device)
    L.check(L.load().vbg_conv2d_wgrad(_p(dy_split.t), dy_split.plane, _p(x_split.t), x_split.plane, B, H, W, Cin, Cout, kh, kw, stride,
                                      pad, _f32(dw), _p(ws), 0 if ws is None else ws.numel(), _stream()), "vbg_conv2d_wgrad")
    return dw


# ------------------------------------------------------------------ training-step kernels (csrc/vbg_train.cu, vbg_attn_bwd.cu)
def _raw(t, name="tensor"):
    """fp32 CUDA tensor whose last dimension is dense (row-strided views allowed)."""
    if not t.is_cuda or t.dtype != torch.float32 or t.stride(-1) != 1:
        raise TypeError(f"{name} must be an fp32 CUDA tensor with a dense last dimension")
    return t.data_ptr()


def _ws_f32(nbytes, device):
    return torch.empty(max(int(nbytes) // 4, 1), dtype=torch.float32, device=device)


def bn_stats(x2d, eps):
    """Batch statistics of [rows, C]: (mean, biased var, rstd)."""
    rows, Cc = x2d.shape
    mean, var, rstd = (torch.empty(Cc, dtype=torch.float32, device=x2d.device) for _ in range(3))
    ws = _ws_f32(L.load().vbg_bn_workspace(rows, Cc), x2d.device)
    L.check(L.load().vbg_bn_stats(_f32(x2d), rows, Cc, eps, _f32(mean), _f32(var), _f32(rstd), _f32(ws), ws.numel() * 4, _stream()),
            "vbg_bn_stats")
    return mean, var, rstd


def bn_apply(x2d, mean, rstd, gamma, beta, residual=None, relu=False):
    rows, Cc = x2d.shape
    y = torch.empty_like(x2d)
    L.check(L.load().vbg_bn_apply(_f32(x2d), rows, Cc, _f32(mean), _f32(rstd), _f32(gamma), _f32(beta),
                                  None if residual is None else _f32(residual), int(relu), _f32(y), _stream()), "vbg_bn_apply")
    return y


def bn_bwd(x2d, dy2d, y_relu, mean, rstd, gamma, want_dres=False):
    """-> (dx, dres or None, dgamma, dbeta)."""
    rows, Cc = x2d.shape
    dx = torch.empty_like(x2d)
    dres = torch.empty_like(x2d) if want_dres else None
    dg, db = (torch.empty(Cc, dtype=torch.float32, device=x2d.device) for _ in range(2))
    ws = _ws_f32(L.load().vbg_bn_workspace(rows, Cc), x2d.device)
    L.check(L.load().vbg_bn_bwd(_f32(x2d), _f32(dy2d), None if y_relu is None else _f32(y_relu), rows, Cc, _f32(mean), _f32(rstd),
                                _f32(gamma), _f32(dx), None if dres is None else _f32(dres), _f32(dg), _f32(db), _f32(ws),
                                ws.numel() * 4, _stream()), "vbg_bn_bwd")
    return dx, dres, dg, db


def bn_bwd_reduce(x2d, dy2d, y_relu, mean, rstd):
    """This rank's (sum dy' * xhat, sum dy') of vbg_bn_bwd, before a SyncBatchNorm all-reduce."""
    rows, Cc = x2d.shape
    s_xhat, s_dy = (torch.empty(Cc, dtype=torch.float32, device=x2d.device) for _ in range(2))
    ws = _ws_f32(L.load().vbg_bn_workspace(rows, Cc), x2d.device)
    L.check(L.load().vbg_bn_bwd_reduce(_f32(x2d), _f32(dy2d), None if y_relu is None else _f32(y_relu), rows, Cc, _f32(mean),
                                       _f32(rstd), _f32(s_xhat), _f32(s_dy), _f32(ws), ws.numel() * 4, _stream()),
            "vbg_bn_bwd_reduce")
    return s_xhat, s_dy


def bn_bwd_dx(x2d, dy2d, y_relu, mean, rstd, gamma, s_xhat, s_dy, count, want_dres=False):
    """dx (and the residual's gradient) from per-channel sums taken over ``count`` rows (all ranks)."""
    rows, Cc = x2d.shape
    dx = torch.empty_like(x2d)
    dres = torch.empty_like(x2d) if want_dres else None
    L.check(L.load().vbg_bn_bwd_dx(_f32(x2d), _f32(dy2d), None if y_relu is None else _f32(y_relu), rows, Cc, 1.0 / float(count),
                                   _f32(mean), _f32(rstd), _f32(gamma), _f32(s_xhat), _f32(s_dy), _f32(dx),
                                   None if dres is None else _f32(dres), _stream()), "vbg_bn_bwd_dx")
    return dx, dres


def maxpool3x3s2_bwd(x, dy, y=None):
    """``y``: the forward's output, when the caller kept it (the gather then scans a window only where x equals its maximum)."""
    B, H, W, Cc = x.shape
    dx = torch.empty_like(x)
    if y is not None:
        L.check(L.load().vbg_maxpool3x3s2_bwd_y(_f32(x), _f32(y), _f32(dy), B, H, W, Cc, _f32(dx), _stream()), "vbg_maxpool3x3s2_bwd_y")
    else:
        L.check(L.load().vbg_maxpool3x3s2_bwd(_f32(x), _f32(dy), B, H, W, Cc, _f32(dx), _stream()), "vbg_maxpool3x3s2_bwd")
    return dx


def sumpool2x2(x, scale=1.0):
    B, H, W, Cc = x.shape
    y = torch.empty((B, H // 2, W // 2, Cc), dtype=torch.float32, device=x.device)
    L.check(L.load().vbg_sumpool2x2(_f32(x), B, H, W, Cc, scale, _f32(y), _stream()), "vbg_sumpool2x2")
    return y


def expand2x(x, H, W, scale=1.0, zero_insert=False):
    B, Hi, Wi, Cc = x.shape
    y = torch.empty((B, H, W, Cc), dtype=torch.float32, device=x.device)
    L.check(L.load().vbg_expand2x(_f32(x), B, Hi, Wi, Cc, H, W, scale, int(zero_insert), _f32(y), _stream()), "vbg_expand2x")
    return y


def gelu(x, dy=None, split_out=False):
    """erf-GELU (``dy`` None) or its backward ``dy * gelu'(x)``; ``x`` / ``dy`` fp32 tensors or Splits, the result a Split when
    ``split_out``."""
    if not split_out and not isinstance(x, Split) and not isinstance(dy, Split):
        out = torch.empty_like(x)
        L.check(L.load().vbg_gelu(_f32(x), None if dy is None else _f32(dy), x.numel(), _f32(out), _stream()), "vbg_gelu")
        return out
    shape = tuple(x.shape)
    n = 1
    for d in shape:
        n *= int(d)
    out = _new_act(shape, (x.t if isinstance(x, Split) else x).device, split_out)
    xp, xpl = _act(x)
    dp, dpl = _act(dy) if dy is not None else (None, 0)
    op, opl = _act(out)
    L.check(L.load().vbg_gelu_x(xp, xpl, dp, dpl, n, op, opl, _stream()), "vbg_gelu_x")
    return out


def uniform_keys(n, seed, step_seed, device):
    """int32 [n] of 31-bit hash keys (sampling keys of losses_device.py)."""
    out = torch.empty(int(n), dtype=torch.int32, device=device)
    L.check(L.load().vbg_uniform_keys(int(n), int(seed), _u64ptr(step_seed), _i32(out), _stream()), "vbg_uniform_keys")
    return out


def dropout(x, p, seed, step_seed=None):
    y = torch.empty_like(x)
    L.check(L.load().vbg_dropout_ds(_f32(x), x.numel(), p, seed, _u64ptr(step_seed), _f32(y), _stream()), "vbg_dropout")
    return y


def grid_scatter_bwd(dgrid_cells, ld, idx, boxes, seg_off, B, K, stride, Cc):
    """``dgrid_cells``: fp32 view whose cell rows are ``ld`` floats apart (a column slice of a wider gradient is fine)."""
    Hg, Wg = idx.shape[1], idx.shape[2]
    demb = torch.empty((K, Cc), dtype=torch.float32, device=idx.device)
    L.check(L.load().vbg_grid_scatter_bwd(_raw(dgrid_cells), ld, _i32(idx), _i32(boxes), _i32(seg_off), B, K, stride, Hg, Wg, Cc,
                                          _f32(demb), _stream()), "vbg_grid_scatter_bwd")
    return demb


def segment_reduce_bwd(dseg, tok_row, seg_start, R, mode=AGG_MEAN):
    K, Cc = dseg.shape
    dh = torch.zeros((R, Cc), dtype=torch.float32, device=dseg.device)
    L.check(L.load().vbg_segment_reduce_bwd(_f32(dseg), _i32(tok_row), _i32(seg_start), K, Cc, mode, _f32(dh), _stream()),
            "vbg_segment_reduce_bwd")
    return dh


def embed_bwd(dx, ids, pos, vocab, max_pos):
    R, Hd = dx.shape
    dword = torch.zeros((vocab, Hd), dtype=torch.float32, device=dx.device)
    dpos = torch.zeros((max_pos, Hd), dtype=torch.float32, device=dx.device)
    L.check(L.load().vbg_embed_bwd(_f32(dx), _i32(ids), _i32(pos), R, Hd, _f32(dword), _f32(dpos), _stream()), "vbg_embed_bwd")
    return dword, dpos


def roi_align_bwd(dout, boxes, seg_off, B, Hf, Wf, spatial_scale):
    K, Pp, _, Cc = dout.shape
    dfeat = torch.empty((B, Hf, Wf, Cc), dtype=torch.float32, device=dout.device)        # every pixel is written by the kernel
    nb = int(L.load().vbg_roi_align_bwd_workspace(K))
    ws = torch.empty(nb, dtype=torch.uint8, device=dout.device)
    L.check(L.load().vbg_roi_align_bwd(_f32(dout), B, Hf, Wf, Cc, _i32(boxes), _i32(seg_off), K, spatial_scale, Pp, _f32(dfeat),
                                       _p(ws), nb, _stream()), "vbg_roi_align_bwd")
    return dfeat


def seg_ce_bwd(logits_lowres, pos_neg, cls, H, W, up, c_split, gscale):
    B, h, w, Ct = logits_lowres.shape
    dl = torch.empty_like(logits_lowres)
    L.check(L.load().vbg_seg_ce_bwd(_f32(logits_lowres), _p(pos_neg, torch.int64), _p(cls, torch.int64), B, H, W, up, Ct, c_split,
                                    _f32(gscale), _f32(dl), _stream()), "vbg_seg_ce_bwd")
    return dl


def upsample_split_bwd(d1, d2, up):
    B, c_split, H, W = d1.shape
    Ct = c_split + d2.shape[1]
    dl = torch.empty((B, H // up, W // up, Ct), dtype=torch.float32, device=d1.device)
    L.check(L.load().vbg_upsample_split_bwd(_f32(d1.contiguous()), _f32(d2.contiguous()), B, H // up, W // up, Ct, up, c_split,
                                            _f32(dl), _stream()), "vbg_upsample_split_bwd")
    return dl


def small_wgrad(dy2d, x2d):
    """dW [N <= 16, K] = dY^T X (row strides taken from the tensors: column slices are fine)."""
    M, N = dy2d.shape
    K = x2d.shape[1]
    dw = torch.empty((N, K), dtype=torch.float32, device=x2d.device)
    ws = _ws_f32(L.load().vbg_small_wgrad_workspace(M, N, K), x2d.device)
    L.check(L.load().vbg_small_wgrad(_raw(dy2d), dy2d.stride(0), _raw(x2d), x2d.stride(0), M, N, K, _f32(dw), _f32(ws),
                                     ws.numel() * 4, _stream()), "vbg_small_wgrad")
    return dw


def stem_wgrad(x4, dy):
    """dW [64,7,7,4] of the 7x7/2 stem from the zero-bordered NHWC4 batch and dY [B,Ho,Wo,64]."""
    B, Hp, Wp, _ = x4.shape
    _, Ho, Wo, Cout = dy.shape
    assert Cout == 64
    dw = torch.empty((64, 7, 7, 4), dtype=torch.float32, device=dy.device)
    ws = _ws_f32(L.load().vbg_stem_wgrad_workspace(), dy.device)
    L.check(L.load().vbg_stem_wgrad(_f32(x4), _f32(dy), B, Hp, Wp, Ho, Wo, _f32(dw), _f32(ws), ws.numel() * 4, _stream()),
            "vbg_stem_wgrad")
    return dw


def attention_bwd(qkv, out, d_out, cu, nseq, max_len, heads):
    R, three_hid = qkv.shape
    hid = three_hid // 3
    dqkv = torch.empty_like(qkv)
    ws = torch.empty(R * heads * 2, dtype=torch.float32, device=qkv.device)
    L.check(L.load().vbg_attention_bwd(_f32(qkv), _f32(out), _f32(d_out), _i32(cu), nseq, max_len, heads, hid // heads, R, _f32(dqkv),
                                       _f32(ws), ws.numel() * 4, _stream()), "vbg_attention_bwd")
    return dqkv
