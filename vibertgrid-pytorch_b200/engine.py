"""Forward engine: sequences the sm_100a kernels of the joint forward
(reference model/ViBERTgrid_net.py:512-544) over one batch, with zero
device->host synchronisations.

Data layout in HBM: fp32, channels-last.  One-time
parameter preparation (cached until a weight changes): conv weights repacked
OIHW->OHWI, eval BatchNorm folded to per-channel (scale, shift), BERT Q/K/V
weights packed to one [2304,768] operand, the two seg-head 1x1 convs packed,
the ROI FC weight permuted to the NHWC flatten order.

Algebraic restructurings (results equal up to fp32 re-association):
  * BERT windows -> one packed varlen batch of real rows only (plan.py)
  * cat(x, grid) -> 1x1 conv     ==  two-source K-split GEMM (no concat buffer)
  * up(a) + lateral(b) -> conv   ==  lateral GEMM with nearest-x2 residual epilogue
  * fuse(cat[up8,up4,up2,id])    ==  4 chained 1x1 GEMMs at native resolution (3x fewer FLOPs)
  * 1x1(up4(x))                  ==  up4(1x1(x))  (seg head: 16x fewer FLOPs, no 268 MB/img tensor)
"""
from __future__ import annotations

import os
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn as nn

from . import ops
from . import params as P
from .ops import (ACT_GELU, ACT_NONE, ACT_RELU, PREC_BF16X3, PREC_FP32, PREC_TF32, RES_NONE, RES_SAME, RES_UP2,
                  Split, make_epilogue)
from .plan import plan_batch, roberta_position_ids


class _Prepared:
    def __init__(self):
        self.convw: Dict[int, torch.Tensor] = {}     # id(conv) -> fp32 weight, [Cout, kh, kw, Cin]
        self.split: Dict[int, torch.Tensor] = {}     # id(module) or id(fp32 tensor) -> bf16 [2, ...] hi/lo planes
        self.bn: Dict[int, tuple] = {}
        self.bert_layers = []
        self.misc: Dict[str, torch.Tensor] = {}


def _cat_into(dst, parts):
    """torch.cat into a static graph input (dtype converted when the caller's tensors differ from the staged dtype)."""
    if all(p.dtype == dst.dtype for p in parts):
        torch.cat(parts, 0, out=dst)
    else:
        dst.copy_(torch.cat(parts, 0), non_blocking=True)


class Intermediates(dict):
    """Engine outputs keyed by name.  Activations that live as bf16 hi/lo planes between the pre-split tensor-core
    kernels (P_fuse, ROI features, BERTgrid ...) are merged to fp32 on access -- inspection and tests only, never on
    the hot path.  ``raw(key)`` returns the stored object."""

    def __getitem__(self, k):
        v = dict.__getitem__(self, k)
        return v.float() if isinstance(v, Split) else v

    def raw(self, k):
        return dict.__getitem__(self, k)


class ForwardEngine:
    def __init__(self, net):
        self.net = net
        self._prep: Optional[_Prepared] = None
        self._fp = None
        self._fp_tensors = None
        self._prep_gen = 0
        self.kernel_launches = 0          # C-ABI kernel launches executed on the device (graph replays included)
        env = os.environ.get("VBG_PRECISION", "").lower()
        self.precision = {"fp32": PREC_FP32, "tf32": PREC_TF32, "bf16x3": PREC_BF16X3}.get(env)
        self.launches = 0
        self.use_graphs = os.environ.get("VBG_CUDA_GRAPHS", "1") != "0"
        self.fuse_aux_loss = os.environ.get("VBG_FUSED_AUX_LOSS", "1") != "0"
        # bf16x3 only: keep activations as bf16 hi/lo planes between the tensor-core kernels (no in-kernel conversion)
        self.presplit = os.environ.get("VBG_PRESPLIT", "1") != "0"
        # run the two independent branches of the forward (BERT || early backbone, seg head || ROI path) on two streams
        self.multi_stream = os.environ.get("VBG_STREAMS", "2") != "1"
        self._side = {}
        self.max_graphs = 8
        self._graphs: Dict[tuple, dict] = {}
        self.graph_replays = 0

    # ------------------------------------------------------------------ parameter preparation
    def invalidate(self):
        self._prep, self._fp, self._fp_tensors = None, None, None
        self._graphs.clear()

    def _fingerprint(self):
        if self._fp_tensors is None:      # Parameter / buffer objects survive .to() and in-place loads; cache the list
            self._fp_tensors = list(self.net.parameters()) + list(self.net.buffers())
        return tuple((t.data_ptr(), t._version) for t in self._fp_tensors)

    def _prepare(self):
        fp = self._fingerprint()
        if self._prep is not None and fp == self._fp:
            return self._prep
        net, pr = self.net, _Prepared()
        want_split = self._prec() == PREC_BF16X3

        def split_of(key, w):
            if want_split:
                pr.split[key] = ops.split_bf16(w)

        stem = net.backbone.resnet.conv1 if net.backbone.pretrained_layout else net.backbone.conv_1[0]
        for m in net.modules():
            if isinstance(m, nn.Conv2d):
                w = m.weight.detach()
                if m is stem:
                    w774, w256 = ops.stem_pack_weights(w)
                    pr.convw[id(m)] = w774
                    split_of(id(m), w256)
                    continue
                if w.shape[2] == 1 and w.shape[3] == 1:
                    pr.convw[id(m)] = w.reshape(w.shape[0], 1, 1, w.shape[1]).contiguous()  # OIHW == OHWI for 1x1
                else:
                    pr.convw[id(m)] = ops.repack_oihw_to_ohwi(w)
                split_of(id(m), pr.convw[id(m)])
            elif isinstance(m, nn.modules.batchnorm._BatchNorm):
                pr.bn[id(m)] = ops.bn_fold(m)
            elif isinstance(m, nn.Linear) and m.out_features >= 64 and m.in_features % 64 == 0:
                split_of(id(m), m.weight)
        for lyr in net.bert_model.encoder.layer:
            sa = lyr.attention.self
            pk = dict(wqkv=torch.cat([sa.query.weight, sa.key.weight, sa.value.weight], 0).detach().contiguous(),
                      bqkv=torch.cat([sa.query.bias, sa.key.bias, sa.value.bias], 0).detach().contiguous())
            pk["wqkv_split"] = ops.split_bf16(pk["wqkv"]) if want_split else None
            pr.bert_layers.append(pk)
            for m in (sa.query, sa.key, sa.value):
                pr.split.pop(id(m), None)          # only the packed QKV operand is used
        bm = net.bert_model
        if hasattr(bm, "pooler"):
            pr.split.pop(id(bm.pooler.dense), None)  # never on the forward path
        roi = net.late_fusion_net.ROI_embedding_net
        Cc, Pp = net.p_fuse_channel, net.roi_shape
        pr.misc["roi_fc_w"] = ops.repack_oihw_to_ohwi(
            roi.linear.weight.detach().reshape(-1, Cc, Pp, Pp)).reshape(roi.linear.out_features, -1)
        pr.split.pop(id(roi.linear), None)
        split_of("roi_fc_w", pr.misc["roi_fc_w"])
        if net.semantic_segmentation_head is not None:
            enc = net.semantic_segmentation_head.encoder
            pr.misc["seg_w"] = torch.cat([enc.conv_3_1.weight.flatten(1), enc.conv_3_2.weight.flatten(1)], 0).detach().contiguous()
            pr.misc["seg_b"] = torch.cat([enc.conv_3_1.bias, enc.conv_3_2.bias], 0).detach().contiguous()
            split_of("seg_w", pr.misc["seg_w"])
            for m in (enc.conv_3_1, enc.conv_3_2):
                pr.split.pop(id(m), None)          # N = 3 + C < 64: CUDA-core kernel
        head = net.field_type_classification_head
        if isinstance(head, P.FullHeadParams) and net.layer_mode == "single":
            nets = [getattr(head, f"category_classification_net_{i}") for i in range(net.num_tokens - 1)]
            pr.misc["full_w"] = torch.cat([n.layer.linear.weight for n in nets], 0).detach().contiguous()
            pr.misc["full_b"] = torch.cat([n.layer.linear.bias for n in nets], 0).detach().contiguous()
        self._prep, self._fp = pr, fp
        self._prep_gen += 1
        self._graphs.clear()              # captured graphs hold pointers into the previous preparation
        return pr

    def _aux_loss_is_default(self):
        """True when the auxiliary loss is the plain mean cross entropy of the `simp` head (no sampling / OHEM / weights):
        the configuration vbg_seg_ce_loss implements."""
        net = self.net
        cfg = net.loss_cfg
        return (net.classifier_mode == "simp" and cfg["aux_sample_list"] is None and tuple(cfg["aux"]) == (-1, -1)
                and net.loss_weights is None)

    # ------------------------------------------------------------------ building blocks
    def _prec(self):
        if self.precision is None:
            self.precision = PREC_BF16X3 if ops.tc_available() else PREC_FP32
        return self.precision

    def _ps(self):
        return self.presplit and self._prec() == PREC_BF16X3

    def _side_stream(self, dev):
        key = dev.index if dev.index is not None else torch.cuda.current_device()
        if key not in self._side:
            self._side[key] = torch.cuda.Stream(dev)
        return self._side[key]

    @staticmethod
    def _tc_shape(N, K):
        return N >= 64 and K % 64 == 0

    def _conv(self, x, conv: nn.Conv2d, bn=None, act=ACT_NONE, residual=None, res_mode=RES_NONE, out_f32=False):
        """``x`` / ``residual``: fp32 NHWC tensors, or Splits in pre-split mode (then the result is a Split unless out_f32)."""
        pr = self._prep
        w, ws = pr.convw[id(conv)], pr.split.get(id(conv))
        scale, shift = pr.bn[id(bn)] if bn is not None else (None, None if conv.bias is None else conv.bias.detach())
        B, H, W, Cin = x.shape
        k, s, p = conv.kernel_size[0], conv.stride[0], conv.padding[0]
        Cout = conv.out_channels
        if isinstance(x, Split) and not (self._tc_shape(Cout, Cin) and ws is not None):
            x = x.float()                       # shape the tensor-core kernels do not take: CUDA-core path, fp32 in / out
            residual = ops.as_f32(residual)
        so = isinstance(x, Split) and not out_f32
        if k == 1 and s == 1:
            ep = make_epilogue(scale, shift, residual, res_mode, ldr=Cout, out_h=H, out_w=W, act=act)
            y = ops.gemm(x.view(B * H * W, Cin), w.view(Cout, Cin), ep=ep, precision=self._prec(), W_split=ws, split_out=so)
            return y.view(B, H, W, Cout)
        ep = make_epilogue(scale, shift, residual, res_mode, ldr=Cout, act=act)
        return ops.conv2d(x, w, s, p, ep=ep, precision=self._prec(), W_split=ws, split_out=so)

    def _stem(self, x4, conv: nn.Conv2d, bn):
        pr = self._prep
        scale, shift = pr.bn[id(bn)]
        ep = make_epilogue(scale, shift, act=ACT_RELU)
        return ops.stem_conv(x4, pr.convw[id(conv)], ep=ep, precision=self._prec(), W_split=pr.split.get(id(conv)))

    def _lin(self, x, lin: nn.Linear, act=ACT_NONE, residual=None, A2=None, W=None, W_split=None, out_f32=False):
        ep = make_epilogue(None, lin.bias.detach(), residual, RES_SAME if residual is not None else RES_NONE,
                           ldr=lin.out_features, act=act)
        if W is None:
            W, W_split = lin.weight.detach(), self._prep.split.get(id(lin))
        if isinstance(x, Split) and not (self._tc_shape(W.shape[0], W.shape[1]) and W_split is not None):
            x, A2 = x.float(), ops.as_f32(A2)
            ep = make_epilogue(None, lin.bias.detach(), ops.as_f32(residual), RES_SAME if residual is not None else RES_NONE,
                               ldr=lin.out_features, act=act)
        so = isinstance(x, Split) and not out_f32
        return ops.gemm(x, W, A2=A2, ep=ep, precision=self._prec(), W_split=W_split, split_out=so)

    def _pick(self, x, xs, lin: nn.Linear):
        """The storage format a linear's kernel wants: the Split for tensor-core shapes, fp32 for the CUDA-core ones."""
        return xs if (xs is not None and self._tc_shape(lin.out_features, lin.in_features)) else x

    def _mlp_or_lin(self, x, m, xs=None):
        """Head MLPs; ``xs`` is the Split twin of the fp32 ``x`` (pre-split mode).  Results are fp32."""
        if hasattr(m, "linear_1"):
            h = self._lin(self._pick(x, xs, m.linear_1), m.linear_1, ACT_RELU, out_f32=True)
            return self._lin(h, m.linear_2, out_f32=True)
        return self._lin(self._pick(x, xs, m.linear), m.linear, out_f32=True)

    def _block(self, x, conv1, bn1, conv2, bn2, shortcut):
        y = self._conv(x, conv1, bn1, ACT_RELU)
        if shortcut is None:
            sc = x
        else:
            kind, sconv, sbn = shortcut
            sc = self._conv(ops.avgpool2x2(x) if kind == "avg" else x, sconv, sbn)
        return self._conv(y, conv2, bn2, ACT_RELU, residual=sc, res_mode=RES_SAME)

    def _our_block(self, x, blk: P.ResBlockParams):
        sc = None
        if blk.downsample:
            sc = ("avg", blk.conv_shortcut[1], blk.conv_shortcut[2]) if blk.d_variant \
                else ("conv", blk.conv_shortcut[0], blk.conv_shortcut[1])
        return self._block(x, blk.conv_1, blk.bn_1, blk.conv_2, blk.bn_2, sc)

    def _tv_block(self, x, blk):
        sc = ("conv", blk.downsample[0], blk.downsample[1]) if hasattr(blk, "downsample") else None
        return self._block(x, blk.conv1, blk.bn1, blk.conv2, blk.bn2, sc)

    def _early_fusion(self, x2, grid, conv: nn.Conv2d):
        B, H, W, C1 = x2.shape
        ep = make_epilogue(None, None if conv.bias is None else conv.bias.detach())
        y = ops.gemm(x2.view(B * H * W, C1), self._prep.convw[id(conv)].view(conv.out_channels, -1),
                     A2=grid.view(B * H * W, grid.shape[-1]), ep=ep,
                     precision=self._prec(), W_split=self._prep.split.get(id(conv)), split_out=isinstance(x2, Split))
        return y.view(B, H, W, conv.out_channels)

    # ------------------------------------------------------------------ stages
    def _bert(self, plan, dev_tab, corpus):
        net, pr = self.net, self._prep
        bm = net.bert_model
        e = bm.embeddings
        heads = bm.cfg["num_attention_heads"]
        seq_tab, cu = dev_tab["seq_tab"], dev_tab["cu"]
        ids, pos = ops.bert_assemble(corpus, seq_tab, cu, plan.nseq, plan.R)
        if bm.cfg.get("roberta"):
            pos = roberta_position_ids(ids, pos, int(bm.cfg["pad_token_id"]), cu)
        prec = self._prec()
        ps = self._ps()
        x = ops.embed_ln(ids, pos, e.word_embeddings.weight.detach(), e.position_embeddings.weight.detach(),
                         e.token_type_embeddings.weight.detach()[0], e.LayerNorm.weight.detach(),
                         e.LayerNorm.bias.detach(), e.LayerNorm.eps, split=ps)
        hid = bm.cfg["hidden_size"]
        split_attn = (prec == PREC_BF16X3 and hid // heads == 64 and hid % 64 == 0 and plan.max_len <= 512
                      and os.environ.get("VBG_ATTN_SPLIT", "1") != "0")
        n_layers = len(bm.encoder.layer)
        for li, (lyr, pk) in enumerate(zip(bm.encoder.layer, pr.bert_layers)):
            # pre-split mode: x, ctx and the FFN intermediate are bf16 hi/lo planes (GEMM A operands and residuals);
            # only the two LayerNorm inputs per layer and the final hidden state are fp32.
            if split_attn:      # QKV projection writes bf16 hi/lo planes; TMA-fed tcgen05 attention consumes them directly
                qkv = ops.gemm(x, pk["wqkv"], ep=make_epilogue(None, pk["bqkv"]), precision=prec, W_split=pk["wqkv_split"],
                               split_out=True)
                ctx = ops.attention_split(qkv, cu, plan.nseq, plan.max_len, heads, split_out=ps)
            else:
                qkv = ops.gemm(x, pk["wqkv"], ep=make_epilogue(None, pk["bqkv"]), precision=prec, W_split=pk["wqkv_split"])
                ctx = ops.attention(qkv, cu, plan.nseq, plan.max_len, heads, prec)
                if ps:
                    ctx = ops.to_split(ctx)
            ao = lyr.attention.output
            a = self._lin(ctx, ao.dense, residual=x, out_f32=True)
            x = ops.layernorm(a, ao.LayerNorm.weight.detach(), ao.LayerNorm.bias.detach(), ao.LayerNorm.eps,
                              out=None if ps else a, split=ps)
            h = self._lin(x, lyr.intermediate.dense, ACT_GELU)
            o = self._lin(h, lyr.output.dense, residual=x, out_f32=True)
            last = li == n_layers - 1
            x = ops.layernorm(o, lyr.output.LayerNorm.weight.detach(), lyr.output.LayerNorm.bias.detach(),
                              lyr.output.LayerNorm.eps, out=o if (last or not ps) else None, split=ps and not last)
        return x

    def _backbone_pre(self, img):
        """Stem .. first block of stage 3 (conv_3_x.block_1 / layer2[0]): everything before the early fusion, i.e. the part
        of the backbone that does not depend on the BERTgrid and can run beside the BERT encoder."""
        bb = self.net.backbone
        if bb.pretrained_layout:
            r = bb.resnet
            x1 = ops.maxpool3x3s2(self._stem(img, r.conv1, r.bn1), split_out=self._ps())
            for blk in r.layer1:
                x1 = self._tv_block(x1, blk)
            x2 = self._tv_block(x1, r.layer2[0])
        else:
            x1 = ops.maxpool3x3s2(self._stem(img, bb.conv_1[0], bb.conv_1[1]), split_out=self._ps())
            for blk in bb.conv_2_x:
                x1 = self._our_block(x1, blk)
            x2 = self._our_block(x1, bb.conv_3_x.block_1)
        return x1, x2

    def _backbone_post(self, x1, x2, grid):
        bb = self.net.backbone
        if bb.pretrained_layout:
            r = bb.resnet
            x2 = self._early_fusion(x2, grid, bb.early_fusion)
            for blk in list(r.layer2)[1:]:
                x2 = self._tv_block(x2, blk)
            x3 = x2
            for blk in r.layer3:
                x3 = self._tv_block(x3, blk)
            x4 = x3
            for blk in r.layer4:
                x4 = self._tv_block(x4, blk)
        else:
            x2 = self._early_fusion(x2, grid, bb.conv_3_x.early_fusion)
            for blk in bb.conv_3_x.layers:
                x2 = self._our_block(x2, blk)
            x3 = x2
            for blk in bb.conv_4_x:
                x3 = self._our_block(x3, blk)
            x4 = x3
            for blk in bb.conv_5_x:
                x4 = self._our_block(x4, blk)
        # FPN top-down: lateral 1x1 GEMM with the nearest-x2 upsample-add fused in its epilogue
        x4 = self._conv(x4, bb.conv_6_x)
        x5 = self._conv(self._conv(x3, bb.skip_1, residual=x4, res_mode=RES_UP2), bb.merge_1)
        x6 = self._conv(self._conv(x2, bb.skip_2, residual=x5, res_mode=RES_UP2), bb.merge_2)
        x7 = self._conv(self._conv(x1, bb.skip_3, residual=x6, res_mode=RES_UP2), bb.merge_3)
        # fuse(cat[up8 x4, up4 x5, up2 x6, x7]) as four chained K-slices of fuse.weight at native resolution
        wf = self._prep.convw[id(bb.fuse)].view(bb.fuse.out_channels, -1)   # [256, 1024]
        wfs = self._prep.split.get(id(bb.fuse))
        Pc = wf.shape[1] // 4
        prec = self._prec()
        t = None
        for i, lvl in enumerate((x4, x5, x6, x7)):
            B, H, W, Cc = lvl.shape
            ep = make_epilogue(residual=t, res_mode=RES_UP2 if t is not None else RES_NONE, out_h=H, out_w=W)
            t = ops.gemm(lvl.view(B * H * W, Cc), wf, ep=ep, precision=prec, N=wf.shape[0], K=Pc, ldw=wf.shape[1],
                         w_offset=i * Pc, W_split=wfs, split_out=isinstance(lvl, Split)).view(B, H, W, wf.shape[0])
        return t

    def _seg_head(self, p_fuse):
        enc = self.net.semantic_segmentation_head.encoder
        x = self._conv(p_fuse, enc.conv_1, enc.bn_1, ACT_RELU)
        pr = self._prep
        # the packed 1x1 heads have N = 3 + C outputs: on the tensor cores (a 64-wide tile with 3 + C live columns) when the
        # pre-split kernel takes the shape (N % 4 == 0), else the CUDA-core GEMM over an fp32 copy
        tc_heads = (pr.split.get("seg_w") is not None and pr.misc["seg_w"].shape[0] % 4 == 0 and self._ps()
                    and x.shape[0] * x.shape[1] * x.shape[2] >= 8192)      # vbg_gemm_ps takes N < 64 only for tall problems
        x = self._conv(x, enc.conv_2, enc.bn_2, ACT_RELU, out_f32=not tc_heads)
        B, H, W, Cc = x.shape
        lg = ops.gemm(x.view(B * H * W, Cc), pr.misc["seg_w"], ep=make_epilogue(None, pr.misc["seg_b"]), precision=self._prec(),
                      W_split=pr.split.get("seg_w") if tc_heads else None)
        lg = lg.view(B, H, W, -1)
        return ops.upsample_split_nchw(lg, self.net.p_fuse_downsampling_ratio, 3) + (lg,)

    # ------------------------------------------------------------------ whole forward
    @torch.no_grad()
    def run(self, image, seg_indices, seg_classes, coors, corpus, mask, want_seg=True, crf_one_sequence=False):
        """One joint forward.  A batch signature (all tensor shapes) seen for the second time is captured into a CUDA
        graph; from then on the step is: copy the inputs into the graph's static buffers, one graph launch.
        Tensors in the returned dict are owned by the engine and are overwritten by the next call with the same
        signature (``ViBERTgridNet.forward`` clones what it returns)."""
        net = self.net
        dev = corpus.device
        standins = getattr(self, "_test_standins", False)
        if dev.type != "cuda" and not standins:
            raise RuntimeError("ViBERTgridNet (B200) runs on CUDA tensors only; there is no CPU fallback")
        self._prepare()
        min_size = float(net.test_image_min_size)            # eval/inference branch of transform.py:192-196
        shapes = (tuple(ops.image_hw(im) if not standins else tuple(im.shape[-2:]) for im in image), tuple(int(s.shape[0]) for s in seg_indices),
                  tuple(int(c.shape[0]) for c in coors), int(corpus.shape[1]))
        want_seg = bool(want_seg and net.semantic_segmentation_head is not None)
        u8 = image[0].dtype == torch.uint8                    # decoded pixels [h, w, 3] (shards.py) instead of ToTensor's fp32 planes
        if any((im.dtype == torch.uint8) != u8 for im in image):
            raise TypeError("all images of a batch must share one format: float32 [3, h, w] or uint8 [h, w, 3]")
        key = (shapes, want_seg, self._prec(), self._prep_gen, dev.index, self.fuse_aux_loss, bool(crf_one_sequence), u8)
        ent = self._graphs.get(key) if self.use_graphs and not standins else None
        if ent is not None and ent.get("graph") is not None:
            st = ent["static"]
            for dst, src in zip(st["image"], image):
                dst.copy_(src, non_blocking=True)
            _cat_into(st["coors"], [c.reshape(-1, 4) for c in coors])
            _cat_into(st["seg_ids"], [s.reshape(-1) for s in seg_indices])
            if want_seg:
                _cat_into(st["cls"], [c.reshape(-1) for c in seg_classes])
            st["corpus"].copy_(corpus, non_blocking=True)
            if st["mask"] is not None:
                st["mask"].copy_(mask, non_blocking=True)
            ent["graph"].replay()
            self.graph_replays += 1
            self.kernel_launches += ent["launches"]
            return ent["out"]

        plan = plan_batch(list(shapes[0]), list(shapes[1]), list(shapes[2]), shapes[3], min_size, float(net.image_max_size))
        tab = torch.from_numpy(plan.table).to(dev)          # the step's only H2D besides the inputs
        uniform = len({tuple(im.shape) for im in image}) == 1 and len(set(plan.sizes)) == 1 and not standins
        stack = torch.stack([im for im in image], 0).contiguous() if uniform else None     # one normalise launch for the batch
        static = dict(
            image=[stack[b] for b in range(len(image))] if uniform else [im.contiguous() for im in image], image_stack=stack,
            coors=torch.cat([c.reshape(-1, 4) for c in coors], 0).to(torch.int64).contiguous(),
            seg_ids=torch.cat([s.reshape(-1) for s in seg_indices], 0).to(torch.int32).contiguous(),
            cls=torch.cat([c.reshape(-1) for c in seg_classes], 0).to(torch.int32).contiguous() if want_seg else None,
            corpus=corpus.contiguous(), tab=tab,
            mask=None if mask is None else mask.to(torch.int32).contiguous())
        ragged_u8 = u8 and not uniform
        static["img_tab"] = ops.image_table(static["image"], plan.sizes) if ragged_u8 else None
        if self.use_graphs and not standins:
            if ent is None:                                   # first sighting: run eagerly (also warms up lazy kernel attributes)
                self._graphs[key] = {"graph": None}
                if len(self._graphs) > self.max_graphs:
                    self._graphs.pop(next(iter(self._graphs)))
            else:                                             # second sighting: capture
                static = {k: ([t.clone() for t in v] if isinstance(v, list) else (None if v is None else v.clone()))
                          for k, v in static.items()}
                if static["image_stack"] is not None:             # the per-image static inputs are views of the stacked buffer
                    static["image"] = [static["image_stack"][b] for b in range(len(image))]
                if ragged_u8:                                     # the table holds the addresses of the (cloned) static images
                    static["img_tab"] = ops.image_table(static["image"], plan.sizes)
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                c0 = ops.L.launch_count
                with torch.cuda.graph(g):
                    out = self._forward(plan, static, want_seg, crf_one_sequence)
                out["static"] = True
                ent.update(graph=g, static=static, out=out, launches=ops.L.launch_count - c0)
                g.replay()
                self.graph_replays += 1
                self.kernel_launches += ent["launches"]
                return out
        c0 = ops.L.launch_count if not standins else 0
        out = self._forward(plan, static, want_seg, crf_one_sequence)
        if not standins:
            self.kernel_launches += ops.L.launch_count - c0
        return out

    def _forward(self, plan, st, want_seg, crf_one_sequence=False):
        """The kernel sequence of one forward over staged inputs (capturable: no host sync, no host-dependent control flow)."""
        net, pr = self.net, self._prep
        B = plan.B
        tab = st["tab"]
        dev = tab.device
        dt = {k: tab[s:s + n] for k, (s, n) in plan.offsets.items()}
        dt["ratios"] = dt["ratios"].view(torch.float32)
        seg_off = dt["seg_off"]
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        out = Intermediates({"plan": plan, "status": status})

        # Two independent branches run on two streams and join where the reference's data flow joins them (in a captured
        # graph these are parallel branches): every tensor-core kernel here is persistent with one CTA per SM, so a single
        # stream leaves SMs idle during each kernel's ramp-up and tail; a second stream's CTAs fill those slots.
        #   fork 1:  side = BERT encoder -> segment mean -> index map -> BERTgrid      main = transform -> stem .. stage-3 block 1
        #   fork 2:  side = auxiliary segmentation head (+ its loss kernel)             main = ROI-align -> late fusion -> heads
        main = torch.cuda.current_stream(dev) if dev.type == "cuda" else None
        side = self._side_stream(dev) if (main is not None and self.multi_stream) else None
        ps = self._ps()
        corpus, seg_ids = st["corpus"], st["seg_ids"]
        gs = net.early_fusion_downsampling_ratio

        # a1 transform (coords first: both branches read the boxes)
        boxes = ops.resize_coords(st["coors"], seg_off, dt["ratios"], B)
        out["boxes"] = boxes

        def bert_branch():
            # a2 / a3 BERT + segment aggregation, a4 BERTgrid
            hidden = self._bert(plan, dt, corpus)
            if st.get("mask") is not None:                          # input contract of `mask` (prefix of n_tok ones): status bit 2
                ops.mask_check(st["mask"], dt["tok_off"], status)
            seg_start = ops.segment_starts(seg_ids, dt["tok_off"], B, plan.K, status)
            seg_emb = ops.segment_reduce(hidden, dt["tok_row"], seg_start, plan.K,
                                         ops.AGG_MEAN if net.grid_mode == "mean" else ops.AGG_FIRST)
            out["seg_emb"] = seg_emb
            idx = ops.box_index_map(boxes, seg_off, B, gs, int(plan.H / gs), int(plan.W / gs))
            seg_emb_s = ops.to_split(seg_emb) if ps else None      # [K, 768]: scatter source and late-fusion operand
            out["seg_emb_split"] = seg_emb_s                       # kept alive: crosses streams
            out["index_map"], out["bertgrid"] = idx, ops.grid_scatter(seg_emb_s if ps else seg_emb, idx, seg_off)

        if side is not None:
            side.wait_stream(main)
            with torch.cuda.stream(side):
                bert_branch()
        else:
            bert_branch()
        batch = torch.zeros((B, plan.H + 6, plan.W + 6, 4), dtype=torch.float32, device=dev)   # zero-bordered NHWC4 stem input
        if st.get("image_stack") is not None:
            ops.normalize_resize_pad_batch(st["image_stack"], batch, 0, plan.sizes[0][0], plan.sizes[0][1], net.image_mean, net.image_std)
        elif st.get("img_tab") is not None:                        # differently-sized uint8 documents: one table-driven launch
            ops.decode_batch_u8(st["image"], st["img_tab"], batch, plan.sizes, net.image_mean, net.image_std)
        else:
            for b, im in enumerate(st["image"]):
                ops.normalize_resize_pad(im, batch, b, plan.sizes[b][0], plan.sizes[b][1], net.image_mean, net.image_std)
        out["image_batch"] = batch[:, 3:-3, 3:-3, :3]              # view without border / pad channel
        x1, x2 = self._backbone_pre(batch)                         # a5, the part before the early fusion
        if side is not None:
            main.wait_stream(side)
        seg_emb, seg_emb_s, grid = out["seg_emb"], dict.__getitem__(out, "seg_emb_split"), dict.__getitem__(out, "bertgrid")

        # a5 backbone from the early fusion on
        p_fuse = self._backbone_post(x1, x2, grid)
        out["p_fuse"] = p_fuse

        # a6 auxiliary segmentation head
        def seg_branch():
            out["pred_mask"], out["pred_ss"], lg = self._seg_head(p_fuse)
            out["seg_logits_lowres"] = lg
            cls_cat = st["cls"]
            if self.fuse_aux_loss and self._aux_loss_is_default():
                # labels are consumed in registers by the fused CE kernel; the int64 label maps are never written
                out["aux_ce"] = ops.seg_ce_loss(boxes, seg_off, cls_cat, lg, B, plan.H, plan.W,
                                                net.p_fuse_downsampling_ratio, 3)
            else:
                out["pos_neg_labels"], out["class_labels"] = ops.label_paint(boxes, seg_off, cls_cat, B, plan.H, plan.W)
            out["gt_label"] = cls_cat

        if want_seg:
            if side is not None:
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    seg_branch()
            else:
                seg_branch()

        # a7 ROI align, a8 late fusion
        roi = ops.roi_align(p_fuse, boxes, seg_off, 1.0 / float(net.p_fuse_downsampling_ratio), net.roi_shape, split_out=ps)
        out["roi"] = roi
        rn = net.late_fusion_net.ROI_embedding_net
        r = self._conv(roi, rn.conv_1, rn.bn_1, ACT_RELU)
        r = self._conv(r, rn.conv_2, rn.bn_2, ACT_RELU)
        roi_emb = self._lin(r.view(plan.K, -1), rn.linear, W=pr.misc["roi_fc_w"], W_split=pr.split.get("roi_fc_w"))
        late = self._lin(roi_emb, net.late_fusion_net.fuse_embedding_net.linear, A2=seg_emb_s if ps else seg_emb,
                         out_f32=True)
        late_s = ops.to_split(late) if ps else None      # [K, 1024]: the heads' tensor-core operand
        out["late"] = late

        # a9 / a10 field-type head
        head = net.field_type_classification_head
        if net.classifier_mode == "simp":
            logits = self._mlp_or_lin(late, head.category_classification_net, late_s)
            out["logits"] = logits
            out["pred_label"] = ops.softmax_rows(logits)
            if want_seg and hasattr(head, "pos_neg_classification_net"):
                out["pos_neg_logits"] = self._mlp_or_lin(late, head.pos_neg_classification_net, late_s)
        elif net.classifier_mode == "crf":
            logits = self._mlp_or_lin(late, head.category_classification_net, late_s)
            out["logits"] = logits
            # forward() decodes per document (field_type_classification_head.py:703-713); the label-free inference() entry
            # decodes all K rows of the batch as ONE sequence (:655-668) -- replicated, tags at document boundaries differ
            one = bool(crf_one_sequence)
            tags, scores = ops.crf_viterbi(logits, head.crf_layer.transitions.detach().contiguous(),
                                           dt["all_off"] if one else seg_off, 1 if one else B)
            out["pred_label"], out["crf_scores"] = tags[:, None], scores
        else:
            pn = self._mlp_or_lin(late, head.pos_neg_classification_net.layer, late_s)
            if "full_w" in pr.misc:
                cl = ops.gemm(late, pr.misc["full_w"], ep=make_epilogue(None, pr.misc["full_b"]), precision=self._prec())
            else:
                cl = torch.cat([self._mlp_or_lin(late, getattr(head, f"category_classification_net_{i}").layer, late_s)
                                for i in range(net.num_tokens - 1)], 1).contiguous()
            out["pos_neg_logits"], out["logits"] = pn, cl
            out["pred_label"] = ops.full_head_scores(pn.reshape(-1).contiguous(), cl)
        if side is not None and want_seg:
            main.wait_stream(side)
        return out
