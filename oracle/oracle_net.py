"""ORACLE (test infrastructure only) -- fp32 CPU restatement of
``ViBERTgridNet.forward`` (reference model/ViBERTgrid_net.py:501-544) as a
pure function of a reference-layout state dict.

Floating-point stages use plain ``torch.nn.functional`` fp32 ops (the same
third-party arithmetic the reference reaches: conv / BN / linear / LN / softmax
/ GELU / interpolate); integer stages use ``oracle_ops`` (numpy).  No
HuggingFace, no torchvision: the BERT encoder and ROIAlign are restated from
their published algorithms so this file travels to the GPU box.

Pinned against the live reference by ``oracle/make_golden.py`` ->
``tests/golden/*.npz`` (see oracle_ops.py header).  NOT importable from the
product package.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

from . import oracle_ops as ops

BN_EPS, LN_EPS = 1e-5, 1e-12


@dataclass
class OracleConfig:
    backbone: str = "resnet_18_fpn"
    classifier_mode: str = "simp"        # simp | full | crf
    num_classes: int = 5
    image_mean: Sequence[float] = (0.9248, 0.9224, 0.9215)
    image_std: Sequence[float] = (0.1532, 0.1545, 0.1536)
    min_size: int = 512                   # test_image_min_size (eval) -- transform.py:196
    max_size: int = 800
    grid_mode: str = "mean"
    stride: int = 8
    roi_shape: int = 7
    p_fuse_stride: int = 4
    num_heads: int = 12
    layer_mode: str = "single"
    with_seg_head: bool = True
    ln_eps: float = LN_EPS                # BertConfig.layer_norm_eps (roberta-base: 1e-5)
    roberta_pad: Optional[int] = None     # RobertaModel: padding_idx of its position-id rule; None = BertModel positions


def _bn(x, sd, p):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"],
                        sd[p + ".bias"], False, 0.0, BN_EPS)


def _conv(x, sd, p, stride=1, pad=0):
    return F.conv2d(x, sd[p + ".weight"], sd.get(p + ".bias"), stride=stride, padding=pad)


def _lin(x, sd, p):
    return F.linear(x, sd[p + ".weight"], sd[p + ".bias"])


# ----------------------------------------------------------------------------- a1
def transform(images, coors, cfg: OracleConfig):
    """pipeline/transform.py:273-312 (eval branch: size = test_min_size)."""
    mean = torch.tensor(cfg.image_mean, dtype=torch.float32)[:, None, None]
    std = torch.tensor(cfg.image_std, dtype=torch.float32)[:, None, None]
    outs, new_coors, sizes = [], [], []
    for img, c in zip(images, coors):
        img = (img.float() - mean) / std                                   # :122
        h, w = img.shape[-2:]
        scale = ops.resize_scale(h, w, float(cfg.min_size), float(cfg.max_size))
        img = F.interpolate(img[None], scale_factor=scale, mode="bilinear",
                            recompute_scale_factor=True, align_corners=False)[0]   # :149-155
        nh, nw = img.shape[-2:]
        assert (nh, nw) == ops.resized_shape(h, w, scale)
        new_coors.append(ops.resize_coords(c.numpy(), (h, w), (nh, nw)))
        outs.append(img)
        sizes.append((nh, nw))
    H, W = ops.padded_shape(sizes)
    batch = torch.zeros(len(outs), 3, H, W)
    for i, img in enumerate(outs):
        batch[i, :, : img.shape[1], : img.shape[2]] = img                  # :269
    return batch, new_coors, sizes


# ----------------------------------------------------------------------------- a2
def bert_encoder(sd: Dict[str, torch.Tensor], prefix: str, ids: torch.Tensor, attn_mask: torch.Tensor,
                 num_heads: int, ln_eps: float = LN_EPS, roberta_pad: Optional[int] = None) -> torch.Tensor:
    """Post-LN BERT encoder in eval mode (dropout off) -- the computation
    HuggingFace ``BertModel(input_ids, attention_mask).last_hidden_state`` performs,
    as called at model/BERTgrid_generator.py:134-135."""
    e = prefix + "embeddings."
    B, T = ids.shape
    if roberta_pad is None:
        pos_emb = sd[e + "position_embeddings.weight"][:T][None]
    else:
        # transformers RobertaEmbeddings.create_position_ids_from_input_ids: cumsum over non-pad ids, offset by padding_idx
        nz = (ids != roberta_pad).long()
        pos_emb = sd[e + "position_embeddings.weight"][torch.cumsum(nz, 1) * nz + roberta_pad]
    x = sd[e + "word_embeddings.weight"][ids] + sd[e + "token_type_embeddings.weight"][0] + pos_emb
    x = F.layer_norm(x, x.shape[-1:], sd[e + "LayerNorm.weight"], sd[e + "LayerNorm.bias"], ln_eps)
    add_mask = (1.0 - attn_mask[:, None, None, :].float()) * torch.finfo(torch.float32).min
    n_layers = 1 + max(int(k.split("encoder.layer.")[1].split(".")[0]) for k in sd if k.startswith(prefix + "encoder.layer."))
    hd = x.shape[-1] // num_heads
    for i in range(n_layers):
        p = f"{prefix}encoder.layer.{i}."
        q = _lin(x, sd, p + "attention.self.query").view(B, T, num_heads, hd).transpose(1, 2)
        k = _lin(x, sd, p + "attention.self.key").view(B, T, num_heads, hd).transpose(1, 2)
        v = _lin(x, sd, p + "attention.self.value").view(B, T, num_heads, hd).transpose(1, 2)
        s = q @ k.transpose(-1, -2) / math.sqrt(hd) + add_mask
        ctx = (s.softmax(-1) @ v).transpose(1, 2).reshape(B, T, -1)
        a = _lin(ctx, sd, p + "attention.output.dense") + x
        x = F.layer_norm(a, a.shape[-1:], sd[p + "attention.output.LayerNorm.weight"],
                         sd[p + "attention.output.LayerNorm.bias"], ln_eps)
        h = F.gelu(_lin(x, sd, p + "intermediate.dense"))
        o = _lin(h, sd, p + "output.dense") + x
        x = F.layer_norm(o, o.shape[-1:], sd[p + "output.LayerNorm.weight"], sd[p + "output.LayerNorm.bias"], ln_eps)
    return x


def bert_token_embeddings(sd, corpus, mask, cfg: OracleConfig):
    """model/BERTgrid_generator.py:81-146 -- windowed BERT, strip [CLS], concat."""
    outs = []
    for ids, am, n in ops.bert_windows(corpus.numpy(), mask.numpy()):
        h = bert_encoder(sd, "bert_model.", torch.from_numpy(ids), torch.from_numpy(am), cfg.num_heads, cfg.ln_eps,
                         cfg.roberta_pad)
        outs.append(h[:, 1:1 + n])
    return torch.cat(outs, 1)            # [B, L, 768]


# ----------------------------------------------------------------------------- a5
def _block(x, sd, p, downsample, d_variant):
    """BasicBlock / DBlock -- model/ResNetFPN_ViBERTgrid.py:174-184, :259-269."""
    y = F.relu(_bn(_conv(x, sd, p + "conv_1", 2 if downsample else 1, 1), sd, p + "bn_1"))
    y = _bn(_conv(y, sd, p + "conv_2", 1, 1), sd, p + "bn_2")
    if downsample:
        if d_variant:
            sc = _bn(_conv(F.avg_pool2d(x, 2, 2), sd, p + "conv_shortcut.1"), sd, p + "conv_shortcut.2")
        else:
            sc = _bn(_conv(x, sd, p + "conv_shortcut.0", 2, 0), sd, p + "conv_shortcut.1")
    else:
        sc = x
    return F.relu(y + sc)


def _tv_block(x, sd, p, stride):
    y = F.relu(_bn(_conv(x, sd, p + "conv1", stride, 1), sd, p + "bn1"))
    y = _bn(_conv(y, sd, p + "conv2", 1, 1), sd, p + "bn2")
    if (p + "downsample.0.weight") in sd:
        x = _bn(_conv(x, sd, p + "downsample.0", stride, 0), sd, p + "downsample.1")
    return F.relu(y + x)


def _n_blocks(sd, prefix):
    return 1 + max(int(k[len(prefix):].split(".")[0]) for k in sd if k.startswith(prefix) and k[len(prefix)].isdigit())


def backbone(sd, x, grid, cfg: OracleConfig):
    """model/ResNetFPN_ViBERTgrid.py:478-508 (plain / D) and :612-648 (pretrained layout)."""
    b = "backbone."
    if cfg.backbone.endswith("_pretrained"):
        r = b + "resnet."
        x1 = F.relu(_bn(_conv(x, sd, r + "conv1", 2, 3), sd, r + "bn1"))
        x1 = F.max_pool2d(x1, 3, 2, 1)
        for i in range(_n_blocks(sd, r + "layer1.")):
            x1 = _tv_block(x1, sd, f"{r}layer1.{i}.", 1)
        x2 = _tv_block(x1, sd, r + "layer2.0.", 2)
        x2 = _conv(torch.cat((x2, grid), 1), sd, b + "early_fusion")
        for i in range(1, _n_blocks(sd, r + "layer2.")):
            x2 = _tv_block(x2, sd, f"{r}layer2.{i}.", 1)
        x3 = x2
        for i in range(_n_blocks(sd, r + "layer3.")):
            x3 = _tv_block(x3, sd, f"{r}layer3.{i}.", 2 if i == 0 else 1)
        x4 = x3
        for i in range(_n_blocks(sd, r + "layer4.")):
            x4 = _tv_block(x4, sd, f"{r}layer4.{i}.", 2 if i == 0 else 1)
    else:
        d = "_D_" in cfg.backbone
        x1 = F.relu(_bn(_conv(x, sd, b + "conv_1.0", 2, 3), sd, b + "conv_1.1"))
        x1 = F.max_pool2d(x1, 3, 2, 1)
        for i in range(_n_blocks(sd, b + "conv_2_x.")):
            x1 = _block(x1, sd, f"{b}conv_2_x.{i}.", False, d)
        x2 = _block(x1, sd, b + "conv_3_x.block_1.", True, d)
        x2 = _conv(torch.cat((x2, grid), 1), sd, b + "conv_3_x.early_fusion")      # :317-318
        if any(k.startswith(b + "conv_3_x.layers.") for k in sd):
            for i in range(_n_blocks(sd, b + "conv_3_x.layers.")):
                x2 = _block(x2, sd, f"{b}conv_3_x.layers.{i}.", False, d)
        x3 = x2
        for i in range(_n_blocks(sd, b + "conv_4_x.")):
            x3 = _block(x3, sd, f"{b}conv_4_x.{i}.", i == 0, d)
        x4 = x3
        for i in range(_n_blocks(sd, b + "conv_5_x.")):
            x4 = _block(x4, sd, f"{b}conv_5_x.{i}.", i == 0, d)
    up = lambda t, s: F.interpolate(t, scale_factor=s, mode="nearest")
    x4 = _conv(x4, sd, b + "conv_6_x")
    x5 = _conv(up(x4, 2) + _conv(x3, sd, b + "skip_1"), sd, b + "merge_1", 1, 1)
    x6 = _conv(up(x5, 2) + _conv(x2, sd, b + "skip_2"), sd, b + "merge_2", 1, 1)
    x7 = _conv(up(x6, 2) + _conv(x1, sd, b + "skip_3"), sd, b + "merge_3", 1, 1)
    return _conv(torch.cat([up(x4, 8), up(x5, 4), up(x6, 2), x7], 1), sd, b + "fuse")


# ----------------------------------------------------------------------------- a6
def seg_head(sd, p_fuse, cfg: OracleConfig):
    """SemanticSegmentationEncoder.forward -- semantic_segmentation_head.py:66-78."""
    p = "semantic_segmentation_head." + ("semantic_segmentation_encoder." if cfg.classifier_mode == "simp" else "ss_encoder.")
    x = F.relu(_bn(_conv(p_fuse, sd, p + "conv_1", 1, 1), sd, p + "bn_1"))
    x = F.relu(_bn(_conv(x, sd, p + "conv_2", 1, 1), sd, p + "bn_2"))
    x = F.interpolate(x, scale_factor=4, mode="nearest")
    return _conv(x, sd, p + "conv_3_1"), _conv(x, sd, p + "conv_3_2")


# ----------------------------------------------------------------------------- a8 / a9
def late_fusion(sd, roi, seg_emb_cat):
    """field_type_classification_head.py:64-75, :164-190."""
    p = "late_fusion_net.ROI_embedding_net."
    x = F.relu(_bn(_conv(roi, sd, p + "conv_1", 1, 1), sd, p + "bn_1"))
    x = F.relu(_bn(_conv(x, sd, p + "conv_2", 1, 1), sd, p + "bn_2"))
    x = _lin(x.flatten(1), sd, p + "linear")
    return _lin(torch.cat((x, seg_emb_cat), 1), sd, "late_fusion_net.fuse_embedding_net.linear")


def _mlp_or_lin(x, sd, p):
    if (p + "linear_1.weight") in sd:
        return _lin(F.relu(_lin(x, sd, p + "linear_1")), sd, p + "linear_2")
    return _lin(x, sd, p + "linear")


def head_logits(sd, late, cfg: OracleConfig):
    """simp: :564-571 ; crf: :683-684 ; full: :370-400 (eval scores)."""
    h = "field_type_classification_head."
    out = {}
    if cfg.classifier_mode in ("simp", "crf"):
        out["logits"] = _mlp_or_lin(late, sd, h + "category_classification_net.")
        if (h + "pos_neg_classification_net.linear_1.weight") in sd:
            out["pos_neg_logits"] = _mlp_or_lin(late, sd, h + "pos_neg_classification_net.")
    else:
        pn = _mlp_or_lin(late, sd, h + "pos_neg_classification_net.layer.").squeeze(1)
        out["pos_neg_logits"] = pn
        C = cfg.num_classes
        cls = torch.stack([_mlp_or_lin(late, sd, f"{h}category_classification_net_{i}.layer.").squeeze(1)
                           for i in range(C - 1)], 1)
        out["logits"] = cls
        gate = pn.sigmoid().ge(0.5)
        pred = torch.zeros(late.shape[0], C)
        pred[:, 0] = pn.sigmoid()
        pred[:, 1:] = torch.where(gate[:, None], cls.sigmoid(), torch.zeros_like(cls))
        out["pred_label"] = pred
    return out


# ----------------------------------------------------------------------------- a11
@torch.no_grad()
def forward(sd: Dict[str, torch.Tensor], cfg: OracleConfig, image, seg_indices, seg_classes, coors, corpus, mask):
    """Eval-mode joint forward; returns every intermediate the parity gates name
    (SURVEY 8d "Parity gates").  All tensors NCHW fp32 like the reference."""
    torch.set_grad_enabled(False)
    out = {}
    batch, coors_t, sizes = transform(image, coors, cfg)
    H, W = batch.shape[-2:]
    out["image_batch"], out["coors_t"], out["image_sizes"] = batch, coors_t, sizes

    tok = bert_token_embeddings(sd, corpus, mask, cfg)
    seg_emb = []
    for b in range(len(image)):
        valid = tok[b][mask[b] == 1].numpy()                                   # BERTgrid_generator.py:151
        assert valid.shape[0] == seg_indices[b].shape[0]
        seg_emb.append(ops.segment_aggregate(valid, seg_indices[b].numpy(), cfg.grid_mode))
        assert seg_emb[-1].shape[0] == coors_t[b].shape[0]
    out["seg_emb"] = seg_emb
    idx = ops.box_index_map(coors_t, H, W, cfg.stride)
    out["index_map"] = idx
    grid = torch.from_numpy(ops.scatter_grid(seg_emb, idx))
    out["bertgrid"] = grid

    p_fuse = backbone(sd, batch, grid, cfg)
    out["p_fuse"] = p_fuse

    if cfg.with_seg_head:
        out["pred_mask"], out["pred_ss"] = seg_head(sd, p_fuse, cfg)
        idx1 = ops.box_index_map(coors_t, H, W, 1)
        out["pos_neg_labels"], out["class_labels"] = ops.paint_labels(idx1, [c.numpy() for c in seg_classes])

    boxes = np.concatenate([c.astype(np.float32) for c in coors_t], 0)          # grid_roi_align.py:72-74
    bidx = np.concatenate([np.full(c.shape[0], b, np.int32) for b, c in enumerate(coors_t)])
    roi, grids = ops.roi_align(p_fuse.numpy(), boxes, bidx, 1.0 / cfg.p_fuse_stride, cfg.roi_shape)
    out["roi"], out["roi_sample_grid"] = torch.from_numpy(roi), grids

    late = late_fusion(sd, out["roi"], torch.from_numpy(np.concatenate(seg_emb, 0)))
    out["late"] = late
    out.update(head_logits(sd, late, cfg))
    if cfg.classifier_mode == "simp":
        out["pred_label"] = out["logits"].softmax(1)                               # :581
    elif cfg.classifier_mode == "crf":
        T = cfg.num_classes + 2
        trans = sd["field_type_classification_head.crf_layer.transitions"].numpy()
        tags, off = [], 0
        for c in coors_t:
            _, path = ops.crf_viterbi(out["logits"][off:off + c.shape[0]].numpy(), trans, T - 2, T - 1)
            tags += path
            off += c.shape[0]
        out["pred_label"] = torch.tensor(tags, dtype=torch.float32)[:, None]
    out["gt_label"] = torch.cat([c for c in seg_classes]).int()
    return out
