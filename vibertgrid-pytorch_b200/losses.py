"""Host-side mirror of the reference's loss modules, applied to the logits the
kernels produce.

SURVEY.md section 8(f) rank 1 lists the losses as the NEXT row after the forward:
they are tiny tensors plus Python ``random`` sampling, so this round keeps them
as torch tensor ops on the device (plumbing), restated from the reference with
its exact sampling order so Python-``random`` consumption matches
(aux_loss_1, aux_loss_2, pos_neg, field-type):

  * CrossEntropyLossRandomSample  pipeline/custom_loss.py:35-101
  * CrossEntropyLossOHEM          pipeline/custom_loss.py:127-201
  * BCELossRandomSample           pipeline/custom_loss.py:228-290
  * BCELossOHEM                   pipeline/custom_loss.py:312-382
  * aux head wiring               model/semantic_segmentation_head.py:216-233, :343-352
  * main head wiring              model/field_type_classification_head.py:370-407, :564-588, :686-718

Quirks preserved: OHEM indexes the SORTED losses with ORIGINAL indices
(custom_loss.py:174-176); RandomSample accumulates in float64 and returns f64[1].
"""
from __future__ import annotations

import random

import torch
import torch.nn.functional as F


def _sample(loss, n):
    idx = torch.tensor(random.sample(range(int(loss.shape[0])), n), device=loss.device)
    return loss[idx]


def _random_sample_reduce(per_class_losses, sample_list, device):
    total = torch.zeros((1,), dtype=torch.float64, device=device)
    kept = 0
    for cur, want in zip(per_class_losses, sample_list):
        n = min(want, cur.shape[0])
        kept += n
        keep = _sample(cur, n) if n == want else cur
        total = total + keep.sum()
    total /= kept
    return total


def ce_random_sample(logits, target, sample_list, weight=None):
    if sample_list is None:
        return F.cross_entropy(logits.float(), target, weight=weight)
    ce = F.cross_entropy(logits.float(), target, weight=weight, reduction="none")
    if len(sample_list) == 2 and logits.shape[1] >= 2:
        masks = [target == 0, target != 0]
    else:
        assert len(sample_list) == logits.shape[1], "sample_list must have 2 or num-channel entries"
        masks = [target == c for c in range(len(sample_list))]
    return _random_sample_reduce([ce[m] for m in masks], sample_list, target.device)


def _ohem_reduce(loss, target_is_neg, n_pos, n_neg, rnd):
    pos, neg = loss[~target_is_neg], loss[target_is_neg]
    if rnd:
        if 2 * n_pos < pos.shape[0]:
            pos = _sample(pos, 2 * n_pos)
        if 2 * n_neg < neg.shape[0]:
            neg = _sample(neg, 2 * n_neg)
    sp, ip = torch.sort(pos, descending=True)
    kp = min(sp.shape[0], n_pos)
    if 0 < kp < sp.shape[0]:
        sp = sp[ip[:kp]]                 # reference quirk: original indices into the sorted array
    sn, in_ = torch.sort(neg, descending=True)
    kn = min(sn.shape[0], n_neg)
    if 0 < kn < sn.shape[0]:
        sn = sn[in_[:kn]]
    return (sp.sum() + sn.sum()) / (kp + kn)


def ce_ohem(logits, target, n_pos, n_neg, weight=None, rnd=False):
    if n_pos == -1 and n_neg == -1:
        return F.cross_entropy(logits.float(), target, weight=weight)
    ce = F.cross_entropy(logits.float(), target, weight=weight, reduction="none")
    return _ohem_reduce(ce, target == 0, n_pos, n_neg, rnd)


def bce_random_sample(logits, target, sample_list, weight=None):
    if logits.dim() == 2:
        logits = logits.squeeze(1)
    if sample_list is None:
        return F.binary_cross_entropy_with_logits(logits.float(), target, weight=weight)
    bce = F.binary_cross_entropy_with_logits(logits.float(), target, weight=weight, reduction="none")
    m = logits > 0
    return _random_sample_reduce([bce[~m], bce[m]], sample_list, target.device)


def bce_ohem(logits, target, n_pos, n_neg, weight=None, rnd=False):
    if n_pos == -1 and n_neg == -1:
        return F.binary_cross_entropy_with_logits(logits.float(), target, weight=weight)
    bce = F.binary_cross_entropy_with_logits(logits.float(), target, weight=weight, reduction="none")
    return _ohem_reduce(bce, target == 0, n_pos, n_neg, rnd)


def _w(net, like):
    return None if net.loss_weights is None else net.loss_weights.to(like.device)


def aux_loss(net, out, ctx=None):
    """``ctx``: a losses_device.SamplingCtx selects the sync-free device-side sampling (simp head); None = the reference's own
    host-side draws (Python ``random``), bit-compatible with its RNG consumption."""
    cfg = net.loss_cfg
    if "aux_ce" in out:                 # fused kernel (vbg_seg_ce_loss): [mean CE mask head, mean CE class head]
        return out["aux_ce"][0] + out["aux_ce"][1]
    pm, ps = out["pred_mask"], out["pred_ss"]
    n_pos, n_neg = cfg["aux"]
    if ctx is not None and net.classifier_mode == "simp":
        from . import losses_device as D
        l1 = D.ce_random_sample(pm.permute(0, 2, 3, 1).reshape(-1, pm.shape[1]), out["pos_neg_labels"].reshape(-1), cfg["aux_sample_list"], ctx=ctx)
        l2 = D.ce_ohem(ps.permute(0, 2, 3, 1).reshape(-1, ps.shape[1]), out["class_labels"].reshape(-1), n_pos, n_neg, weight=_w(net, ps), ctx=ctx)
        return l1 + l2
    l1 = ce_random_sample(pm, out["pos_neg_labels"], cfg["aux_sample_list"])
    if net.classifier_mode == "simp":
        l2 = ce_ohem(ps, out["class_labels"], n_pos, n_neg, weight=_w(net, ps))
        return l1 + l2
    head = net.semantic_segmentation_head
    l2 = torch.zeros((1,), device=pm.device)
    pos_mask = pm.softmax(1).argmax(1) == 1
    if int(pos_mask.int().sum()) != 0:
        for c in range(net.num_tokens - 1):
            conv = getattr(head, f"ss_binary_classifier_{c}").conv1
            pred = F.conv2d(ps, conv.weight, conv.bias)[pos_mask.unsqueeze(1)]
            lab = (out["class_labels"][pos_mask] == (c + 1)).float()
            l2 = l2 + bce_ohem(pred, lab, n_pos, n_neg, weight=_w(net, ps))
    return l1 + l2


def main_loss(net, out, ctx=None):
    cfg = net.loss_cfg
    label = out["gt_label"].long()
    if net.classifier_mode == "simp":
        p1, n1 = cfg["main_1"]
        p2, n2 = cfg["main_2"]
        if ctx is not None:
            from . import losses_device as D
            l_pn = D.ce_ohem(out["pos_neg_logits"], (label > 0).long(), p1, n1, rnd=cfg["random"], ctx=ctx) if net.add_pos_neg else None
            l_c = D.ce_ohem(out["logits"], label, p2, n2, weight=_w(net, out["logits"]), rnd=cfg["random"], ctx=ctx)
            return l_pn + l_c if net.add_pos_neg else l_c
        l_pn = ce_ohem(out["pos_neg_logits"], (label > 0).long(), p1, n1, rnd=cfg["random"])
        l_c = ce_ohem(out["logits"], label, p2, n2, weight=_w(net, out["logits"]), rnd=cfg["random"])
        return l_pn + l_c if net.add_pos_neg else l_c
    if net.classifier_mode == "crf":
        # eval branch: mean Viterbi path score (field_type_classification_head.py:701-718)
        return (out["crf_scores"].sum() / len(out["plan"].seg_counts)).reshape(1)
    p1, n1 = cfg["main_1"]
    p2, n2 = cfg["main_2"]
    pn = out["pos_neg_logits"].reshape(-1)
    # sample_list = [num_hard_negative_1, num_hard_positive_1]; the reference's "[s0, s0]" rewrite at
    # custom_loss.py:222-223 only rebinds a local, so self.sample_list keeps both entries
    l_pn = bce_random_sample(pn, (label > 0).float(), [n1, p1])
    gate = pn.detach().sigmoid().ge(0.5)
    l_c = torch.zeros((1,), device=pn.device)
    if int(gate.sum()) != 0:
        for c in range(net.num_tokens - 1):
            l_c = l_c + bce_ohem(out["logits"][gate][:, c], (label[gate] == (c + 1)).float(), p2, n2,
                                 weight=_w(net, pn), rnd=cfg["random"])
    return l_pn + l_c
