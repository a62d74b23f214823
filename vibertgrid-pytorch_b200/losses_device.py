"""The reference's sampled / OHEM losses (pipeline/custom_loss.py:9-382) WITHOUT host randomness or device->host syncs.

The reference (and losses.py, its restatement) masks the per-element losses by class (``ce[mask]``: a data-dependent shape, i.e.
a sync), draws indices with Python's ``random.sample`` (host RNG, needs the population size: another sync) and uploads them.
Here every step has a FIXED shape and runs on the device, so the loss tail can be recorded into the training step's CUDA graph:

  * sampling keys: ``key[i] = hash(seed, step_seed, i)`` (``vbg_uniform_keys``; 31 bits).  RNG-ORDER CONTRACT -- the one
    deliberate deviation from the reference: "a uniform random k-subset of a group" is realised as "the k members with the
    smallest keys", "in random order" as "in ascending key order".  Same distribution, different draws than ``random.sample``
    (bit-compatibility with Python's Mersenne twister would need the population sizes on the host).
  * ``x[mask]`` followed by top-k / sort becomes: keys (or losses) of non-members set to +-infinity, ``torch.topk`` / ``torch.sort``
    over the full fixed-size array, validity masks instead of shorter tensors, ``cumsum`` for positions inside the compacted array.
  * the OHEM quirk of the reference (``sorted_loss[sorted_index[:k]]``: ORIGINAL indices used as positions in the SORTED array,
    custom_loss.py:174-176) is reproduced exactly -- with it, "hard example mining" after random pre-sampling is in fact a second
    uniform sub-sample, and without pre-sampling it picks the sorted losses at the compacted positions of the k hardest.

Every function equals its losses.py twin exactly whenever no random draw happens (populations not larger than the sample
sizes; OHEM with ``random=False``), which is how tests/test_losses_device.py pins the restatement; where a draw happens the
tests check the kept counts, the weights and the subset property.  The index selection uses ``torch.topk`` / ``torch.sort``
(library radix select / sort over at most B*H*W floats); the BASELINE configurations use the plain mean cross entropy, which
is the fused ``vbg_seg_ce_loss`` kernel -- these functions serve the reference's ``example_config.yaml`` loss settings.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

_BIG = 1 << 40


class SamplingCtx:
    """Per-forward source of sampling keys: call sites are numbered in order, so every loss term of a step draws its own keys."""

    def __init__(self, step_seed=None, base_seed=0x5EED, keys_fn=None):
        self.step_seed, self.base_seed, self.calls, self.keys_fn = step_seed, int(base_seed), 0, keys_fn

    def keys(self, n, device):
        self.calls += 1
        if self.keys_fn is not None:                      # tests inject keys (CPU)
            return self.keys_fn(n, self.calls).to(device)
        from . import ops
        return ops.uniform_keys(n, self.base_seed + 0x9E37 * self.calls, self.step_seed, device)


def _smallest_k_mask(key, k):
    """bool [N]: the k members with the smallest keys (all members when there are fewer; non-members carry _BIG)."""
    n = key.numel()
    k = min(int(k), n)
    if k <= 0:
        return torch.zeros(n, dtype=torch.bool, device=key.device)
    vals, idx = torch.topk(key, k, largest=False, sorted=False)
    return torch.zeros(n, dtype=torch.bool, device=key.device).scatter_(0, idx, vals < _BIG)


def _random_sample_reduce(loss, masks, sample_list, ctx):
    """custom_loss.py:63-92 (reduction 'mean'): per group keep a uniform subset of ``want`` members when there are at least
    ``want``, else all; mean over everything kept -> float64 [1] like the reference."""
    loss = loss.reshape(-1)
    total = torch.zeros((), dtype=torch.float64, device=loss.device)
    kept = torch.zeros((), dtype=torch.int64, device=loss.device)
    for m, want in zip(masks, sample_list):
        m = m.reshape(-1)
        key = torch.where(m, ctx.keys(loss.numel(), loss.device).long(), torch.full((), _BIG, dtype=torch.int64, device=loss.device))
        keep = _smallest_k_mask(key, want)
        total = total + torch.where(keep, loss, torch.zeros((), dtype=loss.dtype, device=loss.device)).double().sum()
        kept = kept + torch.clamp(m.sum(), max=int(want))
    return (total / kept).reshape(1)


def ce_random_sample(logits, target, sample_list, weight=None, ctx=None):
    if sample_list is None:
        return F.cross_entropy(logits.float(), target, weight=weight)
    ce = F.cross_entropy(logits.float(), target, weight=weight, reduction="none")
    if len(sample_list) == 2 and logits.shape[1] >= 2:
        masks = [target == 0, target != 0]
    else:
        assert len(sample_list) == logits.shape[1], "sample_list must have 2 or num-channel entries"
        masks = [target == c for c in range(len(sample_list))]
    return _random_sample_reduce(ce, masks, sample_list, ctx)


def bce_random_sample(logits, target, sample_list, weight=None, ctx=None):
    if logits.dim() == 2:
        logits = logits.squeeze(1)
    if sample_list is None:
        return F.binary_cross_entropy_with_logits(logits.float(), target, weight=weight)
    bce = F.binary_cross_entropy_with_logits(logits.float(), target, weight=weight, reduction="none")
    m = logits > 0
    return _random_sample_reduce(bce, [~m, m], sample_list, ctx)


def _ohem_group(loss, member, n_hard, rnd, ctx):
    """One class of custom_loss.py:154-185 -> (sum of the kept losses, number kept [int64 0-d])."""
    N = loss.numel()
    dev = loss.device
    zero = torch.zeros((), dtype=loss.dtype, device=dev)
    neg_inf = torch.full((), float("-inf"), dtype=loss.dtype, device=dev)
    cnt = member.sum()
    n_hard = int(n_hard)
    assert n_hard > 0, "device OHEM: positive hard-example counts only (losses.py handles the reference's degenerate settings)"
    if rnd:
        cap = min(2 * max(n_hard, 0), N)
        if cap == 0:
            return zero.double(), torch.zeros((), dtype=torch.int64, device=dev)
        sampled = cnt > 2 * n_hard                                                   # device bool: a random pre-sample happens
        order = torch.where(sampled, ctx.keys(N, dev).long(), torch.arange(N, device=dev))      # random order | original order
        key = torch.where(member, order, torch.full((), _BIG, dtype=torch.int64, device=dev))
        vals, sel = torch.topk(key, cap, largest=False, sorted=True)
        valid = vals < _BIG
        pre = torch.where(valid, loss[sel], neg_inf)                                  # the array the reference sorts, in its order
        s = valid.sum()
        sp, ip = torch.sort(pre, descending=True)
        kp = torch.clamp(s, max=n_hard)
        j = torch.arange(min(n_hard, cap), device=dev)
        quirk = torch.where(j < kp, sp[ip[:j.numel()]], zero).double().sum()          # sorted[original index of the j-th hardest]
        everything = torch.where(torch.arange(cap, device=dev) < s, sp, zero).double().sum()
        return torch.where(kp < s, quirk, everything), kp
    # no pre-sampling: the reference sorts every member
    full = torch.where(member, loss, neg_inf)
    sp, ip = torch.sort(full, descending=True)                                        # members first; ip = positions in the FULL array
    cpos = torch.cumsum(member.long(), 0) - 1                                         # position of an element inside loss[member]
    kp = torch.clamp(cnt, max=n_hard)
    j = torch.arange(min(n_hard, N), device=dev)
    quirk = torch.where(j < kp, sp[cpos[ip[:j.numel()]].clamp_min(0)], zero).double().sum()
    everything = torch.where(member, loss, zero).double().sum()
    return torch.where(kp < cnt, quirk, everything), kp


def _ohem_reduce(loss, target_is_neg, n_pos, n_neg, rnd, ctx):
    loss, target_is_neg = loss.reshape(-1), target_is_neg.reshape(-1)
    sp, kp = _ohem_group(loss, ~target_is_neg, n_pos, rnd, ctx)
    sn, kn = _ohem_group(loss, target_is_neg, n_neg, rnd, ctx)
    return ((sp + sn) / (kp + kn)).to(loss.dtype)


def ce_ohem(logits, target, n_pos, n_neg, weight=None, rnd=False, ctx=None):
    if n_pos == -1 and n_neg == -1:
        return F.cross_entropy(logits.float(), target, weight=weight)
    ce = F.cross_entropy(logits.float(), target, weight=weight, reduction="none")
    return _ohem_reduce(ce, target == 0, n_pos, n_neg, rnd, ctx)


def bce_ohem(logits, target, n_pos, n_neg, weight=None, rnd=False, ctx=None):
    if n_pos == -1 and n_neg == -1:
        return F.binary_cross_entropy_with_logits(logits.float(), target, weight=weight)
    bce = F.binary_cross_entropy_with_logits(logits.float(), target, weight=weight, reduction="none")
    return _ohem_reduce(bce, target == 0, n_pos, n_neg, rnd, ctx)


def supported(cfg) -> bool:
    """True when every sampled / OHEM setting of ``net.loss_cfg`` is one these fixed-shape forms take (counts > 0, or the
    (-1, -1) pair = plain mean)."""
    def pair_ok(p):
        p = tuple(int(v) for v in p)
        return p == (-1, -1) or (p[0] > 0 and p[1] > 0)
    sl = cfg["aux_sample_list"]
    return pair_ok(cfg["main_1"]) and pair_ok(cfg["main_2"]) and pair_ok(cfg["aux"]) and (sl is None or all(int(v) > 0 for v in sl))


def two_stage_aux_default(net, pm, ps, pos_neg_labels, class_labels):
    """The auxiliary loss of the two-stage segmentation head (``full`` / ``crf`` classifier modes,
    model/semantic_segmentation_head.py:216-233) for the plain-mean loss configuration, WITHOUT its device->host sync and its
    boolean-mask gathers: stage 1 = mean cross entropy of the 3-way mask head; stage 2 = for every class c the binary cross
    entropy of ``ss_binary_classifier_c`` over the pixels whose PREDICTED mask class is 1 -- here as sum(bce * m) / sum(m) over
    all pixels (m the 0/1 selection), which is the same mean up to the order of the sum, and 0 when nothing is selected (the
    reference skips the stage then).  Fixed shapes, no host reads: capturable in the training step's CUDA graph.  The 1x1
    classifier convolutions are per-pixel dot products over <= C channels, written as a broadcast multiply-add (no library conv)."""
    head = net.semantic_segmentation_head
    l1 = F.cross_entropy(pm.float(), pos_neg_labels)
    m = (pm.softmax(1).argmax(1) == 1).float()                                   # [B, H, W]
    cnt = m.sum()
    l2 = torch.zeros((1,), device=pm.device)
    for c in range(net.num_tokens - 1):
        conv = getattr(head, f"ss_binary_classifier_{c}").conv1
        w = conv.weight.reshape(1, -1, 1, 1)
        pred = (ps * w).sum(1) + conv.bias.reshape(1, 1, 1)                      # [B, H, W]
        lab = (class_labels == (c + 1)).float()
        bce = F.binary_cross_entropy_with_logits(pred.float(), lab, reduction="none")
        l2 = l2 + (bce * m).sum() / cnt.clamp_min(1.0)
    return l1 + l2

