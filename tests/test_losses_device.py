"""losses_device.py (the sampled / OHEM losses of pipeline/custom_loss.py as fixed-shape, sync-free device code) against
losses.py (the host-side restatement pinned to the reference's fixtures): EXACT agreement wherever no random draw happens
(populations within the sample sizes; OHEM with random=False incl. the reference's sorted[original-index] quirk), and the
sampling semantics where one does (kept counts, equal weights, subset of the right group, new draw per key set).
CPU: sampling keys are injected (the product draws them with vbg_uniform_keys on the GPU)."""
import random

import pytest
import torch

import vibertgrid_pytorch_b200  # noqa: F401
from vibertgrid_pytorch_b200 import losses as H
from vibertgrid_pytorch_b200 import losses_device as D


def ctx(seed=0):
    def keys(n, call):
        g = torch.Generator().manual_seed(1000 * seed + call)
        return torch.randint(0, 2 ** 31 - 1, (n,), generator=g, dtype=torch.int64).to(torch.int32)
    return D.SamplingCtx(keys_fn=keys)


def data(n, c, seed, frac_zero=0.6):
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn(n, c, generator=g, requires_grad=True)
    target = torch.randint(1, c, (n,), generator=g)
    target[torch.rand(n, generator=g) < frac_zero] = 0
    return logits, target


@pytest.mark.parametrize("n,c,npos,nneg", [(40, 5, 3, 4), (40, 5, 30, 30), (7, 2, 2, 2), (300, 4, 16, 16), (50, 3, 1, 60)])
def test_ohem_without_presampling_equals_host_exactly(n, c, npos, nneg):
    logits, target = data(n, c, n)
    w = torch.rand(c) + 0.5
    for weight in (None, w):
        a = H.ce_ohem(logits, target, npos, nneg, weight=weight, rnd=False)
        b = D.ce_ohem(logits, target, npos, nneg, weight=weight, rnd=False, ctx=ctx())
        assert abs(float(a) - float(b)) <= 2e-6 * max(1.0, abs(float(a)))
        ga, = torch.autograd.grad(a, logits)
        gb, = torch.autograd.grad(b, logits)
        assert float((ga - gb).abs().max()) <= 2e-6
    la, tb = logits[:, 1], (target > 0).float()
    a = H.bce_ohem(la, tb, npos, nneg, rnd=False)
    b = D.bce_ohem(la, tb, npos, nneg, rnd=False, ctx=ctx())
    assert abs(float(a) - float(b)) <= 2e-6 * max(1.0, abs(float(a)))


@pytest.mark.parametrize("n,c,npos,nneg", [(20, 5, 16, 16), (50, 4, 16, 32), (9, 2, 8, 8)])
def test_ohem_random_with_small_populations_equals_host_exactly(n, c, npos, nneg):
    """random=True, but no class exceeds 2 x its count: the reference draws nothing and sorts in original order."""
    logits, target = data(n, c, 7 * n)
    assert int((target == 0).sum()) <= 2 * nneg and int((target != 0).sum()) <= 2 * npos
    random.seed(0)
    a = H.ce_ohem(logits, target, npos, nneg, rnd=True)
    b = D.ce_ohem(logits, target, npos, nneg, rnd=True, ctx=ctx())
    assert abs(float(a) - float(b)) <= 2e-6 * max(1.0, abs(float(a)))


def test_random_sample_small_populations_equal_host_and_large_ones_keep_exact_counts():
    logits, target = data(200, 3, 5, frac_zero=0.5)
    # every class below its sample size: nothing is drawn
    a = H.ce_random_sample(logits, target, [400, 400, 400])
    b = D.ce_random_sample(logits, target, [400, 400, 400], ctx=ctx())
    assert b.dtype == torch.float64 and tuple(b.shape) == (1,)
    assert abs(float(a) - float(b)) <= 1e-6
    # classes above their sample size: exactly `want` members kept, each with weight 1 / total kept
    want = [10, 20, 5]
    ce = torch.nn.functional.cross_entropy(logits, target, reduction="none").detach().requires_grad_()
    masks = [target == k for k in range(3)]
    loss = D._random_sample_reduce(ce, masks, want, ctx(1))
    (wgt,) = torch.autograd.grad(loss, ce)
    kept_total = sum(min(int(m.sum()), k) for m, k in zip(masks, want))
    for m, k in zip(masks, want):
        kept = (wgt[m] != 0)
        assert int(kept.sum()) == min(int(m.sum()), k)
    assert torch.allclose(wgt[wgt != 0], torch.full_like(wgt[wgt != 0], 1.0 / kept_total))
    assert abs(float(loss) - float((ce.double() * wgt.double()).sum())) < 1e-6
    (wgt2,) = torch.autograd.grad(D._random_sample_reduce(ce, masks, want, ctx(2)), ce)
    assert not torch.equal(wgt != 0, wgt2 != 0)                       # other keys, another subset
    # two-group form (background / foreground) and the BCE variant
    b2 = D.ce_random_sample(logits, target, [30, 30], ctx=ctx(3))
    assert torch.isfinite(b2).all()
    lb = logits[:, 0]
    assert torch.isfinite(D.bce_random_sample(lb, (target > 0).float(), [15, 15], ctx=ctx(4))).all()


def test_ohem_random_presampling_semantics():
    """Populations above 2 x count: a uniform pre-sample of 2 x count in random order, sorted, then the reference's quirk --
    the kept set is `count` members of the pre-sample, each with weight 1 / (kept_pos + kept_neg)."""
    logits, target = data(2000, 4, 11, frac_zero=0.7)
    npos, nneg = 16, 32
    ce = torch.nn.functional.cross_entropy(logits, target, reduction="none").detach().requires_grad_()
    loss = D._ohem_reduce(ce, target == 0, npos, nneg, True, ctx(5))
    (wgt,) = torch.autograd.grad(loss, ce)
    assert int((wgt[target != 0] != 0).sum()) == npos and int((wgt[target == 0] != 0).sum()) == nneg
    assert torch.allclose(wgt[wgt != 0], torch.full_like(wgt[wgt != 0], 1.0 / (npos + nneg)))
    sets = set()
    for s in range(6):
        (w2,) = torch.autograd.grad(D._ohem_reduce(ce, target == 0, npos, nneg, True, ctx(20 + s)), ce)
        sets.add(tuple((w2 != 0).nonzero().flatten().tolist()))
    assert len(sets) == 6


def test_supported_settings():
    base = dict(main_1=(16, 16), main_2=(32, 32), aux=(256, 256), aux_sample_list=[256, 512, 256], random=True)
    assert D.supported(base)
    assert D.supported({**base, "main_1": (-1, -1), "aux_sample_list": None})
    assert not D.supported({**base, "aux": (-1, 256)})
    assert not D.supported({**base, "aux_sample_list": [0, 5, 5]})


class _TwoStageNet:
    """What losses.aux_loss / losses_device.two_stage_aux_default read of the module in the `full` / `crf` classifier modes."""

    def __init__(self, c, seed):
        import types
        torch.manual_seed(seed)
        self.classifier_mode, self.num_tokens, self.loss_weights = "crf", c, None
        self.loss_cfg = {"aux": (-1, -1), "aux_sample_list": None}
        self.semantic_segmentation_head = types.SimpleNamespace()
        for i in range(c - 1):
            setattr(self.semantic_segmentation_head, f"ss_binary_classifier_{i}", types.SimpleNamespace(conv1=torch.nn.Conv2d(c, 1, 1)))


@pytest.mark.parametrize("seed,none_selected", [(1, False), (2, False), (3, True)])
def test_two_stage_aux_default_equals_host_form(seed, none_selected):
    """The fixed-shape two-stage auxiliary loss (no host sync, no boolean gathers: what the captured `crf` step runs) against
    the host restatement of model/semantic_segmentation_head.py:216-233: same value and same gradients, also when no pixel is
    selected for the second stage (the reference skips it)."""
    c, B, Hh, Ww = 5, 2, 12, 16
    net = _TwoStageNet(c, seed)
    g = torch.Generator().manual_seed(100 + seed)
    pm = torch.randn(B, 3, Hh, Ww, generator=g)
    if none_selected:
        pm[:, 1] = -10.0                                            # the predicted mask class is never 1
    pm.requires_grad_()
    ps = torch.randn(B, c, Hh, Ww, generator=g, requires_grad=True)
    pos_neg = torch.randint(0, 3, (B, Hh, Ww), generator=g)
    cls = torch.randint(0, c, (B, Hh, Ww), generator=g)
    params = [p for i in range(c - 1) for p in getattr(net.semantic_segmentation_head, f"ss_binary_classifier_{i}").conv1.parameters()]
    a = H.aux_loss(net, {"pred_mask": pm, "pred_ss": ps, "pos_neg_labels": pos_neg, "class_labels": cls})
    b = D.two_stage_aux_default(net, pm, ps, pos_neg, cls)
    assert a.shape == b.shape == (1,)
    assert abs(float(a) - float(b)) <= 2e-6 * max(1.0, abs(float(a)))
    ga = torch.autograd.grad(a.sum(), [pm, ps] + params, allow_unused=True)
    gb = torch.autograd.grad(b.sum(), [pm, ps] + params, allow_unused=True)
    for x, y in zip(ga, gb):
        if x is None or y is None:
            assert (x is None or float(x.abs().max()) == 0.0) and (y is None or float(y.abs().max()) == 0.0)
        else:
            assert float((x - y).abs().max()) <= 2e-6 * max(1.0, float(x.abs().max()))
