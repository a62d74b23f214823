"""One training step (loss, every parameter gradient, BatchNorm running statistics) of the drop-in module on the B200 against
the UNMODIFIED reference's ``model.train(); loss = model(...); loss.backward()`` on the same seeded weights and documents
(fixtures: oracle/make_train_golden.py, dropout zeroed on both sides)."""
import os
import random

import numpy as np
import pytest
import torch

from conftest import build_case, load_golden

pytestmark = pytest.mark.gpu

N_SAMPLES = 64


def summarize(t):
    f = t.detach().double().reshape(-1).cpu()
    idx = torch.linspace(0, f.numel() - 1, min(N_SAMPLES, f.numel())).long()
    return np.concatenate([[float(f.sum()), float(f.norm())], f[idx].numpy()])


def _to_dev(batch):
    img, seg, cls, coors, corpus, mask = batch
    c = lambda ts: tuple(t.cuda() for t in ts)
    return c(img), c(seg), c(cls), c(coors), corpus.cuda(), mask.cuda()


def run_step(name, tmp_path, monkeypatch, precision="bf16x3"):
    fx = load_golden(name)
    monkeypatch.setenv("VBG_PRECISION", precision)
    monkeypatch.chdir(tmp_path)
    cfg, kw, net, batch = build_case(fx["meta"])
    net = net.cuda()
    net.train()
    net.bert_hidden_dropout = net.bert_attn_dropout = 0.0
    net.loss_sampling = "host"                                 # the fixtures hold the reference's own Python-`random` draws
    random.seed(fx["meta"].get("py_random_seed", 0))          # the sampled losses draw from Python's `random` like the reference
    loss = net(*_to_dev(batch))
    loss.backward()
    torch.cuda.synchronize()
    assert int(net._train_engine.last["status"]) == 0
    return fx, net, loss


# Tolerances.  With random weights the network amplifies rounding noise ~1e4-fold into the gradients (tiny BatchNorm populations:
# 24 rows at the last stage of `tiny`, 4 rows in `tiny_pre`): the reference's OWN fp32 gradients deviate from the float64
# restatement (oracle/oracle_train.py) by 1e-3 (median) to 4e-3 of a tensor's largest element, 1.6e-2 for one BN bias, and the
# key-bias gradients are pure rounding noise around an exact zero.  So: with exact-fp32 contractions (VBG_PRECISION=fp32) every
# gradient is held to 1e-2 pointwise / 2e-3 in norm on every fixture; the bf16x3 tensor-core default (unit round-off 2^-17
# instead of 2^-24 in the products) to 6e-2 pointwise / 5e-3 in norm on the two better-conditioned fixtures and to 4e-2 in norm
# on the two degenerate ones.  A wiring error shows up as O(1) in the norm.
# The pointwise figure is the 90th percentile over a tensor's 64 sampled elements (a single ReLU / max-pool decision that flips
# under rounding moves one element of a small-population BatchNorm gradient by several per cent; the maximum is bounded too).
def tolerances(name, precision):
    """-> (q90 pointwise, max pointwise, norm)"""
    if precision == "fp32":
        return 1e-2, 0.15, 2e-3
    return (6e-2, None, 1e-2) if name in ("train_tiny", "train_mid") else (None, None, 4e-2)


@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
@pytest.mark.parametrize("name", ["train_tiny", "train_mid", "train_tiny_d", "train_tiny_pre"])
def test_training_step_matches_reference(name, precision, tmp_path, monkeypatch):
    fx, net, loss = run_step(name, tmp_path, monkeypatch, precision)
    assert loss.dim() == 0 and loss.dtype == torch.float32
    check_step(fx, net, loss, *tolerances(name, precision), 1e-3)


# The other heads and loss configurations (fixtures: the reference in `full` / `crf` mode, single- and multi-layer classifiers,
# and the `simp` head with index-sampled + class-weighted losses).  Gradient tolerances: fp32 as above with headroom for
# fixtures first seen here; bf16x3 in norm only.  Loss: 1e-3, except the two-stage heads under bf16x3 (1e-2): their second
# auxiliary stage runs on the pixels whose predicted mask class is 1 -- a per-cell argmax that a 1e-4 logit perturbation can
# flip, moving 16 of ~3000 pixels in or out of the loss.
HEAD_CASES = ["train_tiny_crf", "train_tiny_crf_multi", "train_tiny_full", "train_tiny_full_multi", "train_tiny_sampled"]


@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
@pytest.mark.parametrize("name", HEAD_CASES)
def test_training_step_other_heads_match_reference(name, precision, tmp_path, monkeypatch):
    fx, net, loss = run_step(name, tmp_path, monkeypatch, precision)
    assert list(loss.shape) == [int(v) for v in fx["loss_shape"]]
    two_stage = fx["meta"]["classifier_mode"] != "simp"
    if precision == "fp32":
        check_step(fx, net, loss, 2e-2, 0.3, 5e-3, 1e-3)
    else:
        check_step(fx, net, loss, None, None, 4e-2, 1e-2 if two_stage else 1e-3)


def check_step(fx, net, loss, tol_q90, tol_max, tol_norm, tol_loss):
    want = float(fx["loss"][0])
    got_loss = float(loss.reshape(-1)[0])
    assert abs(got_loss - want) <= tol_loss * max(1.0, abs(want)), (got_loss, want)
    params = dict(net.named_parameters())
    bad, worst = [], [0.0, 0.0, 0.0]
    for k in fx["grad_names"]:
        k = str(k)
        ref = fx["g:" + k]
        assert params[k].grad is not None, f"{k}: no gradient"
        got = summarize(params[k].grad)
        scale = max(np.abs(ref[2:]).max(), ref[1] / np.sqrt(params[k].numel()), 1e-12)
        errs = np.abs(got[2:] - ref[2:]) / scale
        err, q90 = errs.max(), np.quantile(errs, 0.9)
        nerr = abs(got[1] - ref[1]) / max(ref[1], 1e-12)
        if k.endswith("attention.self.key.bias"):      # exactly zero in exact arithmetic (softmax shift invariance)
            wref = fx["g:" + k.replace("key.bias", "key.weight")][1]
            assert got[1] <= 1e-3 * wref, (k, got[1], wref)
            continue
        worst = [max(worst[0], float(q90)), max(worst[1], float(err)), max(worst[2], float(nerr))]
        if (tol_q90 is not None and q90 > tol_q90) or (tol_max is not None and err > tol_max) or nerr > tol_norm:
            bad.append((k, float(err), float(nerr)))
    if os.environ.get("VBG_TEST_VERBOSE"):
        print(f"MARGIN {fx['meta']['name']} loss {got_loss:.6f} vs {want:.6f}; worst q90 {worst[0]:.2e} max {worst[1]:.2e} norm {worst[2]:.2e}")
    assert not bad, f"{len(bad)} gradients off: {bad[:12]}"
    for k in fx["no_grad_names"]:
        g = params[str(k)].grad
        assert g is None or float(g.abs().max()) == 0.0
    bufs = dict(net.named_buffers())
    for k in fx["buffer_names"]:
        k = str(k)
        ref, got = fx["b:" + k], summarize(bufs[k].float())
        assert np.abs(got - ref).max() <= 1e-4 * max(1.0, np.abs(ref).max()), k


def test_training_step_is_trainable(tmp_path, monkeypatch):
    """A few SGD steps on one batch reduce the loss (dropout on: the product default)."""
    fx = load_golden("train_tiny")
    monkeypatch.chdir(tmp_path)
    cfg, kw, net, batch = build_case(fx["meta"])
    net = net.cuda()
    net.train()
    opt = torch.optim.SGD(net.parameters(), lr=2e-3)
    dev = _to_dev(batch)
    losses = []
    torch.manual_seed(0)
    for _ in range(6):
        opt.zero_grad()
        loss = net(*dev)
        loss.backward()
        opt.step()
        losses.append(float(loss))
    assert all(np.isfinite(losses)) and losses[-1] < losses[0], losses
    net.eval()                                   # and the eval engine picks the updated weights up
    out = net(*dev)
    assert torch.isfinite(out[0]).all()
