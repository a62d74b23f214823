"""The whole-step CUDA graph of the training path (train_engine.TrainEngine._graphed_loss): a batch signature seen for the
second time is captured -- train-mode forward + backward to every parameter gradient -- and replayed from then on.  Replays must
give the SAME loss and gradients as the eager tape, bit for bit (the step has no atomics), keep BatchNorm's running
statistics moving, scale with grad_output (GradScaler), follow parameter updates made by an optimizer between steps, and draw
a new dropout mask every replay (the device step seed)."""
import dataclasses

import pytest
import torch

from conftest import build_case, load_golden

pytestmark = pytest.mark.gpu


def _to_dev(batch):
    return [tuple(t.cuda() for t in x) if isinstance(x, tuple) else x.cuda() for x in batch]


def _net(tmp_path, monkeypatch, dropout):
    from vibertgrid_pytorch_b200 import synth
    fx = load_golden("train_mid")
    monkeypatch.chdir(tmp_path)
    cfg, kw, net, batch = build_case(fx["meta"])
    cfg = dataclasses.replace(cfg, ragged=False)
    net = net.cuda().train()
    if not dropout:
        net.bert_hidden_dropout = net.bert_attn_dropout = 0.0
    batches = [_to_dev(synth.make_batch(cfg, s)) for s in (3, 4, 5)]
    return net, batches


def _grads(net):
    return {k: p.grad.clone() for k, p in net.named_parameters() if p.grad is not None}


def test_graphed_step_equals_eager_step(tmp_path, monkeypatch):
    net, batches = _net(tmp_path, monkeypatch, dropout=False)
    eng_cls = type(net)
    sd0 = {k: v.clone() for k, v in net.state_dict().items()}

    def run(use_graphs):
        net.load_state_dict(sd0)
        net._train_engine = None
        out = []
        for i, b in enumerate([batches[0], batches[0], batches[1], batches[2]]):
            net.zero_grad(set_to_none=True)
            loss = net(*b)
            net._train_engine.use_graphs = use_graphs
            (loss * (3.0 if i == 3 else 1.0)).backward()          # a scaled backward (GradScaler) on the last step
            out.append((loss.detach().clone(), _grads(net), {k: v.clone() for k, v in net.state_dict().items() if "running" in k or "num_batches" in k}))
        return out, net._train_engine.graph_replays + net._train_engine.capture_failures

    net._train_engine = None
    eager, r0 = run(False)
    graphed, r1 = run(True)
    # step 0 eager, step 1 capture + replay, steps 2, 3 replay with NEW data (a capture invalidated from outside the step is
    # retried at the next sighting and counted in capture_failures: the bit-equality below holds on either path)
    assert r0 == 0 and r1 >= 3
    for step, ((le, ge, be), (lg, gg, bg)) in enumerate(zip(eager, graphed)):
        assert torch.equal(le, lg), f"loss differs at step {step}"
        assert ge.keys() == gg.keys(), (sorted(set(ge) ^ set(gg))[:8], len(ge), len(gg))
        for k in ge:
            # the same kernels in the same order on the same data, and no atomics anywhere in the step: bit for bit
            if step < 3:
                assert torch.equal(ge[k], gg[k]), f"gradient of {k} differs at step {step}"
            elif not k.endswith("attention.self.key.bias"):
                # the scaled step: the eager tape propagates 3 * dL, the graphed step multiplies dL's gradients by 3 afterwards
                assert float((ge[k] - gg[k]).abs().max()) <= 1e-4 * float(ge[k].abs().max()) + 1e-9, f"gradient of {k} differs at step {step}"
        for k in be:
            assert torch.equal(be[k], bg[k]), f"{k} differs at step {step}"
    # the last step was scaled by 3
    k = "bert_model.encoder.layer.0.attention.self.query.weight"
    assert float(eager[3][1][k].abs().max()) > 0


def test_graphed_step_follows_optimizer_updates_and_redraws_dropout(tmp_path, monkeypatch):
    net, batches = _net(tmp_path, monkeypatch, dropout=True)
    opt = torch.optim.SGD([p for p in net.parameters() if p.requires_grad], lr=1e-2)
    losses = []
    for i in range(6):
        opt.zero_grad()
        loss = net(*batches[0])
        loss.backward()
        if i >= 3:
            opt.step()
        losses.append(float(loss))
    assert net._train_engine.graph_replays + net._train_engine.capture_failures >= 4
    # steps 1, 2 replay the same graph on the same data and weights: they differ only through the dropout masks
    assert losses[1] != losses[2], "the dropout mask did not change between replays"
    assert abs(losses[1] - losses[2]) < 0.5 * abs(losses[1])
    # steps 3.. follow the in-place parameter updates
    assert losses[5] < losses[3]
    assert all(torch.isfinite(torch.tensor(losses)))


def test_example_config_losses_are_drawn_on_the_device_and_captured(tmp_path, monkeypatch):
    """The reference's example_config.yaml loss settings (index-sampled auxiliary loss, OHEM with random pre-sampling on the
    heads) with device-side sampling (losses_device.py): no host draws, no syncs -> the step is captured like the plain one;
    every replay draws a new subset; `loss_sampling = "host"` keeps the reference's Python-`random` path (eager)."""
    import dataclasses
    from vibertgrid_pytorch_b200 import synth
    from vibertgrid_pytorch_b200.net import ViBERTgridNet
    monkeypatch.chdir(tmp_path)
    cfg = dataclasses.replace(synth.CONFIGS["mid"], ragged=False)
    synth.write_bert_dir(cfg, str(tmp_path))
    kw = {**synth.model_kwargs(cfg, "eval"), "loss_aux_sample_list": [256, 512, 256], "num_hard_positive_aux": 256, "num_hard_negative_aux": 256,
          "num_hard_positive_main_1": 4, "num_hard_negative_main_1": 4, "num_hard_positive_main_2": 8, "num_hard_negative_main_2": 8,
          "ohem_random": True}
    net = ViBERTgridNet(**kw)
    synth.fill_state_dict_(net, 2)
    net = net.cuda().train()
    net.bert_hidden_dropout = net.bert_attn_dropout = 0.0
    batch = _to_dev(synth.make_batch(cfg, 9))
    losses = []
    for _ in range(5):
        net.zero_grad(set_to_none=True)
        loss = net(*batch)
        loss.backward()
        losses.append(float(loss))
    eng = net._train_engine
    assert eng._sampled_losses() and eng._device_sampling(batch[4].device)
    assert eng.graph_replays + eng.capture_failures >= 4 and eng.graph_replays >= 2     # sampled losses no longer force the eager path
    assert all(torch.isfinite(torch.tensor(losses)))
    assert len({round(l, 7) for l in losses[1:]}) >= 3  # same data, same weights: the loss moves only through the new draws
    assert max(losses) - min(losses) < 0.5 * abs(losses[0])
    g = [p.grad for p in net.parameters() if p.grad is not None]
    assert len(g) > 100 and all(bool(torch.isfinite(t).all()) for t in g)
    net.loss_sampling = "host"
    net._train_engine = None
    for _ in range(3):
        net.zero_grad(set_to_none=True)
        net(*batch).backward()
    assert net._train_engine.graph_replays == 0


def test_crf_step_is_captured_and_equals_eager(tmp_path, monkeypatch):
    """BASELINE configs[4]'s head: the `crf` classifier mode (two-stage auxiliary head with the plain-mean loss, CRF negative
    log-likelihood kernel) has no host sync left in its step (losses_device.two_stage_aux_default) and is captured like the
    `simp` step; replays give the eager tape's loss and gradients bit for bit."""
    from vibertgrid_pytorch_b200 import synth
    fx = load_golden("train_tiny_crf")
    monkeypatch.chdir(tmp_path)
    cfg, kw, net, batch = build_case(fx["meta"])
    cfg = dataclasses.replace(cfg, ragged=False)
    net = net.cuda().train()
    net.bert_hidden_dropout = net.bert_attn_dropout = 0.0
    batches = [_to_dev(synth.make_batch(cfg, s)) for s in (3, 4)]
    sd0 = {k: v.clone() for k, v in net.state_dict().items()}

    def run(use_graphs):
        net.load_state_dict(sd0)
        net._train_engine = None
        out = []
        for b in (batches[0], batches[0], batches[1], batches[0]):
            net.zero_grad(set_to_none=True)
            loss = net(*b)
            net._train_engine.use_graphs = use_graphs
            loss.backward()
            out.append((loss.detach().clone(), _grads(net)))
        return out, net._train_engine.graph_replays + net._train_engine.capture_failures

    eager, r0 = run(False)
    graphed, r1 = run(True)
    assert r0 == 0 and r1 >= 3, (r0, r1)
    for step, ((le, ge), (lg, gg)) in enumerate(zip(eager, graphed)):
        assert torch.equal(le, lg), f"loss differs at step {step}: {float(le):.7f} vs {float(lg):.7f}"
        assert ge.keys() == gg.keys()
        for k in ge:
            assert torch.equal(ge[k], gg[k]), f"gradient of {k} differs at step {step}"
    assert "field_type_classification_head.crf_layer.transitions" in eager[0][1]


def test_two_signatures_share_one_live_seed_word(tmp_path, monkeypatch):
    """Every captured graph has the ADDRESS of the device seed word baked in: meeting a second batch signature (an eager first
    sighting, then its own capture) must not drop or move that word -- replays of the FIRST signature keep drawing new dropout
    masks afterwards."""
    from vibertgrid_pytorch_b200 import synth
    fx = load_golden("train_mid")
    monkeypatch.chdir(tmp_path)
    cfg, kw, net, batch = build_case(fx["meta"])
    net = net.cuda().train()                                    # dropout on
    a = _to_dev(synth.make_batch(dataclasses.replace(cfg, ragged=False), 3))
    b = _to_dev(synth.make_batch(dataclasses.replace(cfg, ragged=False, segments=cfg.segments - 2), 4))
    losses = {"a": [], "b": []}
    for name, bt in [("a", a), ("a", a), ("a", a), ("b", b), ("b", b), ("b", b), ("a", a), ("a", a), ("a", a)]:
        net.zero_grad(set_to_none=True)
        loss = net(*bt)
        loss.backward()
        losses[name].append(float(loss))
    eng = net._train_engine
    assert eng.graph_replays + eng.capture_failures >= 7
    word = eng._step_seed.data_ptr()
    assert len({round(l, 7) for l in losses["a"][-3:]}) == 3, losses["a"]       # same data, same weights: only the masks move
    assert len({round(l, 7) for l in losses["b"][-2:]}) == 2, losses["b"]
    net.zero_grad(set_to_none=True)
    net(*a).backward()
    assert eng._step_seed.data_ptr() == word
