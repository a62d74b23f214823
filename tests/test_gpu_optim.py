"""Fused multi-tensor optimizers (vibertgrid_pytorch_b200.optim, vbg_sgd_step_mt / vbg_adamw_step_mt) against
torch.optim.SGD / torch.optim.AdamW -- the optimizers the reference builds (train_SROIE.py:217-235) -- over several steps on
tensors of awkward sizes (unaligned storage offsets, one element, more than one chunk), including a parameter whose gradient
appears late, a learning-rate / weight-decay change between steps (the reference's per-iteration schedulers), and state-dict
interchange with torch's optimizers."""
import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPES = [(1,), (7,), (1023,), (64, 3, 7, 7), (8192,), (8193,), (300, 257), (3, 50001)]


def _params(seed, offset):
    g = torch.Generator().manual_seed(seed)
    out = []
    for s in SHAPES:
        n = int(torch.tensor(s).prod())
        base = torch.randn(n + offset, generator=g).cuda()
        out.append(base[offset:].view(s).detach().requires_grad_())       # offset 1: 4-byte aligned storage -> scalar path
    return out


def _grads(params, step):
    g = torch.Generator().manual_seed(1000 + step)
    return [torch.randn(p.shape, generator=g).cuda() for p in params]


@pytest.mark.parametrize("offset", [0, 1])
@pytest.mark.parametrize("kind", ["sgd", "sgd_nomom", "adamw"])
def test_fused_optimizer_matches_torch(kind, offset):
    from vibertgrid_pytorch_b200.optim import FusedAdamW, FusedSGD
    ours, ref = _params(3, offset), _params(3, offset)
    if kind == "adamw":
        o1 = FusedAdamW(ours, lr=5e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01)
        o2 = torch.optim.AdamW(ref, lr=5e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01)
    else:
        mom = 0.9 if kind == "sgd" else 0.0
        o1 = FusedSGD(ours, lr=5e-2, momentum=mom, weight_decay=5e-3)
        o2 = torch.optim.SGD(ref, lr=5e-2, momentum=mom, weight_decay=5e-3)
    for step in range(5):
        gs = _grads(ours, step)
        for i, (a, b, g) in enumerate(zip(ours, ref, gs)):
            late = i == 2 and step < 2                # this parameter gets its first gradient at step 2
            a.grad = None if late else g.clone()
            b.grad = None if late else g.clone()
        if step == 3:                                 # the reference's schedulers rewrite these every iteration
            for grp in o1.param_groups + o2.param_groups:
                grp["lr"] *= 0.5
                grp["weight_decay"] *= 2.0
        o1.step(); o2.step()
    torch.cuda.synchronize()
    for a, b in zip(ours, ref):
        assert float((a - b).abs().max()) <= 2e-6 * max(1.0, float(b.abs().max())), a.shape
    # state names / values interchange with torch's optimizers
    s1, s2 = o1.state_dict()["state"], o2.state_dict()["state"]
    for k in s2:
        for name in ("momentum_buffer", "exp_avg", "exp_avg_sq"):
            if name in s2[k] and s2[k][name] is not None:
                assert float((s1[k][name] - s2[k][name]).abs().max()) <= 2e-6 * max(1.0, float(s2[k][name].abs().max()))
    if kind == "adamw":
        o3 = torch.optim.AdamW(ref, lr=1e-3)
        sd = o1.state_dict()
        for st in sd["state"].values():
            st["step"] = torch.tensor(float(st["step"]))
        sd["param_groups"] = [{k: v for k, v in g.items() if k != "grad_scale"} | {kk: vv for kk, vv in o3.state_dict()["param_groups"][0].items()
                                                                                    if kk not in g} for g in sd["param_groups"]]
        o3.load_state_dict(sd)


def test_fused_optimizers_step_the_model(tmp_path, monkeypatch):
    """One training step of the real module with the fused optimizers, split like the reference (names with 'bert_model' ->
    AdamW, the rest -> SGD), equals the same step with torch's optimizers."""
    import dataclasses
    from conftest import build_case, load_golden
    from vibertgrid_pytorch_b200.optim import FusedAdamW, FusedSGD
    fx = load_golden("train_tiny")
    monkeypatch.chdir(tmp_path)
    cfg, kw, net, batch = build_case(fx["meta"])
    net = net.cuda().train()
    net.bert_hidden_dropout = net.bert_attn_dropout = 0.0
    dev = [tuple(t.cuda() for t in x) if isinstance(x, tuple) else x.cuda() for x in batch]
    sd0 = {k: v.clone() for k, v in net.state_dict().items()}
    res = []
    for fused in (False, False, True):          # torch twice: the training step is deterministic, so the two must agree bit for bit
        net.load_state_dict(sd0)
        bert = [p for n, p in net.named_parameters() if "bert_model" in n]
        cnn = [p for n, p in net.named_parameters() if "bert_model" not in n]
        o_c = (FusedSGD if fused else torch.optim.SGD)(cnn, lr=1e-2, momentum=0.9, weight_decay=5e-3)
        o_b = (FusedAdamW if fused else torch.optim.AdamW)(bert, lr=1e-4, weight_decay=1e-2)
        net._train_engine = None
        from vibertgrid_pytorch_b200.train_engine import TrainEngine
        net._train_engine = TrainEngine(net)
        net._train_engine.use_graphs = False
        snaps = []
        for _ in range(2):
            o_c.zero_grad(); o_b.zero_grad()
            net(*dev).backward()
            o_c.step(); o_b.step()
            snaps.append({k: v.clone() for k, v in net.state_dict().items() if v.dtype == torch.float32})
        res.append(snaps)

    def dev(a, b, keys):
        return max(float((a[k] - b[k]).abs().max()) / max(1e-3, float(a[k].abs().max())) for k in keys)
    keys = [k for k in res[0][0] if not k.endswith("attention.self.key.bias")]
    repeat = dev(res[0][1], res[1][1], keys)
    one, two = dev(res[0][0], res[2][0], keys), dev(res[0][1], res[2][1], keys)
    print(f"parameters, torch vs torch after 2 steps {repeat:.2e}; torch vs fused after 1 step {one:.2e}, after 2 steps {two:.2e}")
    assert repeat == 0.0, "the training step is deterministic: two identical runs must agree exactly"
    # step 1 starts from identical gradients: only the optimizers' own rounding differs (AdamW's first update is lr * g / (|g| + eps),
    # ill-conditioned where |g| ~ eps, hence the 1e-3 floor in the denominator of `dev`)
    assert one <= 2e-5
    # step 2 sees gradients of parameters that differ by ~1e-7, amplified by this fixture's tiny-batch BatchNorms
    assert two <= 5e-3
