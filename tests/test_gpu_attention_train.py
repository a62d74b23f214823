"""Training-mode self-attention on the tensor cores: vbg_attention_split_train_fwd (dropout on the attention probabilities
inside the kernel, base-2 row log-sum-exp stored) and vbg_attention_bwd_tc (tcgen05 bf16x3 products, probabilities rebuilt
from the log-sum-exp, the same counter-based dropout mask regenerated) against torch float64 autograd of

    O = dropout(softmax(Q K^T / 8)) V        per (sequence, head)     [HF BertSelfAttention, model/BERTgrid_generator.py:134]

with the EXACT keep mask (vbg_attention_dropout_mask exposes the hash the kernels use).  Bars: forward within 2e-5 and dQ / dK / dV within
3e-5 of float64 (max-rel, normalised by the tensor's abs-max); the mask's keep rate within 4 sigma of 1 - p; bitwise
reproducible launch to launch (no atomics)."""
import pytest
import torch

from conftest import relerr

pytestmark = pytest.mark.gpu

CASES = [([130, 64, 2, 200], 3), ([512, 2, 511], 2), ([1], 1), ([65, 127, 129, 256, 3], 12), ([512] * 2, 12)]


def _reference(qkv, d_o, cu, lens, heads, masks, inv_keep):
    hid = heads * 64
    qd = qkv.double().requires_grad_()
    outs = []
    for i, n in enumerate(lens):
        s = qd[int(cu[i]):int(cu[i + 1])].view(n, 3, heads, 64)
        q, k, v = (s[:, j].transpose(0, 1) for j in range(3))             # [heads, n, 64]
        p = torch.softmax(q @ k.transpose(1, 2) / 8.0, -1)
        if masks is not None:
            p = p * masks[i].double() * inv_keep
        outs.append((p @ v).transpose(0, 1).reshape(n, hid))
    o = torch.cat(outs)
    (ref,) = torch.autograd.grad(o, qd, d_o.double())
    return o.detach(), ref


@pytest.mark.parametrize("p_drop", [0.0, 0.1, 0.5])
@pytest.mark.parametrize("lens,heads", CASES)
def test_attention_train_forward_backward(lens, heads, p_drop):
    from vibertgrid_pytorch_b200 import ops
    assert ops.tc_available()
    if p_drop == 0.5 and sum(lens) > 700:
        pytest.skip("one dropout rate is enough at the large shapes")
    g = torch.Generator(device="cuda").manual_seed(sum(lens) + heads)
    hid, R = heads * 64, sum(lens)
    qkv = torch.randn(R, 3 * hid, device="cuda", generator=g)
    d_o = torch.randn(R, hid, device="cuda", generator=g)
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device="cuda")
    seed = 0x1234_5678_9ABC + R
    masks, inv_keep = None, 1.0
    if p_drop > 0:
        masks = []
        for i, n in enumerate(lens):
            per_head = [ops.attention_dropout_mask(seed, p_drop, int(cu[i]), n, h, qkv.device) for h in range(heads)]
            inv_keep = per_head[0][1]
            masks.append(torch.stack([m for m, _ in per_head]))                         # [heads, n, n]
        kept = torch.cat([m.reshape(-1) for m in masks])
        p_eff = 1.0 - 1.0 / inv_keep
        assert abs(p_eff - p_drop) < 1e-4
        if kept.numel() > 10_000:
            sigma = (p_eff * (1 - p_eff) / kept.numel()) ** 0.5
            assert abs(float(kept.mean()) - (1 - p_eff)) < 4 * sigma + 1e-4, "keep rate of the dropout hash"
    o_ref, ref = _reference(qkv, d_o, cu, lens, heads, masks, inv_keep)

    qs = ops.to_split(qkv)
    out, lse2 = ops.attention_split_train(qs, cu, len(lens), max(lens), heads, p_drop, seed)
    dqkv = ops.attention_bwd_tc(qs, out, d_o, lse2, cu, len(lens), max(lens), heads, p_drop, seed)
    torch.cuda.synchronize()
    e_o = relerr(out.cpu().numpy(), o_ref.cpu().numpy())
    # dQ and dK vanish identically for a one-row sequence: (exact cancellation dP - delta): each block is normalised by max(its own scale, 0.1 x the whole gradient's)
    floor = 0.1 * float(ref.abs().max())
    hidq = []
    for j in range(3):
        a, b = dqkv[:, j * hid:(j + 1) * hid].double().cpu(), ref[:, j * hid:(j + 1) * hid].cpu()
        hidq.append(float((a - b).abs().max() / max(float(b.abs().max()), floor, 1e-20)))
    print(f"[attention train lens={lens} heads={heads} p={p_drop}] O {e_o:.1e}  dQ {hidq[0]:.1e}  dK {hidq[1]:.1e}  dV {hidq[2]:.1e}")
    assert e_o < 2e-5 and max(hidq) < 3e-5        # bf16x3 products (unit round-off 2^-17) through the cancellation dP - delta
    # base-2 log-sum-exp of the UN-dropped scores
    for i, n in enumerate(lens[:2]):
        s = qkv[int(cu[i]):int(cu[i + 1])].double().view(n, 3, heads, 64)
        lse = torch.logsumexp(s[:, 0].transpose(0, 1) @ s[:, 1].transpose(0, 1).transpose(1, 2) / 8.0, -1) / 0.6931471805599453
        assert (lse2[int(cu[i]):int(cu[i + 1])].double().t() - lse).abs().max() < 1e-4
    # deterministic
    out2, lse2b = ops.attention_split_train(qs, cu, len(lens), max(lens), heads, p_drop, seed)
    assert torch.equal(out, out2) and torch.equal(lse2, lse2b)
    assert torch.equal(dqkv, ops.attention_bwd_tc(qs, out, d_o, lse2, cu, len(lens), max(lens), heads, p_drop, seed))
    if p_drop > 0:          # another seed gives another mask
        out3, _ = ops.attention_split_train(qs, cu, len(lens), max(lens), heads, p_drop, seed + 1)
        assert R < 4 or not torch.equal(out, out3)


def test_attention_autograd_function_uses_the_tensor_core_backward():
    from vibertgrid_pytorch_b200 import autograd as A, ops
    lens, heads = [100, 37], 2
    g = torch.Generator(device="cuda").manual_seed(5)
    R, hid = sum(lens), heads * 64
    qkv = torch.randn(R, 3 * hid, device="cuda", generator=g, requires_grad=True)
    d_o = torch.randn(R, hid, device="cuda", generator=g)
    cu = torch.tensor([0, 100, 137], dtype=torch.int32, device="cuda")
    out = A.AttentionF.apply(qkv, cu, 2, 100, heads, 0.0, 0)
    out.backward(d_o)
    _, ref = _reference(qkv.detach(), d_o, cu, lens, heads, None, 1.0)
    assert relerr(qkv.grad.cpu().numpy(), ref.cpu().numpy()) < 2e-5
    c0 = ops.L.launch_count
    A.AttentionF.apply(qkv, cu, 2, 100, heads, 0.1, 77).backward(d_o)
    assert ops.L.launch_count - c0 >= 4
