"""Host-side logic (no GPU): batch planning vs the oracle, and the forward
engine's orchestration with the kernels replaced by test-only stand-ins
(tests/mock_ops.py) against the reference fixtures."""
import numpy as np
import pytest
import torch

from conftest import build_case, load_golden
from test_oracle_golden import TINY, check_against_golden

import mock_ops
from oracle import oracle_ops
from vibertgrid_pytorch_b200 import engine as engine_mod
from vibertgrid_pytorch_b200 import plan as plan_mod


def test_resize_geometry_matches_oracle():
    for h, w, mn, mx in [(333, 777, 512, 800), (512, 512, 512, 800), (85, 107, 96, 128), (1000, 300, 704, 800), (40, 900, 320, 800)]:
        sc = oracle_ops.resize_scale(h, w, mn, mx)
        assert plan_mod.resize_geometry(h, w, mn, mx) == oracle_ops.resized_shape(h, w, sc)


@pytest.mark.parametrize("L,ntoks", [(512, [512, 300, 2]), (510, [510, 7]), (40, [40, 25]), (1024, [1024, 511, 1020]), (515, [515, 510, 3])])
def test_plan_windows_match_reference_windows(L, ntoks):
    B = len(ntoks)
    rng = np.random.default_rng(0)
    corpus = np.zeros((B, L), np.int64)
    for b, n in enumerate(ntoks):
        corpus[b, :n] = rng.integers(1000, 2000, n)
    mask = (corpus != 0).astype(np.int64)
    pl = plan_mod.plan_batch([(64, 64)] * B, ntoks, [1] * B, L, 64, 64)
    seq_tab, cu = pl.view("seq_tab").reshape(-1, 4), pl.view("cu")
    ids, pos = mock_ops.bert_assemble(torch.from_numpy(corpus), torch.from_numpy(seq_tab.copy()), torch.from_numpy(cu.copy()), pl.nseq, pl.R)
    wins = oracle_ops.bert_windows(corpus, mask)
    # every packed row must equal the reference window's (id, position) at a mask==1 slot, in order
    for q, (b, col0, n, sep) in enumerate(seq_tab):
        w_ids, w_mask, _ = wins[col0 // 510]
        keep = np.nonzero(w_mask[b])[0]
        assert np.array_equal(ids[cu[q]:cu[q + 1]].numpy(), w_ids[b][keep])
        assert np.array_equal(pos[cu[q]:cu[q + 1]].numpy(), keep)
    # token -> packed row map: row holds that token's id
    tok_row, tok_off = pl.view("tok_row"), pl.view("tok_off")
    for b, n in enumerate(ntoks):
        assert np.array_equal(ids[tok_row[tok_off[b]:tok_off[b + 1]]].numpy(), corpus[b, :n])
    # windows without real tokens are dropped
    assert pl.nseq == sum(1 for b in range(B) for w in range(L // 510 + 1) if min(max(ntoks[b] - 510 * w, 0), min(510, L - 510 * w)) > 0)


@pytest.mark.parametrize("name", TINY)
def test_engine_orchestration_with_standins(name, tmp_path, monkeypatch):
    fx = load_golden(name)
    monkeypatch.chdir(tmp_path)
    cfg, kw, net, batch = build_case(fx["meta"])
    net.eval()
    monkeypatch.setattr(engine_mod, "ops", mock_ops)
    monkeypatch.setattr(engine_mod, "make_epilogue", mock_ops.make_epilogue)
    eng = net._get_engine()
    eng._test_standins = True
    eng.fuse_aux_loss = False          # this test also checks the painted label maps
    out = eng.run(*batch, want_seg=True)
    assert int(out["status"]) == 0
    nchw = lambda t: t.permute(0, 3, 1, 2)
    so = out["plan"].view("seg_off")
    o = dict(image_batch=nchw(out["image_batch"]), coors_t=[out["boxes"][so[b]:so[b + 1]].numpy() for b in range(len(so) - 1)],
             index_map=out["index_map"].numpy(), seg_emb=[out["seg_emb"].numpy()], p_fuse=nchw(out["p_fuse"]),
             roi=nchw(out["roi"]), late=out["late"], pred_label=out["pred_label"], pred_mask=out["pred_mask"],
             pred_ss=out["pred_ss"], pos_neg_labels=out["pos_neg_labels"].numpy(),
             class_labels=out["class_labels"].numpy(), gt_label=out["gt_label"].numpy())
    if "logits" in fx:
        o["logits"] = out["logits"]
    check_against_golden(o, fx, {"default": 5e-5}, big=False)
    # the loss path (host-side mirror of the reference's loss modules)
    from vibertgrid_pytorch_b200 import losses
    total = losses.main_loss(net, out) + net.loss_control_lambda * losses.aux_loss(net, out)
    assert abs(float(total.reshape(-1)[0]) - float(fx["loss"][0])) <= 1e-4 * max(1.0, abs(float(fx["loss"][0])))
    # fused auxiliary-loss wiring (the product default): same total loss without materialised label maps
    eng.fuse_aux_loss = True
    out2 = eng.run(*batch, want_seg=True)
    if cfg.classifier_mode == "simp":
        assert "aux_ce" in out2 and "pos_neg_labels" not in out2
    total2 = losses.main_loss(net, out2) + net.loss_control_lambda * losses.aux_loss(net, out2)
    assert abs(float(total2.reshape(-1)[0]) - float(fx["loss"][0])) <= 1e-4 * max(1.0, abs(float(fx["loss"][0])))


def test_roberta_position_ids_match_transformers():
    """plan.roberta_position_ids over PACKED rows (sequences back to back, ``pos`` = index inside the sequence) against
    transformers' own RobertaEmbeddings.create_position_ids_from_input_ids applied per sequence -- including <pad> ids (1)
    inside and at the start of a sequence."""
    import torch
    from vibertgrid_pytorch_b200.plan import roberta_position_ids
    try:
        from transformers.models.roberta.modeling_roberta import RobertaEmbeddings
        hf = RobertaEmbeddings.create_position_ids_from_input_ids
    except Exception:                       # pragma: no cover - other transformers layouts
        pytest.skip("transformers without RobertaEmbeddings.create_position_ids_from_input_ids")
    g = torch.Generator().manual_seed(3)
    lens = [7, 1, 512, 4, 33]
    seqs = [torch.randint(0, 50, (n,), generator=g) for n in lens]       # small vocabulary: id 1 occurs often
    seqs[0][0] = 1
    seqs[2][5:9] = 1
    ids = torch.cat(seqs).to(torch.int32)
    pos = torch.cat([torch.arange(n) for n in lens]).to(torch.int32)
    got = roberta_position_ids(ids, pos, 1)
    want = torch.cat([hf(s[None].long(), 1)[0] for s in seqs])
    assert got.dtype == pos.dtype and torch.equal(got.long(), want)


@pytest.mark.parametrize("n_toks,L", [([15, 24], 24), ([3, 40, 17], 40), ([515, 30], 515), ([9], 30)])
def test_roberta_position_ids_ragged_packed_layout(n_toks, L):
    """The packed layout of a ragged batch (sample 0 shorter than the batch maximum; RoBERTa's <pad> id inside a
    later sample): ids / pos / cu exactly as ``bert_assemble_kernel`` writes them from ``plan_batch``'s window table -- the
    [SEP] row carries its slot in the PADDED window -- against transformers' rule applied to the padded windows the reference
    builds (model/BERTgrid_generator.py:106-129), read back at the rows the packed layout keeps."""
    import numpy as np
    import torch
    from vibertgrid_pytorch_b200.plan import plan_batch, roberta_position_ids
    try:
        from transformers.models.roberta.modeling_roberta import RobertaEmbeddings
        hf = RobertaEmbeddings.create_position_ids_from_input_ids
    except Exception:                       # pragma: no cover
        pytest.skip("transformers without RobertaEmbeddings.create_position_ids_from_input_ids")
    from oracle import oracle_ops
    B = len(n_toks)
    g = torch.Generator().manual_seed(7)
    corpus = torch.zeros(B, L, dtype=torch.int64)
    for b, n in enumerate(n_toks):
        corpus[b, :n] = torch.randint(1000, 2000, (n,), generator=g)
    corpus[B - 1, min(2, n_toks[-1] - 1)] = 1                     # <pad> id inside the LAST sample's text
    plan = plan_batch([(64, 64)] * B, n_toks, [1] * B, L, 64, 64)
    seq_tab, cu = plan.view("seq_tab").reshape(-1, 4), plan.view("cu")
    ids, pos, want = [], [], []
    windows = oracle_ops.bert_windows(corpus.numpy(), (corpus != 0).int().numpy())
    for (b, col0, n, sep) in seq_tab:
        ids += [101] + corpus[b, col0:col0 + n].tolist() + [102]
        pos += list(range(n + 1)) + [int(sep)]
        wid = windows[col0 // 510][0]                               # [B, 512] padded window ids
        full = hf(torch.from_numpy(wid[b:b + 1]).long(), 1)[0]
        want += full[:n + 1].tolist() + [int(full[sep])]
    got = roberta_position_ids(torch.tensor(ids, dtype=torch.int32), torch.tensor(pos, dtype=torch.int32), 1,
                               torch.from_numpy(np.ascontiguousarray(cu)))
    assert got.tolist() == want
    assert min(got.tolist()) >= 1


def test_missing_pretrained_weights_raise_outside_eval_mode(tmp_path, monkeypatch):
    """in work_mode 'train' / 'inference' the reference's from_pretrained / resnetXX(pretrained=True) raise when the
    weights are absent (model/ViBERTgrid_net.py:232-253, model/ResNetFPN_ViBERTgrid.py:521); the drop-in must not silently
    train from scratch.  A directory with a config but no weights stands in for an offline node."""
    import dataclasses
    from vibertgrid_pytorch_b200 import synth
    from vibertgrid_pytorch_b200.net import ViBERTgridNet
    monkeypatch.chdir(tmp_path)
    monkeypatch.setenv("HF_HUB_OFFLINE", "1")
    monkeypatch.setenv("TORCH_HOME", str(tmp_path / "no_hub"))
    monkeypatch.delenv("VBG_ALLOW_RANDOM_INIT", raising=False)
    cfg = synth.CONFIGS["tiny"]
    synth.write_bert_dir(cfg, str(tmp_path))
    with pytest.raises(Exception):
        ViBERTgridNet(**synth.model_kwargs(cfg, "train"))
    pre = dataclasses.replace(synth.CONFIGS["tiny_pre"])
    synth.write_bert_dir(pre, str(tmp_path))
    ViBERTgridNet(**synth.model_kwargs(pre, "eval"))                 # eval mode: a checkpoint is loaded right after
    monkeypatch.setenv("VBG_ALLOW_RANDOM_INIT", "1")
    with pytest.warns(UserWarning):
        net = ViBERTgridNet(**synth.model_kwargs(cfg, "train"))
    assert net.work_mode == "train"
