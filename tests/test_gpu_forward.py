"""End-to-end parity of the joint forward on the GPU: against the reference
fixtures (tests/golden, produced by the live reference) and against the oracle
at the full BASELINE shapes through size-independent properties."""
import os

import numpy as np
import pytest
import torch

from conftest import build_case, load_golden, relerr
from test_oracle_golden import TINY, check_against_golden

pytestmark = pytest.mark.gpu

# max-rel (normalised by the tensor's abs-max) tolerances; north_star: fp32 logits within 1e-3
TOL = {"fp32": {"default": 1e-4, "image": 1e-5, "seg_emb": 1e-4},
       # parity-grade tensor-core mode (3 bf16 tcgen05 products on hi/lo splits): the north_star bar
       "bf16x3": {"default": 1e-3, "image": 1e-5},
       }   # kind::tf32 (VBG_PRECISION=tf32) is a fast, non-parity mode (measured 2e-3..6e-3 on these fixtures; it flips an
           # argmax on cfg1), covered at kernel level only (tests/test_gpu_ops.py)
PREC = {"fp32": 0, "tf32": 1, "bf16x3": 2}


def _to_dev(batch):
    return [tuple(t.cuda() for t in x) if isinstance(x, tuple) else x.cuda() for x in batch]


def _collect(net, out):
    so = out["plan"].view("seg_off")
    nchw = lambda t: t.permute(0, 3, 1, 2).cpu()
    boxes = out["boxes"].cpu().numpy()
    o = dict(image_batch=nchw(out["image_batch"]), coors_t=[boxes[so[b]:so[b + 1]] for b in range(len(so) - 1)],
             index_map=out["index_map"].cpu().numpy(), seg_emb=[out["seg_emb"].cpu().numpy()], p_fuse=nchw(out["p_fuse"]),
             roi=nchw(out["roi"]), late=out["late"].cpu(), pred_label=out["pred_label"].cpu(),
             pred_mask=out["pred_mask"].cpu(), pred_ss=out["pred_ss"].cpu(), pos_neg_labels=out["pos_neg_labels"].cpu().numpy(),
             class_labels=out["class_labels"].cpu().numpy(), gt_label=out["gt_label"].cpu().numpy())
    if "logits" in out:
        o["logits"] = out["logits"].cpu()
    return o


@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
@pytest.mark.parametrize("name", TINY + ["cfg1", "cfg2_b1", "cfg4_b1", "cfg5_b1"])
def test_forward_matches_reference_fixture(name, precision, tmp_path, monkeypatch):
    from vibertgrid_pytorch_b200 import ops
    if precision != "fp32":
        assert ops.tc_available(), "tcgen05 path unavailable on this GPU box"
    fx = load_golden(name)
    monkeypatch.chdir(tmp_path)
    cfg, kw, net, batch = build_case(fx["meta"])
    net = net.cuda().eval()
    net._get_engine().precision = PREC[precision]
    net._get_engine().fuse_aux_loss = False          # this run also checks the painted label maps
    loss, pred_mask, pred_ss, gt, pred = net(*_to_dev(batch))
    out = net.last_intermediates
    assert int(out["status"].item()) == 0
    errs = check_against_golden(_collect(net, out), fx, TOL[precision], big=name.startswith("cfg"))
    print(f"[{name}/{precision}] " + ", ".join(f"{k}={v:.1e}" for k, v in errs.items()))
    assert pred.shape == tuple(fx["pred_label"].shape) and pred_mask.shape[1] == 3
    want = float(fx["loss"][0])
    tol_loss = {"fp32": 1e-4, "bf16x3": 1e-3, "tf32": 2e-2}[precision] * max(1.0, abs(want))
    assert abs(float(loss.detach().reshape(-1)[0]) - want) <= tol_loss
    # default product path: the auxiliary CE is fused with the label painting (labels never materialised)
    net._get_engine().fuse_aux_loss = True
    loss2 = net(*_to_dev(batch))[0]
    if fx["meta"]["classifier_mode"] == "simp":
        assert "aux_ce" in net.last_intermediates and "pos_neg_labels" not in net.last_intermediates
    assert abs(float(loss2.detach().reshape(-1)[0]) - want) <= tol_loss


def test_inference_entry_point_and_mode_quirk(tmp_path, monkeypatch):
    fx = load_golden("tiny_simp")
    monkeypatch.chdir(tmp_path)
    cfg, kw, net, batch = build_case(fx["meta"])
    net = net.cuda()
    net.eval()
    assert net.work_mode == "train" and not net.training        # eval() leaves work_mode == "train"
    img, seg, cls, coors, corpus, mask = _to_dev(batch)
    pred = net.inference(img, seg, coors, corpus, mask)
    assert relerr(pred.cpu().numpy(), fx["pred_label"]) < 1e-3
    net.train()                                                  # training mode returns the loss alone (ViBERTgrid_net.py:541)
    loss = net(img, seg, cls, coors, corpus, mask)
    assert isinstance(loss, torch.Tensor) and loss.dim() == 0 and loss.requires_grad


def test_mask_argument_is_checked_on_device(tmp_path, monkeypatch):
    """forward()'s ``mask`` (reference: selects the real rows, model/BERTgrid_generator.py:152-158, asserted against
    seg_indices :233) must be the prefix mask of the collate; a hole or a wrong count raises status bit 2 without a host
    sync -- eagerly and through the CUDA-graph replay path -- and a valid mask leaves the status clean."""
    fx = load_golden("tiny_simp")
    monkeypatch.chdir(tmp_path)
    cfg, kw, net, batch = build_case(fx["meta"])
    net = net.cuda().eval()
    img, seg, cls, coors, corpus, mask = _to_dev(batch)
    hole = mask.clone(); hole[0, 1] = 0
    extra = mask.clone(); extra[1, -1] = 1
    for call in range(3):                            # eager, capture, replay
        net(img, seg, cls, coors, corpus, mask)
        assert int(net.last_intermediates["status"].item()) == 0
        net(img, seg, cls, coors, corpus, hole)
        assert int(net.last_intermediates["status"].item()) & 2
        net(img, seg, cls, coors, corpus, extra)
        assert int(net.last_intermediates["status"].item()) & 2
    assert net._get_engine().graph_replays >= 3


def test_crf_inference_decodes_the_batch_as_one_sequence(tmp_path, monkeypatch):
    """``CRFFieldTypeClassification.inference`` (model/field_type_classification_head.py:655-668) runs Viterbi over
    all K rows of the batch as ONE sequence, unlike forward() (per document, :703-713).  ``net.inference`` replicates that:
    its tags equal the oracle's Viterbi over the concatenated emissions."""
    from oracle import oracle_ops
    fx = load_golden("tiny_crf")
    monkeypatch.chdir(tmp_path)
    cfg, kw, net, batch = build_case(fx["meta"])
    net = net.cuda().eval()
    img, seg, cls, coors, corpus, mask = _to_dev(batch)
    net(img, seg, cls, coors, corpus, mask)
    logits = net.last_intermediates["logits"].cpu().numpy()
    per_doc = net.last_intermediates["pred_label"].cpu().numpy()
    assert np.array_equal(per_doc, fx["pred_label"])
    trans = net.field_type_classification_head.crf_layer.transitions.detach().cpu().numpy()
    T = trans.shape[0]
    _, path = oracle_ops.crf_viterbi(logits, trans, T - 2, T - 1)
    for _ in range(3):                               # eager, capture, replay
        got = net.inference(img, seg, coors, corpus, mask).cpu().numpy()
        assert got.shape == (logits.shape[0], 1) and np.array_equal(got[:, 0], np.asarray(path, np.float32))


def test_full_size_properties_cfg2(tmp_path, monkeypatch):
    """BASELINE configs[1] shape (r34, B=8, 512x512, L=512, S=128) -- no oracle run at this size;
    instead: (1) documents are independent => a sample's outputs do not depend on its batch mates
    (the data-parallel sharding property), bit-exactly on the fp32 path; (2) scatter/index-map
    consistency; (3) probabilities are a distribution."""
    import dataclasses
    from vibertgrid_pytorch_b200 import ops, synth
    from vibertgrid_pytorch_b200.net import ViBERTgridNet
    monkeypatch.chdir(tmp_path)
    cfg = synth.CONFIGS["cfg2"]
    synth.write_bert_dir(cfg, str(tmp_path))
    net = ViBERTgridNet(**synth.model_kwargs(cfg, "eval"))
    synth.fill_state_dict_(net, 0)
    net = net.cuda().eval()
    net._get_engine().precision = ops.PREC_FP32
    batch = synth.make_batch(cfg, 3)
    img, seg, cls, coors, corpus, mask = _to_dev(batch)
    net(img, seg, cls, coors, corpus, mask)
    full = net.last_intermediates
    S = cfg.segments
    one = lambda x, b: (x[b],) if isinstance(x, tuple) else x[b:b + 1]
    for b in (0, 5):
        net(one(img, b), one(seg, b), one(cls, b), one(coors, b), one(corpus, b), one(mask, b))
        solo = net.last_intermediates
        assert torch.equal(solo["index_map"][0], full["index_map"][b])
        assert torch.equal(solo["bertgrid"][0], full["bertgrid"][b])
        assert torch.equal(solo["p_fuse"][0], full["p_fuse"][b])
        assert torch.equal(solo["logits"], full["logits"][b * S:(b + 1) * S])
    idx, grid = full["index_map"], full["bertgrid"]
    assert bool(((grid.abs().sum(-1) == 0) == (idx < 0)).all())
    k = full["seg_emb"].shape[0]
    assert int(idx.max()) < S and k == 8 * S
    p = full["pred_label"]
    assert torch.allclose(p.sum(1), torch.ones_like(p[:, 0]), atol=1e-5) and bool((p >= 0).all())
    assert int(full["status"].item()) == 0


@pytest.mark.parametrize("name", ["tiny_simp", "tiny_rob"])
def test_cuda_graph_replay_equals_eager(name, tmp_path, monkeypatch):
    """A batch signature seen twice is captured into a CUDA graph; replays with NEW data of the same shapes must equal the
    eager launches bit for bit, and returned tensors must not alias the graph's static buffers.  ``tiny_rob``: the RoBERTa
    position ids (integer torch ops between the kernels, plan.roberta_position_ids) are part of the captured graph."""
    import dataclasses
    from vibertgrid_pytorch_b200 import synth
    fx = load_golden(name)
    monkeypatch.chdir(tmp_path)
    cfg, kw, net, batch0 = build_case(fx["meta"])
    net = net.cuda().eval()
    eng = net._get_engine()
    cfg = dataclasses.replace(cfg, ragged=False)          # identical tensor shapes for every seed
    batches = [_to_dev(synth.make_batch(cfg, s)) for s in (5, 5, 77, 78)]
    sig = lambda b: (tuple(t.shape for t in b[0]), tuple(t.shape for t in b[1]), tuple(t.shape for t in b[3]), b[4].shape)
    same_shapes = all(sig(b) == sig(batches[0]) for b in batches)
    assert same_shapes
    eng.use_graphs = False
    eager = [[t.clone() for t in net(*b)[1:]] + [net.last_intermediates["logits"].clone()] for b in batches]
    eng.use_graphs = True
    kept = []
    for i, b in enumerate(batches):
        res = net(*b)
        kept.append(res)
        got = list(res[1:]) + [net.last_intermediates["logits"]]
        for g, e in zip(got, eager[i]):
            assert torch.equal(g, e), f"graph/eager mismatch at call {i}"
    if same_shapes:
        assert eng.graph_replays >= 3                  # call 0 eager, call 1 capture(+replay), calls 2, 3 replay
        assert not torch.equal(kept[2][4], kept[3][4]) or torch.equal(eager[2][3], eager[3][3])   # copies, not aliases
        for r, e in zip(kept[2][1:], eager[2]):
            assert torch.equal(r, e)                   # still intact after later replays


@pytest.mark.parametrize("name,batch", [("cfg4", 2), ("cfg5", 1), ("cfg2", 3)])
def test_full_size_tensor_core_vs_exact_fp32_path(name, batch, tmp_path, monkeypatch):
    """BASELINE configs[1], [3], [4] shapes (768^2 / L=1024 / S=1024 char boxes / pretrained layout; 1024^2 / CRF head): no CPU
    oracle run at these sizes -- instead the tensor-core mode (bf16x3, TMA implicit GEMM, tcgen05 attention) is checked against
    this library's own exact-fp32 CUDA-core path, which the fixtures pin to the reference at cfg1.  Integer outputs bit-equal,
    floats within the north_star 1e-3."""
    import dataclasses
    from vibertgrid_pytorch_b200 import ops, synth
    from vibertgrid_pytorch_b200.net import ViBERTgridNet
    assert ops.tc_available()
    monkeypatch.chdir(tmp_path)
    cfg = dataclasses.replace(synth.CONFIGS[name], batch=batch)
    synth.write_bert_dir(cfg, str(tmp_path))
    net = ViBERTgridNet(**synth.model_kwargs(cfg, "eval"))
    synth.fill_state_dict_(net, 1)
    net = net.cuda().eval()
    eng = net._get_engine()
    eng.use_graphs = False
    eng.fuse_aux_loss = False
    dev = _to_dev(synth.make_batch(cfg, 11))
    res = {}
    for prec in ("fp32", "bf16x3"):
        eng.precision = PREC[prec]
        eng.invalidate()
        net(*dev)
        o = net.last_intermediates
        assert int(o["status"].item()) == 0
        res[prec] = {k: o[k].clone() for k in ("index_map", "boxes", "seg_emb", "p_fuse", "roi", "late", "logits", "pred_label",
                                               "pred_mask", "pred_ss", "pos_neg_labels", "class_labels")}
    a, b = res["fp32"], res["bf16x3"]
    for k in ("index_map", "boxes", "pos_neg_labels", "class_labels"):
        assert torch.equal(a[k], b[k]), k
    errs = {k: relerr(b[k].cpu().numpy(), a[k].cpu().numpy()) for k in ("seg_emb", "p_fuse", "roi", "late", "logits", "pred_mask", "pred_ss")}
    print(f"[{name} x{batch} bf16x3 vs fp32] " + ", ".join(f"{k}={v:.1e}" for k, v in errs.items()))
    assert max(errs.values()) < 1e-3, errs
    if cfg.classifier_mode == "crf":
        assert torch.equal(a["pred_label"], b["pred_label"])          # identical Viterbi paths
    else:
        assert torch.equal(a["pred_label"].argmax(1), b["pred_label"].argmax(1))


def test_two_stream_forward_equals_single_stream_and_prefetcher(tmp_path, monkeypatch):
    """The fork/join of the forward over two streams (BERT || early backbone, seg head || ROI path) only reorders independent
    kernels: outputs are bit-identical to the single-stream order, eagerly and under graph replay.  DevicePrefetcher yields
    the same batches in order."""
    from vibertgrid_pytorch_b200.prefetch import DevicePrefetcher
    fx = load_golden("tiny_simp")
    monkeypatch.chdir(tmp_path)
    cfg, kw, net, batch = build_case(fx["meta"])
    net = net.cuda().eval()
    eng = net._get_engine()
    ref = None
    for multi in (False, True):
        eng.multi_stream = multi
        eng.invalidate()
        outs = []
        for dev_batch in DevicePrefetcher([batch] * 8, torch.device("cuda")):     # eager, capture, replays; ring slots reused
            res = net(*dev_batch)
            outs.append([t.clone() for t in res] + [net.last_intermediates["logits"].clone(), net.last_intermediates["p_fuse"].clone()])
        torch.cuda.synchronize()
        for o in outs[1:]:
            for a, b in zip(o, outs[0]):
                assert torch.equal(a, b)
        if ref is None:
            ref = outs[0]
        else:
            for a, b in zip(outs[0], ref):
                assert torch.equal(a, b)
