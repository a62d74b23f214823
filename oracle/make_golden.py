"""Generate ``tests/golden/*.npz`` by running the UNMODIFIED reference
(/root/reference, CPU, fp32, eval mode) on seeded synthetic documents and the
seeded weights of ``vibertgrid_pytorch_b200.synth`` -- and check the oracle
restatement (oracle_net / oracle_ops) against it on the spot.

Run in the build container only (the GPU box has no /root/reference):

    python oracle/make_golden.py            # all fixtures
    python oracle/make_golden.py tiny cfg1  # a subset

Fixtures store seeds + the reference's outputs (sub-sampled where large); the
weights and inputs are regenerated from the seeds at test time.
"""
from __future__ import annotations

import json
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, ROOT)

from vibertgrid_pytorch_b200 import synth  # noqa: E402
from vibertgrid_pytorch_b200.net import ViBERTgridNet as OurNet  # noqa: E402  (parameter tree only)
from oracle import oracle_net, oracle_ops  # noqa: E402

GOLDEN = {
    # fixture name -> (config name, classifier override, weight seed, input seed)
    "tiny_simp": ("tiny", "simp", 0, 0),
    "tiny_full": ("tiny", "full", 1, 1),
    "tiny_crf": ("tiny", "crf", 2, 2),
    "tiny_d": ("tiny_d", "simp", 3, 3),
    "tiny_pre": ("tiny_pre", "simp", 4, 4),
    "tiny_win": ("tiny_win", "simp", 5, 5),
    "cfg1": ("cfg1", "simp", 0, 0),
    "cfg2_b1": ("cfg2_b1", "simp", 1, 11),       # one document of BASELINE configs[1] (r34 + bert-base, 512^2, L=512, S=128)
    "cfg4_b1": ("cfg4_b1", "simp", 1, 11),       # configs[3]: 768^2, L=1024 (3 windows), 1024 char boxes, pretrained-layout r34
    "cfg5_b1": ("cfg5_b1", "crf", 1, 11),        # configs[4]: 1024^2, CRF head
    "tiny_rob": ("tiny_rob", "simp", 6, 6),      # bert_model="roberta-base": RobertaModel position ids, LayerNorm eps 1e-5
}


write_bert_dir = synth.write_bert_dir


def seed_hub(root):
    """Pre-seed the torch-hub cache so ``resnetXX(pretrained=True)`` needs no network."""
    import torchvision
    os.environ["TORCH_HOME"] = root
    ck = os.path.join(root, "hub", "checkpoints")
    os.makedirs(ck, exist_ok=True)
    torch.save(torchvision.models.resnet18().state_dict(), os.path.join(ck, "resnet18-f37072fd.pth"))
    torch.save(torchvision.models.resnet34().state_dict(), os.path.join(ck, "resnet34-b627a593.pth"))


def sub(t, *strides):
    """Strided sub-sample of the trailing dims (keeps fixtures small)."""
    sl = [slice(None)] * (t.dim() - len(strides)) + [slice(None, None, s) for s in strides]
    return t[tuple(sl)].contiguous().numpy()


def relerr(a, b):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def run_one(name, cfg_name, mode, wseed, iseed, outdir):
    import dataclasses
    cfg = dataclasses.replace(synth.CONFIGS[cfg_name], classifier_mode=mode)
    if mode == "crf" and cfg.tag_to_idx is None:
        cfg.tag_to_idx = {f"T{i}": i for i in range(cfg.num_classes)}
    with tempfile.TemporaryDirectory() as tmp:
        cwd = os.getcwd()
        os.chdir(tmp)
        try:
            write_bert_dir(cfg, tmp)
            seed_hub(tmp)
            kw = synth.model_kwargs(cfg, "eval")
            ours = OurNet(**kw)
            synth.fill_state_dict_(ours, wseed)
            sd = {k: v.clone() for k, v in ours.state_dict().items()}

            sys.path.insert(0, REF)
            for m in [m for m in sys.modules if m.split(".")[0] in ("model", "pipeline")]:
                del sys.modules[m]
            from model.ViBERTgrid_net import ViBERTgridNet as RefNet
            assert RefNet.__module__ == "model.ViBERTgrid_net" and REF in sys.modules["model.ViBERTgrid_net"].__file__
            kw = synth.model_kwargs(cfg, "eval")
            ref = RefNet(**kw)
            ref_keys, our_keys = set(ref.state_dict()), set(sd)
            assert ref_keys == our_keys, (sorted(ref_keys - our_keys)[:5], sorted(our_keys - ref_keys)[:5])
            for k, v in ref.state_dict().items():
                assert tuple(v.shape) == tuple(sd[k].shape), (k, v.shape, sd[k].shape)
            ref.load_state_dict(sd, strict=True)          # state-dict layout parity
            ref.eval()
            sys.path.remove(REF)

            batch = synth.make_batch(cfg, iseed)
            cap = {}
            hooks = [
                ref.transform.register_forward_hook(lambda m, i, o: cap.__setitem__("transform", o)),
                ref.BERTgrid_generator.register_forward_hook(lambda m, i, o: cap.__setitem__("bertgrid", o)),
                ref.backbone.register_forward_hook(lambda m, i, o: cap.__setitem__("p_fuse", o)),
                ref.grid_roi_align_net.register_forward_hook(lambda m, i, o: cap.__setitem__("roi", o)),
                ref.late_fusion_net.register_forward_hook(lambda m, i, o: cap.__setitem__("late", o)),
            ]
            head = ref.field_type_classification_head
            if mode != "full":
                hooks.append(head.category_classification_net.register_forward_hook(
                    lambda m, i, o: cap.__setitem__("logits", o)))
            seg = ref.semantic_segmentation_head
            hooks.append(seg.aux_loss_1.register_forward_pre_hook(
                lambda m, i: cap.__setitem__("pos_neg_labels", i[1].clone())))
            if mode == "simp":
                hooks.append(seg.aux_loss_2.register_forward_pre_hook(
                    lambda m, i: cap.__setitem__("class_labels", i[1].clone())))
            with torch.no_grad():
                loss, pred_mask, pred_ss, gt_label, pred_label = ref(*batch)
            for h in hooks:
                h.remove()
        finally:
            os.chdir(cwd)

    # ---------------- oracle vs reference, on the spot
    ocfg = oracle_net.OracleConfig(backbone=cfg.backbone, classifier_mode=mode, num_classes=cfg.num_classes,
                                   min_size=kw["test_image_min_size"], max_size=kw["image_max_size"], **({"ln_eps": 1e-5, "roberta_pad": 1} if "roberta-" in cfg.bert_name else {}))
    o = oracle_net.forward(sd, ocfg, *batch)
    image_list, coors_t = cap["transform"]
    seg_emb_ref, grid_ref = cap["bertgrid"]
    report = {}
    assert tuple(o["image_batch"].shape) == tuple(image_list.tensors.shape)
    report["image_batch"] = relerr(o["image_batch"], image_list.tensors)
    for a, b in zip(o["coors_t"], coors_t):
        assert b.dtype == torch.int32 and np.array_equal(a, b.numpy()), "transformed coords differ"
    report["seg_emb"] = max(relerr(a, b) for a, b in zip(o["seg_emb"], seg_emb_ref))
    # index map derived from the reference's BERTgrid by exact row matching (last match wins)
    B, C, Hg, Wg = grid_ref.shape
    ref_idx = np.full((B, Hg, Wg), -1, np.int32)
    for b in range(B):
        cells = grid_ref[b].permute(1, 2, 0).reshape(-1, C)
        for s in range(seg_emb_ref[b].shape[0]):
            hit = (cells == seg_emb_ref[b][s][None]).all(1).numpy().reshape(Hg, Wg)
            ref_idx[b][hit] = s
        zero = (cells == 0).all(1).numpy().reshape(Hg, Wg)
        assert ((ref_idx[b] >= 0) | zero).all()
    assert np.array_equal(ref_idx, o["index_map"]), "index map differs from the reference's BERTgrid"
    report["bertgrid"] = relerr(o["bertgrid"], grid_ref)
    report["p_fuse"] = relerr(o["p_fuse"], cap["p_fuse"])
    report["roi"] = relerr(o["roi"], cap["roi"])
    report["late"] = relerr(o["late"], cap["late"])
    if "logits" in cap:
        report["logits"] = relerr(o["logits"], cap["logits"])
    report["pred_label"] = relerr(o["pred_label"], pred_label)
    report["pred_mask"] = relerr(o["pred_mask"], pred_mask)
    report["pred_ss"] = relerr(o["pred_ss"], pred_ss)
    assert np.array_equal(o["pos_neg_labels"], cap["pos_neg_labels"].numpy())
    if "class_labels" in cap:
        assert np.array_equal(o["class_labels"], cap["class_labels"].numpy())
    assert torch.equal(o["gt_label"], gt_label)
    bad = {k: v for k, v in report.items() if v > 2e-5}
    print(f"[{name}] oracle vs reference max-rel: " + ", ".join(f"{k}={v:.2e}" for k, v in report.items()))
    assert not bad, f"oracle deviates from the reference: {bad}"

    big = cfg_name.startswith("cfg")
    s2 = (4, 4) if big else (2, 2)
    fx = dict(
        meta=json.dumps(dict(name=name, cfg=cfg_name, classifier_mode=mode, weight_seed=wseed, input_seed=iseed,
                             tag_to_idx=cfg.tag_to_idx, torch=torch.__version__, oracle_vs_ref=report)),
        image_shape=np.asarray(image_list.tensors.shape),
        image_sub=sub(image_list.tensors, 8, 8) if big else sub(image_list.tensors, 2, 2),
        coors_t=np.concatenate([c.numpy() for c in coors_t], 0),
        seg_emb=np.concatenate([e.numpy() for e in seg_emb_ref], 0)[:, ::(8 if big else 2)],
        index_map=ref_idx,
        p_fuse_sub=sub(cap["p_fuse"], *s2)[:, ::4],
        roi_stride=np.asarray(32 if cap["roi"].shape[0] >= 512 else (8 if big else 4)),
        roi_sub=cap["roi"][:, ::(32 if cap["roi"].shape[0] >= 512 else (8 if big else 4))].contiguous().numpy(),
        late_sub=cap["late"][:, ::4].contiguous().numpy(),
        pred_label=pred_label.numpy(),
        pred_mask_sub=sub(pred_mask, *((8, 8) if big else (2, 2))),
        pred_ss_sub=sub(pred_ss, *((8, 8) if big else (2, 2))),
        pos_neg_sum=np.asarray([int((cap["pos_neg_labels"] == v).sum()) for v in (0, 1, 2)]),
        pos_neg_sub=sub(cap["pos_neg_labels"], 4, 4).astype(np.int8),
        gt_label=gt_label.numpy(),
        loss=np.asarray(loss.detach().double().numpy()).reshape(-1),
    )
    if "logits" in cap:
        fx["logits"] = cap["logits"].numpy()
    if "class_labels" in cap:
        fx["class_sub"] = sub(cap["class_labels"], 4, 4).astype(np.int8)
    path = os.path.join(outdir, f"{name}.npz")
    np.savez_compressed(path, **fx)
    print(f"[{name}] wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count() or 1)
    names = sys.argv[1:] or list(GOLDEN)
    out = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out, exist_ok=True)
    for n in names:
        run_one(n, *GOLDEN[n], out)
