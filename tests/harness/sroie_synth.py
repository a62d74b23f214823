"""Test harness: a synthetic on-disk SROIE tree in the layout the reference's dataset reads
(data/SROIE_dataset.py:88-153: ``<root>/{train,test}/{image/*.jpg, label/*.csv, key/*.json}``), a stand-in HuggingFace
directory (config + vocab, optionally seeded weights for ``from_pretrained``), a seeded checkpoint and the yaml config the
reference's train_SROIE.py / eval_SROIE.py parse.  Everything is a pure function of integer seeds."""
import csv
import dataclasses
import json
import os

import numpy as np
import torch
import yaml

from vibertgrid_pytorch_b200 import synth

CLASSES = ["others", "company", "date", "address", "total"]


def write_split(root, n_docs, cfg, seed, tokens_per_seg=4):
    from PIL import Image
    for sub in ("image", "label", "key"):
        os.makedirs(os.path.join(root, sub), exist_ok=True)
    g = torch.Generator().manual_seed(50_000 + seed)
    for d in range(n_docs):
        name = f"doc{seed:02d}_{d:03d}"
        img = (torch.rand(cfg.height, cfg.width, 3, generator=g) * 255).to(torch.uint8).numpy()
        Image.fromarray(img, "RGB").save(os.path.join(root, "image", name + ".jpg"), quality=92)
        boxes = synth.make_boxes(cfg.segments, cfg.height, cfg.width, g)
        ids = torch.randint(1000, cfg.vocab_size, (cfg.segments, tokens_per_seg), generator=g)
        cls = torch.randint(0, cfg.num_classes, (cfg.segments,), generator=g)
        with open(os.path.join(root, "label", name + ".csv"), "w", newline="") as f:
            w = csv.writer(f)
            w.writerow(["text", "left", "top", "right", "bot", "data_class"])
            for s in range(cfg.segments):
                w.writerow([" ".join(f"tok{int(i)}" for i in ids[s])] + [int(v) for v in boxes[s]] + [int(cls[s])])
        with open(os.path.join(root, "key", name + ".json"), "w") as f:
            json.dump({"company": f"tok{int(ids[0, 0])}", "date": "01/02/2019", "address": f"tok{int(ids[1, 0])}", "total": "12.50"}, f)


def write_bert(cfg, cwd, with_weights=False, seed=0, dropout=None):
    """Stand-in directory named like the checkpoint (resolved relative to the CWD by from_pretrained).  ``with_weights``
    adds a seeded ``model.safetensors`` so that work_mode="train" (which calls from_pretrained for the weights) works."""
    d = synth.write_bert_dir(cfg, cwd)
    cj = os.path.join(d, "config.json")
    conf = json.load(open(cj))
    if dropout is not None:
        conf["hidden_dropout_prob"] = conf["attention_probs_dropout_prob"] = float(dropout)
        json.dump(conf, open(cj, "w"))
    if with_weights:
        from transformers import BertConfig, BertModel
        torch.manual_seed(1234 + seed)
        BertModel(BertConfig(**conf)).save_pretrained(d)
    return d


def write_config(path, cfg, data_root, weights="", batch_size=2, end_epoch=1, sync_bn=True, amp=True, device="cuda"):
    hyp = dict(
        comment="harness", device=device, syncBN=sync_bn, amp=amp, start_epoch=0, end_epoch=end_epoch, batch_size=batch_size,
        optimizer_cnn_hyp=dict(learning_rate=0.005, min_learning_rate=1e-5, warm_up_epoches=0, warm_up_init_lr=1e-5, momentum=0.9,
                               weight_decay=0.005, min_weight_decay=0.005),
        optimizer_bert_hyp=dict(learning_rate=5e-5, min_learning_rate=1e-7, warm_up_epoches=0, warm_up_init_lr=1e-7, beta1=0.9,
                                beta2=0.999, epsilon=1e-8, weight_decay=0.01, min_weight_decay=0.01),
        loss_weights=None,
        num_hard_positive_main_1=-1, num_hard_negative_main_1=-1, num_hard_positive_main_2=-1, num_hard_negative_main_2=-1,
        loss_aux_sample_list=None, num_hard_positive_aux=-1, num_hard_negative_aux=-1, ohem_random=True,
        classifier_mode=cfg.classifier_mode, eval_mode="strcmp", tag_mode="B", bert_version=cfg.bert_name, backbone=cfg.backbone,
        grid_mode="mean", early_fusion_downsampling_ratio=8, roi_shape=7, p_fuse_downsampling_ratio=4,
        roi_align_output_reshape=False, late_fusion_fuse_embedding_channel=1024, layer_mode="single", loss_control_lambda=1,
        add_pos_neg=True, save_top=None, save_log="", weights=weights, num_workers=0, data_root=data_root,
        num_classes=cfg.num_classes, image_mean=[0.9248, 0.9224, 0.9215], image_std=[0.1532, 0.1545, 0.1536],
        image_min_size=[min(cfg.height, cfg.width)], image_max_size=max(cfg.height, cfg.width),
        test_image_min_size=min(cfg.height, cfg.width))
    with open(path, "w") as f:
        yaml.safe_dump(hyp, f)
    return hyp


def write_checkpoint(path, cfg, cwd, seed=0):
    """``torch.save({"model": state_dict})`` with DDP's ``module.`` prefix, as train_SROIE.py:381 saves and eval_SROIE.py:336
    strips.  Built from the drop-in's parameter tree (state-dict layout parity with the reference: tests/test_dropin.py)."""
    from vibertgrid_pytorch_b200.net import ViBERTgridNet
    here = os.getcwd()
    os.chdir(cwd)
    try:
        net = ViBERTgridNet(**synth.model_kwargs(cfg, "eval"))
    finally:
        os.chdir(here)
    synth.fill_state_dict_(net, seed)
    torch.save({"model": {"module." + k: v.clone() for k, v in net.state_dict().items()}}, path)


EVAL_DOCS = 4


def eval_case_config():
    """BASELINE configs[0]: one 512x512 image per step, resnet_18_fpn + bert-base-uncased (12 layers), 128 boxes x 4 tokens."""
    return dataclasses.replace(synth.CONFIGS["cfg1"])


def prepare_eval_case(tmp, device):
    """Synthetic test split + stand-in BERT directory + seeded checkpoint + yaml; returns the config path."""
    cfg = eval_case_config()
    write_split(os.path.join(tmp, "data", "test"), EVAL_DOCS, cfg, seed=1)
    write_bert(cfg, tmp)
    ck = os.path.join(tmp, "cfg1_seed0.pth")
    write_checkpoint(ck, cfg, tmp, seed=0)
    cpath = os.path.join(tmp, "eval.yaml")
    write_config(cpath, cfg, os.path.join(tmp, "data"), weights=ck, device=device)
    return cpath, os.path.join(tmp, "result", "cfg1_seed0.json")


def script_env(root, with_dropin):
    """PYTHONPATH for the reference's scripts: [dropin,] repo root (the package), the staged reference, the stubs."""
    ref = os.path.join(root, "oracle", "_ref", "reference")
    parts = ([os.path.join(root, "dropin")] if with_dropin else []) + [root, ref, os.path.join(root, "tests", "harness", "stubs")]
    env = dict(os.environ, PYTHONSAFEPATH="1", PYTHONPATH=os.pathsep.join(parts), HF_HUB_OFFLINE="1", TOKENIZERS_PARALLELISM="false")
    return env
