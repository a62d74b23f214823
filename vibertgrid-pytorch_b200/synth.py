"""Synthetic documents and deterministic weights.

Both are pure functions of integer seeds on the CPU generator, so the same
tensors can be regenerated on any box of this image: the golden fixtures under
``tests/golden/`` store only seeds + expected outputs, never weights.

Input format = what the reference's collate produces
(reference data/SROIE_dataset.py:141-148,184-197):
  image        tuple of B  f32[3,h,w] in [0,1]
  seg_indices  tuple of B  i32[n_tok]     non-decreasing run ids
  seg_classes  tuple of B  i32[S]
  coors        tuple of B  i64[S,4]       (left, top, right, bottom) pixels
  corpus       i64[B,L]    zero padded
  mask         i32[B,L]    corpus != 0
"""
from __future__ import annotations

import json
import math
import os
from dataclasses import dataclass, field
from typing import List, Optional

import torch


@dataclass
class DocConfig:
    """One BASELINE.json workload shape (per GPU)."""
    name: str = "cfg2"
    batch: int = 8
    height: int = 512
    width: int = 512
    seq_len: int = 512          # L = padded corpus width
    segments: int = 128         # S per image
    num_classes: int = 5
    backbone: str = "resnet_34_fpn"
    classifier_mode: str = "simp"
    vocab_size: int = 30522
    bert_layers: int = 12
    ragged: bool = False        # vary S / n_tok / image size per sample
    tag_to_idx: Optional[dict] = None
    bert_name: str = "bert-base-uncased"   # one of the reference's seven names (model/ViBERTgrid_net.py:218-226)


CONFIGS = {
    # BASELINE.json configs[0..4] (SURVEY 8 header)
    "cfg1": DocConfig("cfg1", 1, 512, 512, 512, 128, 5, "resnet_18_fpn"),
    "cfg2": DocConfig("cfg2", 8, 512, 512, 512, 128, 5, "resnet_34_fpn"),
    "cfg3": DocConfig("cfg3", 8, 512, 512, 512, 128, 5, "resnet_34_fpn"),
    "cfg4": DocConfig("cfg4", 4, 768, 768, 1024, 1024, 12, "resnet_34_fpn_pretrained", vocab_size=21128),
    "cfg5": DocConfig("cfg5", 2, 1024, 1024, 512, 64, 4, "resnet_34_fpn", classifier_mode="crf",
                      tag_to_idx={"O": 0, "B-q": 1, "B-a": 2, "B-h": 3}),
    # one document of the headline shapes: live-reference fixtures at the sizes the numbers are quoted on
    "cfg2_b1": DocConfig("cfg2_b1", 1, 512, 512, 512, 128, 5, "resnet_34_fpn"),
    "cfg4_b1": DocConfig("cfg4_b1", 1, 768, 768, 1024, 1024, 12, "resnet_34_fpn_pretrained", vocab_size=21128),
    "cfg5_b1": DocConfig("cfg5_b1", 1, 1024, 1024, 512, 64, 4, "resnet_34_fpn", classifier_mode="crf",
                         tag_to_idx={"O": 0, "B-q": 1, "B-a": 2, "B-h": 3}),
    # small shapes for the CPU suite / golden fixtures
    "tiny": DocConfig("tiny", 2, 96, 128, 40, 9, 5, "resnet_18_fpn", vocab_size=2000, bert_layers=2, ragged=True),
    "mid": DocConfig("mid", 4, 160, 192, 64, 12, 5, "resnet_18_fpn", vocab_size=2000, bert_layers=2, ragged=True),
    "tiny_d": DocConfig("tiny_d", 2, 64, 96, 24, 6, 4, "resnet_18_D_fpn", vocab_size=2000, bert_layers=1, ragged=True),
    "tiny_pre": DocConfig("tiny_pre", 1, 64, 64, 16, 5, 5, "resnet_18_fpn_pretrained", vocab_size=2000, bert_layers=1),
    "tiny_win": DocConfig("tiny_win", 2, 64, 64, 515, 12, 5, "resnet_18_fpn", vocab_size=2000, bert_layers=1, ragged=True),
    "tiny_rob": DocConfig("tiny_rob", 2, 64, 96, 24, 6, 5, "resnet_18_fpn", vocab_size=2000, bert_layers=2, ragged=True,
                          bert_name="roberta-base"),
}


def make_boxes(S: int, H: int, W: int, g: torch.Generator) -> torch.Tensor:
    """S boxes on a jittered lattice; ~5% tail of sub-stride and overlapping boxes
    so the empty-slice and last-writer-wins paths are exercised."""
    cols = max(1, min(8, W // 32))
    rows = max(1, math.ceil(S / cols))
    cw, ch = W / cols, H / rows
    out = torch.zeros(S, 4, dtype=torch.int64)
    for s in range(S):
        r, c = divmod(s, cols)
        r = r % rows
        x0 = c * cw + float(torch.rand(1, generator=g)) * cw * 0.15
        y0 = r * ch + float(torch.rand(1, generator=g)) * ch * 0.15
        x1 = (c + 1) * cw - float(torch.rand(1, generator=g)) * cw * 0.15
        y1 = (r + 1) * ch - float(torch.rand(1, generator=g)) * ch * 0.15
        u = float(torch.rand(1, generator=g))
        if u < 0.025:            # tiny box: < 8 px on a side -> may vanish at stride 8
            x1, y1 = x0 + 1 + 5 * float(torch.rand(1, generator=g)), y0 + 1 + 5 * float(torch.rand(1, generator=g))
        elif u < 0.05 and s > 0:  # overlaps its predecessor -> last writer wins
            px0, py0, px1, py1 = [float(v) for v in out[s - 1]]
            x0, y0 = (px0 + px1) / 2, (py0 + py1) / 2
            x1, y1 = x0 + cw * 0.8, y0 + ch * 0.8
        l, t = int(max(0, min(W - 2, x0))), int(max(0, min(H - 2, y0)))
        rr, bb = int(max(l + 1, min(W - 1, x1))), int(max(t + 1, min(H - 1, y1)))
        out[s] = torch.tensor([l, t, rr, bb])
    return out


def make_batch(cfg: DocConfig, seed: int = 0, device="cpu"):
    """Returns the six ``forward`` arguments as the reference's collate would."""
    g = torch.Generator().manual_seed(10_000 + seed)
    images, segs, classes, coors, n_toks = [], [], [], [], []
    for b in range(cfg.batch):
        h, w, S = cfg.height, cfg.width, cfg.segments
        if cfg.ragged and b > 0:
            h = max(32, h - 8 * b - 3)
            w = max(32, w - 16 * b - 5)
            S = max(2, S - 2 * b)
        tps = max(1, min(cfg.seq_len // max(S, 1), 8))
        if cfg.ragged:
            reps = torch.randint(1, tps + 1, (S,), generator=g)
            n_cap = cfg.seq_len if b == 0 else cfg.seq_len - 3 * b
            while int(reps.sum()) > n_cap:
                reps[int(torch.argmax(reps))] -= 1
            if b == 0 and int(reps.sum()) < cfg.seq_len:   # sample 0 fills L exactly
                reps[-1] += cfg.seq_len - int(reps.sum())
        else:
            reps = torch.full((S,), cfg.seq_len // S, dtype=torch.int64)
        images.append(torch.rand(3, h, w, generator=g))
        seg_ids = torch.repeat_interleave(torch.arange(S), reps).to(torch.int32)
        segs.append(seg_ids)
        classes.append(torch.randint(0, cfg.num_classes, (S,), generator=g).to(torch.int32))
        coors.append(make_boxes(S, h, w, g))
        n_toks.append(int(seg_ids.shape[0]))
    L = cfg.seq_len
    corpus = torch.zeros(cfg.batch, L, dtype=torch.int64)
    for b, n in enumerate(n_toks):
        assert n <= L
        corpus[b, :n] = torch.randint(1000, cfg.vocab_size, (n,), generator=g)
    if "roberta-" in cfg.bert_name and n_toks[0] > 3:
        corpus[0, 2] = 1        # RoBERTa's <pad> id inside the text: exercises the position-id rule's padding branch
    mask = (corpus != 0).to(torch.int32)
    mv = lambda seq: tuple(t.to(device) for t in seq)
    return mv(images), mv(segs), mv(classes), mv(coors), corpus.to(device), mask.to(device)


def fill_state_dict_(module: torch.nn.Module, seed: int = 0) -> None:
    """Deterministic, non-degenerate weights for every tensor of ``module``
    (in ``state_dict()`` order, CPU generator).  BatchNorm running statistics
    and affine terms are randomised so the folded-BN epilogues are exercised."""
    g = torch.Generator().manual_seed(777 + seed)
    seen = set()
    norm_types = (torch.nn.modules.batchnorm._BatchNorm, torch.nn.LayerNorm)
    with torch.no_grad():
        for name, t in module.state_dict().items():
            if t.data_ptr() in seen:       # aliased BERT copy
                continue
            seen.add(t.data_ptr())
            parent_name, _, leaf = name.rpartition(".")
            parent = module.get_submodule(parent_name) if parent_name else module
            in_bert = "bert_model" in name or "BERTgrid_generator" in name
            if leaf == "num_batches_tracked":
                t.fill_(1)
                continue
            if type(parent).__name__ == "LossWeightParams":     # class weights of the loss modules: constructor values
                continue
            if leaf == "running_mean":
                new = torch.randn(t.shape, generator=g) * 0.1
            elif leaf == "running_var":
                new = torch.rand(t.shape, generator=g) * 0.5 + 0.75
            elif isinstance(parent, norm_types):
                new = (torch.rand(t.shape, generator=g) * 0.5 + 0.75) if leaf == "weight" \
                    else torch.randn(t.shape, generator=g) * 0.05
            elif leaf == "transitions":
                new = torch.randn(t.shape, generator=g)
                T = t.shape[0]
                new[T - 2, :] = -10000.0
                new[:, T - 1] = -10000.0
            elif isinstance(parent, torch.nn.Embedding):
                new = torch.randn(t.shape, generator=g) * 0.05
            elif t.dim() >= 2:
                std = 0.04 if in_bert else math.sqrt(1.0 / t[0].numel())
                new = torch.randn(t.shape, generator=g) * std
            else:
                new = torch.randn(t.shape, generator=g) * 0.05
            t.copy_(new.to(t.dtype))


def bert_config_dict(cfg: DocConfig) -> dict:
    """The HuggingFace ``config.json`` of the stand-in BERT directory."""
    if "roberta-" in cfg.bert_name:
        return dict(model_type="roberta", architectures=["RobertaModel"], hidden_size=768,
                    num_hidden_layers=cfg.bert_layers, num_attention_heads=12, intermediate_size=3072, hidden_act="gelu",
                    layer_norm_eps=1e-5, max_position_embeddings=514, type_vocab_size=1, vocab_size=cfg.vocab_size,
                    hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1, pad_token_id=1, bos_token_id=0, eos_token_id=2)
    return dict(model_type="bert", architectures=["BertModel"], hidden_size=768,
                num_hidden_layers=cfg.bert_layers, num_attention_heads=12,
                intermediate_size=3072, hidden_act="gelu", layer_norm_eps=1e-12,
                max_position_embeddings=512, type_vocab_size=2, vocab_size=cfg.vocab_size,
                hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1, pad_token_id=0)


def write_bert_dir(cfg, root):
    """Stand-in HuggingFace directory literally named ``bert-base-uncased``:
    config.json with the workload's BERT hyper-parameters + a synthetic vocab."""
    d = os.path.join(root, cfg.bert_name)
    os.makedirs(d, exist_ok=True)
    with open(os.path.join(d, "config.json"), "w") as f:
        json.dump(bert_config_dict(cfg), f)
    if "roberta-" in cfg.bert_name:      # byte-level BPE stand-in: specials + single-character tokens, no merges
        vocab = {"<s>": 0, "<pad>": 1, "</s>": 2, "<unk>": 3}
        for i in range(4, cfg.vocab_size - 1):
            vocab[f"tok{i}"] = i
        vocab["<mask>"] = cfg.vocab_size - 1
        with open(os.path.join(d, "vocab.json"), "w") as f:
            json.dump(vocab, f)
        with open(os.path.join(d, "merges.txt"), "w") as f:
            f.write("#version: 0.2\n")
        return d
    special = {0: "[PAD]", 100: "[UNK]", 101: "[CLS]", 102: "[SEP]", 103: "[MASK]"}
    with open(os.path.join(d, "vocab.txt"), "w") as f:
        for i in range(cfg.vocab_size):
            f.write(special.get(i, f"tok{i}") + "\n")
    with open(os.path.join(d, "tokenizer_config.json"), "w") as f:
        json.dump({"do_lower_case": True, "model_max_length": 512}, f)
    return d


def model_kwargs(cfg: DocConfig, work_mode: str = "eval") -> dict:
    """Constructor kwargs for the reference-compatible ``ViBERTgridNet``."""
    kw = dict(num_classes=cfg.num_classes,
              image_mean=[0.9248, 0.9224, 0.9215], image_std=[0.1532, 0.1545, 0.1536],
              image_min_size=[min(cfg.height, cfg.width)], image_max_size=max(cfg.height, cfg.width),
              test_image_min_size=min(cfg.height, cfg.width),
              bert_model=cfg.bert_name, backbone=cfg.backbone,
              classifier_mode=cfg.classifier_mode, layer_mode="single",
              loss_control_lambda=1, ohem_random=True, work_mode=work_mode)
    if cfg.tag_to_idx is not None:
        kw["tag_to_idx"] = dict(cfg.tag_to_idx)
    if cfg.classifier_mode == "full":
        # the reference's BCELossRandomSample crashes on the default -1 sample counts
        # (custom_loss.py:260-264: random.sample(range(n), -1)); a count above the
        # population keeps every element and stays deterministic
        kw["num_hard_positive_main_1"] = kw["num_hard_negative_main_1"] = 1 << 20
    return kw
