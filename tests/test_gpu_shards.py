"""Input pipeline on the GPU: the uint8 decode kernels against the fp32 transform kernel over ToTensor's output
(bit-exact), and the joint forward / training step fed by ShardLoader(device="cuda") against the same documents in the
reference's collate layout (float images): identical outputs."""
import dataclasses
import os
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "tests", "harness"))
import sroie_synth  # noqa: E402

pytestmark = pytest.mark.gpu

MEAN, STD = [0.9248, 0.9224, 0.9215], [0.1532, 0.1545, 0.1536]


def _to_tensor(u8):
    """torchvision ToTensor on an RGB uint8 HWC array."""
    return u8.permute(2, 0, 1).contiguous().to(torch.float32).div(255)


@pytest.mark.parametrize("geom", [((64, 96), (64, 96), (64, 96)), ((50, 70), (64, 90), (64, 96)), ((333, 777), (320, 746), (320, 768))])
def test_u8_decode_equals_float_transform(geom):
    """(h, w) -> (oh, ow) inside a padded (H, W) batch: same bits from the uint8 pixels as from ToTensor's floats."""
    from vibertgrid_pytorch_b200 import ops
    (h, w), (oh, ow), (H, W) = geom
    g = torch.Generator().manual_seed(h * 1000 + w)
    imgs = torch.randint(0, 256, (3, h, w, 3), generator=g, dtype=torch.uint8)
    want = torch.zeros(3, H + 6, W + 6, 4, device="cuda")
    ops.normalize_resize_pad_batch(torch.stack([_to_tensor(i) for i in imgs]).cuda(), want, 0, oh, ow, MEAN, STD)
    got = torch.zeros_like(want)
    ops.normalize_resize_pad_batch(imgs.cuda(), got, 0, oh, ow, MEAN, STD)
    assert torch.equal(got, want)
    one = torch.zeros_like(want)
    for b in range(3):
        ops.normalize_resize_pad(imgs[b].cuda(), one, b, oh, ow, MEAN, STD)
    assert torch.equal(one, want)


def test_u8_table_decode_ragged_batch():
    """Documents of different sizes, scattered over separate allocations, decoded by ONE table-driven launch."""
    from vibertgrid_pytorch_b200 import ops
    shapes = [((60, 100), (48, 80)), ((96, 64), (96, 64)), ((33, 47), (66, 94)), ((128, 128), (100, 100))]
    H, W = 128, 128
    g = torch.Generator().manual_seed(5)
    imgs = [torch.randint(0, 256, (h, w, 3), generator=g, dtype=torch.uint8) for (h, w), _ in shapes]
    sizes = [s for _, s in shapes]
    want = torch.zeros(len(imgs), H + 6, W + 6, 4, device="cuda")
    for b, im in enumerate(imgs):
        ops.normalize_resize_pad(_to_tensor(im).cuda(), want, b, sizes[b][0], sizes[b][1], MEAN, STD)
    pad = [torch.empty(1000 * (3 - b), device="cuda") for b in range(len(imgs))]       # spread the allocations
    dev = [im.cuda() for im in reversed(imgs)][::-1]                                    # later images at lower addresses too
    got = torch.zeros_like(want)
    tab = ops.image_table(dev, sizes)
    ops.decode_batch_u8(dev, tab, got, sizes, MEAN, STD)
    torch.cuda.synchronize()
    assert torch.equal(got, want)
    del pad


def _tree_and_shard(tmp, uniform, tokenizer_cfg):
    from transformers import BertTokenizer
    from vibertgrid_pytorch_b200 import shards, synth
    split = os.path.join(tmp, "data", "train")
    base = synth.CONFIGS["tiny"]
    geo = [(96, 128, 9)] * 2 if uniform else [(96, 128, 9), (64, 96, 5)]
    for seed, (h, w, segs) in enumerate(geo):
        sroie_synth.write_split(split, 2, dataclasses.replace(base, height=h, width=w, segments=segs), seed=seed, tokens_per_seg=3 + seed)
    tok = BertTokenizer.from_pretrained(synth.write_bert_dir(tokenizer_cfg, tmp))
    out = os.path.join(tmp, "train.vbgshard")
    shards.convert_sroie_split(split, tok, out, train=True)
    return out


def _float_batch(batch):
    """The same batch as the reference's dataset + collate would hand it over: ToTensor'd images."""
    return (tuple(_to_tensor(im) for im in batch[0]),) + tuple(batch[1:])


@pytest.mark.parametrize("uniform", [True, False])
def test_forward_from_shard_loader_equals_reference_layout(tmp_path, monkeypatch, uniform):
    from vibertgrid_pytorch_b200 import shards, synth
    from vibertgrid_pytorch_b200.net import ViBERTgridNet
    monkeypatch.chdir(tmp_path)
    cfg = synth.CONFIGS["tiny"]
    path = _tree_and_shard(str(tmp_path), uniform, cfg)
    net = ViBERTgridNet(**synth.model_kwargs(cfg, "eval"))
    synth.fill_state_dict_(net, 0)
    net = net.cuda().eval()
    docs = [[0, 1], [2, 3], [1, 0], [3, 2]] if uniform else [[0, 2], [3, 1], [2, 0], [1, 3]]
    host = list(shards.ShardLoader(path, batches=docs, depth=len(docs) + 1))                       # pinned host batches
    ld = shards.ShardLoader(path, batches=docs, device="cuda", depth=3)
    n = 0
    for i, batch in enumerate(ld):
        assert batch[0][0].dtype == torch.uint8 and batch[0][0].is_cuda and batch[4].is_cuda
        out_u8 = [t.clone() for t in net(*batch)]
        ref = _float_batch(host[i])
        ref = [tuple(t.cuda() for t in x) if isinstance(x, tuple) else x.cuda() for x in ref]
        out_f = net(*ref)
        for a, b in zip(out_u8, out_f):
            assert torch.equal(a, b), f"batch {i}: uint8-fed forward differs from the float-fed one"
        n += 1
    assert n == len(docs) and ld.h2d_bytes > 0
    if uniform:
        assert net._get_engine().graph_replays > 0, "repeated signatures are expected to replay the captured graph"


def test_training_step_from_shard_loader(tmp_path, monkeypatch):
    from vibertgrid_pytorch_b200 import shards, synth
    from vibertgrid_pytorch_b200.net import ViBERTgridNet
    monkeypatch.chdir(tmp_path)
    cfg = synth.CONFIGS["tiny"]
    path = _tree_and_shard(str(tmp_path), False, cfg)
    losses = {}
    for fmt in ("u8", "float"):
        torch.manual_seed(0)
        net = ViBERTgridNet(**synth.model_kwargs(cfg, "eval"))
        synth.fill_state_dict_(net, 0)
        net = net.cuda().train()
        net.bert_hidden_dropout = net.bert_attn_dropout = 0.0
        if fmt == "u8":
            batch = next(iter(shards.ShardLoader(path, batches=[[0, 3]], device="cuda")))
        else:       # ToTensor on the HOST like the reference's dataset (torch's CUDA div-by-scalar multiplies by the reciprocal)
            batch = _float_batch(next(iter(shards.ShardLoader(path, batches=[[0, 3]]))))
            batch = [tuple(t.cuda() for t in x) if isinstance(x, tuple) else x.cuda() for x in batch]
        loss = net(*batch)
        loss.backward()
        losses[fmt] = (float(loss), float(net.backbone.conv_1[0].weight.grad.abs().sum()) if hasattr(net.backbone, "conv_1") else 0.0)
    assert losses["u8"] == losses["float"]
