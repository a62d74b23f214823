"""Input pipeline on the CPU: the shard converter + native reader / collate against the UNMODIFIED reference
dataset and collate function (oracle/_ref/reference/data/SROIE_dataset.py, staged by __graft_entry__.build()) on a synthetic
on-disk SROIE tree; the default sampler against torch's DistributedSampler + BatchSampler; error behaviour of the reader."""
import csv
import dataclasses
import json
import os
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "tests", "harness"))
import sroie_synth  # noqa: E402

from vibertgrid_pytorch_b200 import _lib, shards, synth  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "reference")


def make_tree(tmp, train, edge_rows=True):
    """Documents of three different sizes plus the rows the reference's filter drops (blank text, text that tokenises to nothing).
    ``edge_rows=False``: without the blank / numeric rows (the reference's EPHOIE / FUNSD datasets raise on a NaN text cell)."""
    split = os.path.join(tmp, "data", "train" if train else "test")
    base = synth.CONFIGS["tiny"]
    for seed, (h, w, segs) in enumerate([(96, 128, 9), (64, 96, 5), (80, 72, 12)]):
        sroie_synth.write_split(split, 2, dataclasses.replace(base, height=h, width=w, segments=segs), seed=seed, tokens_per_seg=3 + seed)
    # edge rows in one label file: blank text, whitespace, a numeric text (pandas parses it as a number) and an unknown word
    name = sorted(os.listdir(os.path.join(split, "label")))[0]
    path = os.path.join(split, "label", name)
    rows = list(csv.reader(open(path)))
    if edge_rows:
        rows.insert(2, ["", 1, 2, 30, 40, 1])
        rows.insert(4, ["   ", 3, 4, 50, 60, 2])
        rows.append(["12.50", 5, 6, 70, 80, 4])
    rows.append(["Zebra tok1500", 7, 8, 90, 95, 3])
    with open(path, "w", newline="") as f:
        csv.writer(f).writerows(rows)
    # a grey-scale image: the reference converts anything without exactly three bands to RGB
    from PIL import Image
    img = sorted(os.listdir(os.path.join(split, "image")))[1]
    Image.open(os.path.join(split, "image", img)).convert("L").save(os.path.join(split, "image", img), quality=90)
    return split


@pytest.fixture
def tokenizer(tmp_path):
    from transformers import BertTokenizer
    d = synth.write_bert_dir(synth.CONFIGS["tiny"], str(tmp_path))
    return BertTokenizer.from_pretrained(d)


def reference_dataset(split, tokenizer, train):
    if not os.path.isdir(REF):
        pytest.skip("oracle/_ref/reference is not staged (run __graft_entry__.build() where /root/reference is mounted)")
    if REF not in sys.path:
        sys.path.append(REF)                      # after the repo: only `data.*` resolves there
    from data.SROIE_dataset import SROIEDataset
    return SROIEDataset(split, train=train, tokenizer=tokenizer)


@pytest.mark.parametrize("train", [True, False])
def test_shard_documents_match_reference_dataset(tmp_path, tokenizer, train):
    split = make_tree(str(tmp_path), train)
    ds = reference_dataset(split, tokenizer, train)
    out = str(tmp_path / "split.vbgshard")
    n = shards.convert_sroie_split(split, tokenizer, out, train=train, files=ds.filename_list)
    sh = shards.Shard(out)
    assert n == len(ds) == len(sh) == 6
    ld = shards.ShardLoader(sh, batch_size=1, train=train, drop_last=False)
    sizes = set()
    for i, got in enumerate(ld):
        want = ds[i]
        img, seg, cls, coors, corpus = want[:5]
        assert got[0][0].dtype == torch.uint8 and got[0][0].shape == (img.shape[1], img.shape[2], 3)
        # ToTensor == byte / 255: the decode kernel's arithmetic, checked here with torch's own division
        assert torch.equal(got[0][0].permute(2, 0, 1).to(torch.float32).div(255), img)
        assert got[1][0].dtype == seg.dtype and torch.equal(got[1][0], seg)
        assert got[2][0].dtype == cls.dtype and torch.equal(got[2][0], cls)
        assert got[3][0].dtype == coors.dtype and torch.equal(got[3][0], coors)
        assert got[4].dtype == corpus.dtype and torch.equal(got[4][0], corpus)
        assert torch.equal(got[5][0], (corpus != 0).int())
        assert sh.shape(i) == (img.shape[1], img.shape[2], corpus.shape[0], cls.shape[0])
        if not train:
            assert list(got[6][0]) == list(want[5]) and got[7][0] == want[6]
        sizes.add(tuple(img.shape))
    assert len(sizes) == 3, "the tree is meant to hold documents of different sizes"
    sh.close()


@pytest.mark.parametrize("train", [True, False])
def test_collate_matches_reference_collate(tmp_path, tokenizer, train):
    split = make_tree(str(tmp_path), train)
    ds = reference_dataset(split, tokenizer, train)
    out = str(tmp_path / "split.vbgshard")
    shards.convert_sroie_split(split, tokenizer, out, train=train, files=ds.filename_list)
    for docs in ([0, 3, 5], [4, 1], [2, 2, 0, 1]):
        want = ds._ViBERTgrid_coll_func([ds[i] for i in docs])
        got = next(iter(shards.ShardLoader(out, batches=[docs], train=train, threads=3)))
        assert len(got) == len(want) == (6 if train else 8)
        for a, b in zip(got[0], want[0]):
            assert torch.equal(a.permute(2, 0, 1).float().div(255), b)
        for k in (1, 2, 3):
            assert len(got[k]) == len(want[k])
            for a, b in zip(got[k], want[k]):
                assert a.dtype == b.dtype and torch.equal(a, b)
        assert got[4].dtype == want[4].dtype and torch.equal(got[4], want[4])            # pad_sequence: zero padded to the longest
        assert got[5].dtype == want[5].dtype and torch.equal(got[5], want[5])            # mask.int()
        if not train:
            assert [list(t) for t in got[6]] == [list(t) for t in want[6]] and list(got[7]) == list(want[7])


@pytest.mark.parametrize("train", [True, False])
def test_convert_dataset_equals_direct_conversion(tmp_path, tokenizer, train):
    """The generic converter (consumes the reference dataset's own items: works for its SROIE / EPHOIE / FUNSD datasets alike)
    writes byte for byte the shard the restated SROIE rules write."""
    split = make_tree(str(tmp_path), train)
    ds = reference_dataset(split, tokenizer, train)
    a, b = str(tmp_path / "a.vbgshard"), str(tmp_path / "b.vbgshard")
    assert shards.convert_sroie_split(split, tokenizer, a, train=train, files=ds.filename_list) == len(ds)
    assert shards.convert_dataset(ds, b, train=train) == len(ds)
    assert open(a, "rb").read() == open(b, "rb").read()


def _other_tree(tmp, kind):
    """The SROIE-shaped synthetic split rearranged into the EPHOIE / FUNSD on-disk layouts (data/EPHOIE_dataset.py:96-115,
    data/FUNSD_dataset.py:87-106)."""
    import shutil
    from PIL import Image
    src = make_tree(os.path.join(tmp, "src"), False, edge_rows=False)
    root = os.path.join(tmp, kind)
    names = sorted(f[:-4] for f in os.listdir(os.path.join(src, "image")))
    if kind == "ephoie":
        for sub in ("image", "_label_csv", "kvpair"):
            os.makedirs(os.path.join(root, sub))
        for n in names:
            shutil.copy(os.path.join(src, "image", n + ".jpg"), os.path.join(root, "image", n + ".jpg"))
            shutil.copy(os.path.join(src, "label", n + ".csv"), os.path.join(root, "_label_csv", n + ".csv"))
            shutil.copy(os.path.join(src, "key", n + ".json"), os.path.join(root, "kvpair", n + ".txt"))
        for split in ("train.txt", "test.txt"):
            open(os.path.join(root, split), "w").write("\n".join(names) + "\n")
    else:
        for sub in ("images", "_label_csv"):
            os.makedirs(os.path.join(root, "training_data", sub))
        for n in names:
            Image.open(os.path.join(src, "image", n + ".jpg")).save(os.path.join(root, "training_data", "images", n + ".png"))
            shutil.copy(os.path.join(src, "label", n + ".csv"), os.path.join(root, "training_data", "_label_csv", n + ".csv"))
    return root


@pytest.mark.parametrize("kind,train", [("ephoie", True), ("ephoie", False), ("funsd", True), ("funsd", False)])
def test_generic_converter_on_the_other_reference_datasets(tmp_path, tokenizer, kind, train):
    """convert_dataset + ShardLoader against the reference's EPHOIE / FUNSD datasets and their collate functions."""
    if not os.path.isdir(REF):
        pytest.skip("oracle/_ref/reference is not staged")
    if REF not in sys.path:
        sys.path.append(REF)
    root = _other_tree(str(tmp_path), kind)
    if kind == "ephoie":
        from data.EPHOIE_dataset import EPHOIEDataset as DS
    else:
        from data.FUNSD_dataset import FUNSDDataset as DS
    ds = DS(root, train=train, tokenizer=tokenizer)
    out = str(tmp_path / f"{kind}.vbgshard")
    assert shards.convert_dataset(ds, out, train=train) == len(ds) == 6
    for docs in ([0, 4, 5], [3, 1]):
        want = ds._ViBERTgrid_coll_func([ds[i] for i in docs])
        got = next(iter(shards.ShardLoader(out, batches=[docs], train=train)))
        assert len(got) == len(want)
        for a, b in zip(got[0], want[0]):
            assert torch.equal(a.permute(2, 0, 1).float().div(255), b)
        for k in (1, 2, 3):
            for a, b in zip(got[k], want[k]):
                assert a.dtype == b.dtype and torch.equal(a, b)
        assert torch.equal(got[4], want[4]) and torch.equal(got[5], want[5]) and got[5].dtype == want[5].dtype
        for k in range(6, len(want)):
            if want[k] is None:
                assert got[k] is None
            else:
                assert [x for x in got[k]] == [x for x in want[k]]


@pytest.mark.parametrize("n,world,bs,shuffle", [(10, 1, 3, False), (10, 4, 2, True), (7, 2, 2, True), (3, 4, 1, False), (16, 8, 2, True)])
def test_default_batches_are_torch_samplers(n, world, bs, shuffle):
    from torch.utils.data import BatchSampler, DistributedSampler
    for rank in range(world):
        for epoch in (0, 3):
            s = DistributedSampler(range(n), num_replicas=world, rank=rank, shuffle=shuffle, seed=11)
            s.set_epoch(epoch)
            want = [list(b) for b in BatchSampler(s, batch_size=bs, drop_last=True)]
            got = shards.default_batches(n, bs, rank, world, shuffle, seed=11, epoch=epoch, drop_last=True)
            assert got == want, (rank, epoch)


def test_loader_accepts_torch_batch_sampler(tmp_path, tokenizer):
    from torch.utils.data import BatchSampler, DistributedSampler
    split = make_tree(str(tmp_path), True)
    out = str(tmp_path / "s.vbgshard")
    shards.convert_sroie_split(split, tokenizer, out, train=True)
    sh = shards.Shard(out)
    bsamp = BatchSampler(DistributedSampler(sh, num_replicas=2, rank=1, shuffle=False), batch_size=2, drop_last=True)
    seen = [tuple(int(s.shape[0]) for s in b[1]) for b in shards.ShardLoader(sh, batches=bsamp)]
    assert seen == [tuple(sh.shape(d)[2] for d in docs) for docs in bsamp] and len(seen) == 1


def test_reader_rejects_bad_input(tmp_path, tokenizer):
    split = make_tree(str(tmp_path), True)
    out = str(tmp_path / "s.vbgshard")
    shards.convert_sroie_split(split, tokenizer, out, train=True)
    with pytest.raises(_lib.VbgError, match="cannot open"):
        shards.Shard(str(tmp_path / "missing.vbgshard"))
    raw = open(out, "rb").read()
    bad = str(tmp_path / "bad.vbgshard")
    open(bad, "wb").write(b"NOTASHRD" + raw[8:])
    with pytest.raises(_lib.VbgError, match="not a version-1 shard"):
        shards.Shard(bad)
    open(bad, "wb").write(raw[:len(raw) // 2])                      # truncated: the header's size no longer matches
    with pytest.raises(_lib.VbgError, match="not a version-1 shard"):
        shards.Shard(bad)
    sh = shards.Shard(out)
    with pytest.raises(_lib.VbgError):
        sh.shape(99)
    with pytest.raises(_lib.VbgError, match="staging buffer"):
        sh.collate_into([0, 1], torch.empty(128, dtype=torch.uint8))
    with pytest.raises(ValueError):
        shards.write_shard(bad, [dict(image=np.zeros((4, 4), np.uint8), corpus=[1], seg_ids=[0], classes=[0], coors=[[0, 0, 1, 1]])])
