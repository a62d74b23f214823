"""Input pipeline of the hot path: shards of pre-tokenised, pre-decoded documents, a native reader / collate
(``csrc/vbg_shard.cpp`` behind the C-ABI) and a loader that yields the reference's collate layout.

What it replaces.  Per document and per epoch the reference's dataset (``data/SROIE_dataset.py:94-162``) decodes a JPEG with
PIL, parses a CSV with ``pandas.read_csv().iterrows()``, runs the Python WordPiece tokenizer over every OCR segment and builds
five tensors; its collate (``:165-208``) pads the token ids and derives the mask; the training loop then moves ~34 tensors of
a batch to the device one ``.to(device)`` at a time (``pipeline/train_val_utils.py:257-262``).  At >1 000 documents/s per GPU
that path cannot feed one B200, let alone eight.  Here the per-document work is done ONCE, offline:

    convert_sroie_split(split_dir, tokenizer, "train.vbgshard", train=True)       # same filter / tokenise rules, restated

and an epoch is: memory-map the shard, and per batch one C call that gathers the documents into ONE pinned staging buffer
(uint8 pixels -- a quarter of ToTensor's fp32 bytes; ``byte / 255`` happens in the decode kernel, bit-identically) followed
by ONE host->device copy on a side stream, overlapped with the previous step's kernels:

    for image, seg, cls, coors, corpus, mask in ShardLoader("train.vbgshard", batch_size=8, device="cuda"):
        loss = net(image, seg, cls, coors, corpus, mask)

The yielded tuple has the reference's collate layout (tuples of per-document tensors + the padded ``corpus`` / ``mask``;
``train=False`` appends the ``ocr_text`` and ``key_dict`` tuples of the eval layout), except that an image is the decoded
``uint8 [h, w, 3]`` array instead of ``ToTensor``'s ``float32 [3, h, w]`` -- ``ViBERTgridNet`` accepts both and produces
bit-identical results.  With ``device=None`` the loader yields pinned host tensors, so the reference's own loops (which call
``.to(device)`` themselves) can consume it unchanged.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import queue
import struct
import threading
from typing import Iterable, Iterator, List, Optional, Sequence

import numpy as np
import torch

from . import _lib as L

MAGIC = b"VBGSHRD1"
_FILE_HEADER = struct.Struct("<8sIIQQI28x")          # magic, version, n_docs, index_offset, file_bytes, flags
_DOC_HEADER = struct.Struct("<5i12x")                # h, w, n_tok, n_seg, meta_bytes
assert _FILE_HEADER.size == 64 and _DOC_HEADER.size == 32


def _pad(f, align):
    f.write(b"\0" * (-f.tell() % align))


# ------------------------------------------------------------------ offline conversion
def sroie_document(split_dir: str, file: str, tokenizer, train: bool) -> dict:
    """One document exactly as the reference's ``SROIEDataset.__getitem__`` would produce it (data/SROIE_dataset.py:94-162),
    before ``ToTensor``: the decoded RGB bytes, the filtered segments and their token ids.  Rules restated: the image is
    converted to RGB unless it already has three bands; every CSV row is a segment (``text, left, top, right, bot,
    data_class``); segments whose text is empty / blank, or tokenises to nothing after lower-casing, are dropped and the
    remaining ones renumbered consecutively; every token carries the index of its (renumbered) segment."""
    import pandas as pd
    from PIL import Image
    image = Image.open(os.path.join(split_dir, "image", file))
    if len(image.split()) != 3:
        image = image.convert("RGB")
    pixels = np.asarray(image, dtype=np.uint8)
    rows = pd.read_csv(os.path.join(split_dir, "label", file.replace("jpg", "csv")))
    corpus_tokens, seg_ids, coors, classes, texts = [], [], [], [], []
    for _, row in rows.iterrows():                          # iterrows: the reference's own value coercion (mixed-type rows)
        text = str(row["text"])
        if text == "" or text.isspace():
            continue
        tokens = tokenizer.tokenize(text.lower())
        if len(tokens) == 0:
            continue
        seg_ids += [len(texts)] * len(tokens)
        corpus_tokens += tokens
        texts.append(text)
        coors.append([row["left"], row["top"], row["right"], row["bot"]])
        classes.append(row["data_class"])
    doc = dict(image=pixels,
               corpus=torch.tensor(tokenizer.convert_tokens_to_ids(corpus_tokens), dtype=torch.long).numpy().astype(np.int32),
               seg_ids=np.asarray(seg_ids, dtype=np.int32),
               classes=torch.tensor(classes, dtype=torch.int).numpy().reshape(-1),
               coors=torch.tensor(coors, dtype=torch.long).numpy().reshape(-1, 4))
    if not train:                                           # the eval layout's side data (:150-162)
        with open(os.path.join(split_dir, "key", file.replace(".jpg", ".json"))) as f:
            key = json.load(f)
        key.update({"filename": file.replace(".jpg", "")})
        doc["meta"] = {"text": texts, "key": key}
    return doc


def write_shard(path: str, docs: Iterable[dict]) -> int:
    """Writes documents (dicts with ``image`` uint8 [h,w,3], ``corpus`` / ``seg_ids`` int32 [n_tok], ``classes`` int32 [n_seg],
    ``coors`` int64 [n_seg,4] and optionally ``meta``, a JSON-serialisable object) in the layout csrc/vbg_shard.cpp reads."""
    offsets = []
    with open(path, "wb") as f:
        f.write(b"\0" * _FILE_HEADER.size)
        for d in docs:
            img = np.ascontiguousarray(d["image"], dtype=np.uint8)
            if img.ndim != 3 or img.shape[2] != 3:
                raise ValueError(f"shard images are uint8 [h, w, 3]; got {img.shape}")
            corpus = np.ascontiguousarray(d["corpus"], dtype=np.int32).reshape(-1)
            seg_ids = np.ascontiguousarray(d["seg_ids"], dtype=np.int32).reshape(-1)
            classes = np.ascontiguousarray(d["classes"], dtype=np.int32).reshape(-1)
            coors = np.ascontiguousarray(d["coors"], dtype=np.int64).reshape(-1, 4)
            if seg_ids.shape != corpus.shape or coors.shape[0] != classes.shape[0]:
                raise ValueError("shard document: corpus / seg_ids and coors / classes must have matching lengths")
            meta = json.dumps(d["meta"]).encode() if d.get("meta") is not None else b""
            _pad(f, 64)
            offsets.append(f.tell())
            f.write(_DOC_HEADER.pack(img.shape[0], img.shape[1], corpus.shape[0], classes.shape[0], len(meta)))
            f.write(corpus.tobytes()); f.write(seg_ids.tobytes()); f.write(classes.tobytes())
            _pad(f, 8)
            f.write(coors.tobytes()); f.write(meta)
            _pad(f, 64)
            f.write(img.tobytes())
        _pad(f, 64)
        index_offset = f.tell()
        f.write(np.asarray(offsets, dtype=np.uint64).tobytes())
        total = f.tell()
        f.seek(0)
        f.write(_FILE_HEADER.pack(MAGIC, 1, len(offsets), index_offset, total, 0))
    return len(offsets)


def convert_sroie_split(split_dir: str, tokenizer, out_path: str, train: bool = True, files: Optional[Sequence[str]] = None) -> int:
    """``<split_dir>/{image,label[,key]}`` (the tree data/SROIE_dataset.py:88-92 lists) -> one shard, documents in the order of
    ``os.listdir(image)`` -- the reference dataset's own index order -- unless ``files`` is given."""
    names = list(files) if files is not None else [f for f in os.listdir(os.path.join(split_dir, "image"))]
    return write_shard(out_path, (sroie_document(split_dir, f, tokenizer, train) for f in names))


def convert_dataset(dataset, out_path: str, train: bool = True, indices: Optional[Sequence[int]] = None) -> int:
    """Any map-style dataset in the reference's item layout -- ``(ToTensor image f32 [3,h,w], seg_indices, seg_classes, coors,
    corpus[, ocr_text, key_dict])``, i.e. the reference's own ``SROIEDataset`` / ``EPHOIEDataset`` / ``FUNSDDataset``
    (data/*_dataset.py) -- -> one shard, offline: each item is produced once by the dataset's own code and stored; the pixels
    go back to the bytes ``ToTensor`` divided by 255 (exact: ``round(x * 255)`` inverts ``byte / 255`` for every byte, checked)."""
    def docs():
        for i in (range(len(dataset)) if indices is None else indices):
            item = dataset[i]
            img = item[0]
            u8 = (img * 255.0).round().clamp_(0, 255).to(torch.uint8)
            if not torch.equal(u8.to(torch.float32).div(255), img):
                raise ValueError(f"document {i}: the image is not ToTensor of 8-bit pixels; shards hold uint8 pixels")
            d = dict(image=u8.permute(1, 2, 0).contiguous().numpy(), seg_ids=item[1].numpy(), classes=item[2].numpy().reshape(-1),
                     coors=item[3].numpy().reshape(-1, 4), corpus=item[4].numpy())
            if not train:                                   # SROIE / EPHOIE: (..., ocr_text, key_dict); FUNSD: (..., ocr_text)
                d["meta"] = {"text": list(item[5])}
                if len(item) > 6:
                    d["meta"]["key"] = item[6]
            yield d
    return write_shard(out_path, docs())


# ------------------------------------------------------------------ native reader
_LAYOUT_N = 13


class Shard:
    """A memory-mapped shard (read-only).  ``len(shard)`` documents; ``shape(i)`` = (h, w, n_tok, n_seg); ``meta(i)`` the eval
    side data.  Usable as the ``data_source`` of torch samplers (``DistributedSampler(shard)``)."""

    def __init__(self, path: str):
        self.path = path
        self._h = C.c_void_p()
        L.check(L.load().vbg_shard_open(os.fsencode(path), C.byref(self._h)), "vbg_shard_open", launch=False)

    def close(self):
        if self._h:
            L.load().vbg_shard_close(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self):
        return int(L.load().vbg_shard_num_docs(self._h))

    def shape(self, i: int):
        out = (C.c_int32 * 4)()
        L.check(L.load().vbg_shard_doc_shape(self._h, int(i), out), "vbg_shard_doc_shape", launch=False)
        return tuple(out)

    def meta(self, i: int):
        p, n = C.c_void_p(), C.c_longlong()
        L.check(L.load().vbg_shard_doc_meta(self._h, int(i), C.byref(p), C.byref(n)), "vbg_shard_doc_meta", launch=False)
        return json.loads(C.string_at(p, n.value).decode()) if n.value else None

    def layout(self, docs: Sequence[int]):
        ids = (C.c_int32 * len(docs))(*[int(d) for d in docs])
        lay = (C.c_int64 * _LAYOUT_N)()
        L.check(L.load().vbg_shard_batch_layout(self._h, ids, len(docs), lay), "vbg_shard_batch_layout", launch=False)
        return list(lay)

    def collate_into(self, docs: Sequence[int], staging: torch.Tensor, threads: int = 4):
        """Gathers ``docs`` into ``staging`` (a uint8 host tensor, ideally pinned; the GIL is released during the call)."""
        ids = (C.c_int32 * len(docs))(*[int(d) for d in docs])
        L.check(L.load().vbg_shard_collate(self._h, ids, len(docs), C.c_void_p(staging.data_ptr()), staging.numel(), int(threads)),
                "vbg_shard_collate", launch=False)


def batch_views(buf: torch.Tensor, lay: Sequence[int], shapes: Sequence[Sequence[int]]):
    """The reference collate layout as zero-copy views of one collated buffer (host staging or its device copy)."""
    (_, Lw, _, _, o_corpus, o_mask, o_seg, o_cls, o_coors, _, _, o_arena, _) = lay
    B = len(shapes)

    def view(off, count, dtype):
        nbytes = count * torch.empty((), dtype=dtype).element_size()
        return buf[off:off + nbytes].view(dtype)

    corpus = view(o_corpus, B * Lw, torch.int64).view(B, Lw)
    mask = view(o_mask, B * Lw, torch.int32).view(B, Lw)
    images, seg, cls, coors = [], [], [], []
    t0 = s0 = a0 = 0
    for h, w, n_tok, n_seg in shapes:
        images.append(buf[o_arena + a0:o_arena + a0 + 3 * h * w].view(h, w, 3))
        seg.append(view(o_seg + 4 * t0, n_tok, torch.int32))
        cls.append(view(o_cls + 4 * s0, n_seg, torch.int32))
        coors.append(view(o_coors + 32 * s0, 4 * n_seg, torch.int64).view(n_seg, 4))
        t0 += n_tok; s0 += n_seg; a0 += -(-3 * h * w // 64) * 64
    return tuple(images), tuple(seg), tuple(cls), tuple(coors), corpus, mask


def default_batches(n_docs: int, batch_size: int, rank: int = 0, world: int = 1, shuffle: bool = False, seed: int = 0,
                    epoch: int = 0, drop_last: bool = True) -> List[List[int]]:
    """Index lists of one epoch for one rank: ``DistributedSampler`` (pad by wrap-around to a multiple of ``world``, strided
    assignment, ``torch.randperm`` seeded by ``seed + epoch`` when shuffling) followed by ``BatchSampler(drop_last)`` --
    the samplers of the reference's multi-GPU loader (data/SROIE_dataset.py:314-318)."""
    if shuffle:
        g = torch.Generator()
        g.manual_seed(seed + epoch)
        order = torch.randperm(n_docs, generator=g).tolist()
    else:
        order = list(range(n_docs))
    if world > 1:
        total = -(-n_docs // world) * world
        order = (order + order[:total - len(order)]) if total > len(order) else order
        while len(order) < total:                           # fewer documents than ranks
            order += order[:total - len(order)]
        order = order[rank:total:world]
    out = [order[i:i + batch_size] for i in range(0, len(order), batch_size)]
    if drop_last and out and len(out[-1]) < batch_size:
        out.pop()
    return out


class ShardLoader:
    """Iterates batches of a shard in the reference's collate layout (see the module docstring).

    ``batches``: any RE-ITERABLE of index lists (e.g. ``torch.utils.data.BatchSampler(DistributedSampler(shard), bs, True)``;
    it is walked once per epoch and once by ``len()``, so a one-shot generator will not do);
    default: ``default_batches`` for (rank, world).  ``device``: a CUDA device -> batches arrive on the device, each by ONE
    asynchronous copy of the pinned staging buffer on a side stream, prepared ``depth - 1`` batches ahead by a background
    thread; ``None`` -> pinned host tensors.  A yielded batch stays valid until ``depth - 1`` further batches were requested.
    """

    def __init__(self, shard, batch_size: int = 1, *, batches: Optional[Iterable[Sequence[int]]] = None, rank: int = 0, world: int = 1,
                 shuffle: bool = False, seed: int = 0, drop_last: bool = True, train: bool = True, device=None, depth: int = 3,
                 threads: int = 4):
        self.shard = shard if isinstance(shard, Shard) else Shard(shard)
        self.batch_size, self.rank, self.world = int(batch_size), int(rank), int(world)
        self.shuffle, self.seed, self.drop_last, self.train = bool(shuffle), int(seed), bool(drop_last), bool(train)
        self.batches = batches
        self.device = None if device is None else torch.device(device)
        if self.device is not None and self.device.type != "cuda":
            raise RuntimeError("ShardLoader uploads to a CUDA device (device=None yields pinned host batches)")
        if self.device is not None and self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.depth, self.threads = max(2, int(depth)), int(threads)
        self.epoch = 0
        self.h2d_bytes = 0                                  # staged bytes copied to the device so far (bench.py reads it)
        self._slots = []

    def set_epoch(self, epoch: int):
        self.epoch = int(epoch)

    def _epoch_batches(self):
        if self.batches is not None:
            return [list(b) for b in self.batches]
        return default_batches(len(self.shard), self.batch_size, self.rank, self.world, self.shuffle, self.seed, self.epoch,
                               self.drop_last)

    def __len__(self):
        return len(self._epoch_batches())

    def _slot(self, i, nbytes):
        while len(self._slots) <= i:
            self._slots.append({"host": None, "dev": None, "free": None})
        s = self._slots[i]
        if s["host"] is None or s["host"].numel() < nbytes:
            cap = -(-nbytes // (1 << 20)) * (1 << 20)
            pin = torch.cuda.is_available()
            s["host"] = torch.empty(cap, dtype=torch.uint8, pin_memory=pin)
            if self.device is not None:
                if s["dev"] is not None:                    # growing a slot (rare): nothing may still read the old buffer
                    torch.cuda.synchronize(self.device)
                s["dev"] = torch.empty(cap, dtype=torch.uint8, device=self.device)
                s["free"] = s["copied"] = None
        return s

    def _finish(self, docs, views):
        if self.train:
            return views
        metas = [self.shard.meta(d) or {} for d in docs]
        # SROIE / EPHOIE: (..., ocr_text, key_dicts); FUNSD has no key dictionaries and its collate returns None in that slot
        keys = tuple(m.get("key", {}) for m in metas) if any("key" in m for m in metas) else None
        return views + (tuple(m.get("text", []) for m in metas), keys)

    def __iter__(self) -> Iterator[tuple]:
        todo = self._epoch_batches()
        if not todo:
            return
        if self.device is None:
            for i, docs in enumerate(todo):
                lay = self.shard.layout(docs)
                slot = self._slot(i % self.depth, lay[0])
                self.shard.collate_into(docs, slot["host"], self.threads)
                yield self._finish(docs, batch_views(slot["host"], lay, [self.shard.shape(d) for d in docs]))
            return
        yield from self._iter_device(todo)

    def _iter_device(self, todo):
        dev = self.device
        stream = torch.cuda.Stream(dev)
        ready: "queue.Queue" = queue.Queue()
        released: "queue.Queue" = queue.Queue()
        for i in range(self.depth):
            released.put(i)
        stop = threading.Event()

        def produce():
            try:
                torch.cuda.set_device(dev)
                for docs in todo:
                    i = released.get()
                    if stop.is_set() or i is None:
                        return
                    lay = self.shard.layout(docs)
                    slot = self._slot(i, lay[0])
                    if slot.get("copied") is not None:
                        slot["copied"].synchronize()         # HOST wait: the upload that last read this pinned buffer has executed
                    self.shard.collate_into(docs, slot["host"], self.threads)
                    if slot["free"] is not None:
                        stream.wait_event(slot["free"])      # the consumer's kernels on this slot's previous batch
                    with torch.cuda.stream(stream):
                        slot["dev"][:lay[0]].copy_(slot["host"][:lay[0]], non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(stream)
                    slot["copied"] = ev
                    ready.put((i, docs, lay, ev))
                ready.put(None)
            except BaseException as e:                       # surface producer failures in the consumer
                ready.put(e)

        th = threading.Thread(target=produce, name="vbg-shard-loader", daemon=True)
        th.start()
        held = []                                           # slots handed to the consumer, oldest first
        try:
            while True:
                item = ready.get()
                if item is None:
                    break
                if isinstance(item, BaseException):
                    raise item
                i, docs, lay, ev = item
                cur = torch.cuda.current_stream(dev)
                while len(held) >= self.depth - 1:           # everything the consumer does with the oldest batch is enqueued
                    old = held.pop(0)
                    self._slots[old]["free"] = torch.cuda.Event()
                    self._slots[old]["free"].record(cur)
                    released.put(old)
                cur.wait_event(ev)
                self.h2d_bytes += int(lay[0])
                held.append(i)
                yield self._finish(docs, batch_views(self._slots[i]["dev"], lay, [self.shard.shape(d) for d in docs]))
        finally:
            stop.set()
            released.put(None)
            th.join(timeout=10)
