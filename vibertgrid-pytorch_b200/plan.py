"""Host-side batch planning: everything the reference obtains through
``.item()`` / ``int(tensor)`` device syncs is derived here from tensor SHAPES
only (SURVEY.md 7.2 "No host syncs in forward"), packed into one int32 table
and uploaded with a single H2D copy.

Pure numpy / Python -- unit-tested on CPU against the oracle
(tests/test_host_logic.py).

Reference semantics restated:
  * resize scale / output size   pipeline/transform.py:137-155
  * coord ratios                  pipeline/transform.py:163
  * padded batch shape            pipeline/transform.py:248-255
  * 510-token windows, [SEP] position after the dataset padding
                                  model/BERTgrid_generator.py:84,106-129

Input contract (as produced by the reference's collate, data/SROIE_dataset.py:184-197):
``mask[b]`` is a prefix mask with ``seg_indices[b].shape[0]`` ones.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Sequence, Tuple

import numpy as np

WINDOW = 510


def resize_geometry(h: int, w: int, min_size: float, max_size: float) -> Tuple[int, int]:
    lo, hi = float(min(h, w)), float(max(h, w))
    scale = float(min_size) / lo
    if hi * scale > float(max_size):
        scale = float(max_size) / hi
    return int(math.floor(float(h) * scale)), int(math.floor(float(w) * scale))


@dataclass
class BatchPlan:
    B: int
    H: int
    W: int
    sizes: List[Tuple[int, int]]
    K: int
    n_tok: int
    L: int
    nseq: int
    R: int
    max_len: int
    seg_counts: List[int]
    table: np.ndarray = field(repr=False)          # packed int32 upload
    offsets: dict = field(default_factory=dict)    # name -> (start, length) inside table

    def view(self, name) -> np.ndarray:
        s, n = self.offsets[name]
        return self.table[s:s + n]


def plan_batch(image_shapes: Sequence[Tuple[int, int]], n_toks: Sequence[int], n_segs: Sequence[int], L: int,
               min_size: float, max_size: float, size_divisible: int = 32) -> BatchPlan:
    B = len(image_shapes)
    assert B == len(n_toks) == len(n_segs) and B > 0
    # ``min_size``: one value (eval: test_image_min_size) or one per image (training: transform.py:192-194 draws per image)
    mins = list(min_size) if isinstance(min_size, (list, tuple)) else [min_size] * B
    assert len(mins) == B
    sizes = [resize_geometry(h, w, m, max_size) for (h, w), m in zip(image_shapes, mins)]
    H = int(math.ceil(float(max(s[0] for s in sizes)) / size_divisible) * size_divisible)
    W = int(math.ceil(float(max(s[1] for s in sizes)) / size_divisible) * size_divisible)
    ratios = np.asarray([[oh / h, ow / w] for (h, w), (oh, ow) in zip(image_shapes, sizes)], dtype=np.float32)

    seg_off = np.zeros(B + 1, np.int32)
    seg_off[1:] = np.cumsum(n_segs)
    tok_off = np.zeros(B + 1, np.int32)
    tok_off[1:] = np.cumsum(n_toks)
    K, n_tok = int(seg_off[-1]), int(tok_off[-1])

    n_win = L // WINDOW + 1
    seq_tab, cu, seq_of = [], [0], {}
    for w in range(n_win):
        col0 = w * WINDOW
        len_w = min(WINDOW, L - col0)               # width of this window's corpus slice
        if len_w <= 0:
            continue                                # "[CLS][SEP]+pad" window: output discarded
        for b in range(B):
            n = min(max(int(n_toks[b]) - col0, 0), len_w)
            if n == 0:
                continue                            # no real token -> nothing of this window is ever read
            seq_of[(b, w)] = len(seq_tab)
            seq_tab.append((b, col0, n, len_w + 1))
            cu.append(cu[-1] + n + 2)
    nseq, R = len(seq_tab), cu[-1]
    max_len = max((n + 2 for _, _, n, _ in seq_tab), default=0)
    tok_row = np.zeros(n_tok, np.int32)
    for b in range(B):
        t = np.arange(int(n_toks[b]))
        if t.size == 0:
            continue
        base = np.asarray([cu[seq_of[(b, int(w))]] for w in range((int(n_toks[b]) - 1) // WINDOW + 1)], np.int64)
        tok_row[tok_off[b]:tok_off[b + 1]] = base[t // WINDOW] + 1 + t % WINDOW

    parts = {
        "seg_off": seg_off, "tok_off": tok_off, "ratios": ratios.reshape(-1).view(np.int32),
        "seq_tab": np.asarray(seq_tab, np.int32).reshape(-1), "cu": np.asarray(cu, np.int32), "tok_row": tok_row,
        "all_off": np.asarray([0, K], np.int32),    # the batch as ONE sequence (CRF head's inference(), see engine.py)
    }
    offsets, chunks, pos = {}, [], 0
    for name, arr in parts.items():
        arr = np.ascontiguousarray(arr, dtype=np.int32)
        pad = (-arr.size) % 4                       # keep every sub-table 16-byte aligned
        offsets[name] = (pos, arr.size)
        chunks.append(arr)
        if pad:
            chunks.append(np.zeros(pad, np.int32))
        pos += arr.size + pad
    table = np.concatenate(chunks) if chunks else np.zeros(0, np.int32)
    return BatchPlan(B, H, W, sizes, K, n_tok, L, nseq, R, max_len, [int(s) for s in n_segs], table, offsets)


def roberta_position_ids(ids, pos, padding_idx, cu=None):
    """HF ``create_position_ids_from_input_ids`` (RobertaEmbeddings) over the packed rows: inside each sequence a non-pad id
    gets ``padding_idx + (number of non-pad ids up to and including it)``, a pad id gets ``padding_idx``.

    ``cu`` [nseq + 1] is the packed-row offset table (``cu_seqlens``): a row's sequence start is looked up there, never
    derived from ``pos`` -- ``pos`` of the [SEP] row is its slot in the PADDED window (``len_w + 1``, after the dataset's
    0-padding, model/BERTgrid_generator.py:106-129), not its packed index.  The dataset's 0 pads between the last real token
    and [SEP] are not ``padding_idx`` (RoBERTa's <pad> is 1), so HF counts them: the [SEP] row adds ``pos - packed_index``.
    Without ``cu`` (unit tests over plain back-to-back sequences) ``pos`` must be the packed index itself.
    Integer torch ops on the [R] id vector (capturable in a CUDA graph); the result feeds the same embedding kernel."""
    import torch
    R = ids.numel()
    rows = torch.arange(R, device=ids.device, dtype=torch.int64)
    nz = (ids != padding_idx).to(torch.int32)
    c = torch.cumsum(nz, 0, dtype=torch.int32)
    if cu is None:
        first = rows - pos.long()
    else:
        seq = torch.searchsorted(cu.long().contiguous(), rows, right=True) - 1
        first = cu.long()[seq]
    before = c[first] - nz[first]
    skipped = (pos.long() - (rows - first)).to(torch.int32)        # dataset pads the packed layout dropped before this row
    return ((c - before + skipped) * nz + padding_idx).to(pos.dtype).contiguous()
