"""Tests of kernel variants that were written after a round's GPU budget was spent: compiled, never launched.  They are part
of the suite only when VBG_TEST_UNVERIFIED=1, so an unproven variant cannot fail the regular `-m gpu` run; the first GPU job of
the next round runs them (`VBG_TEST_UNVERIFIED=1 python -m pytest tests/test_gpu_unverified_variants.py -m gpu`) and, once
green, each case moves into the regular per-kernel tests."""
import os

import numpy as np
import pytest
import torch

from oracle import oracle_ops

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("VBG_TEST_UNVERIFIED") != "1", reason="unproven kernel variants: opt-in (see the module docstring)")]


def _case(seed, B, Hf, Wf, C, counts):
    rng = np.random.default_rng(seed)
    per = []
    for c in counts:
        l = rng.integers(0, Wf * 4 - 2, c); t = rng.integers(0, Hf * 4 - 2, c)
        r = np.minimum(l + rng.integers(1, Wf * 2, c), Wf * 4 - 1); b = np.minimum(t + rng.integers(1, Hf, c), Hf * 4 - 1)
        per.append(np.stack([l, t, r, b], 1).astype(np.int32))
    per[0][0] = [0, 0, Wf * 4 - 1, Hf * 4 - 1]          # full page: window too large -> the in-kernel global-tap path
    per[0][1] = [10, 10, 10, 10]                        # zero size -> clamped to 1
    per[-1][0] = [Wf * 4 - 3, Hf * 4 - 3, Wf * 4 + 30, Hf * 4 + 30]   # runs off the map
    off = np.zeros(len(counts) + 1, np.int32); off[1:] = np.cumsum(counts)
    feat = torch.randn(B, C, Hf, Wf, generator=torch.Generator().manual_seed(seed))
    return feat, np.concatenate(per, 0), off


@pytest.mark.parametrize("variant", ["3", "4"])
@pytest.mark.parametrize("split", [False, True])
def test_roi_align_persistent_variants(monkeypatch, variant, split):
    """roi_align_pipe_kernel (VBG_ROI_ROW=3: 64-channel chunks, =4: 128-channel chunks) against the oracle and, bit for bit,
    against the default row-per-warp kernel; more items than persistent CTAs so every CTA loops and both buffers rotate."""
    from vibertgrid_pytorch_b200 import ops
    B, Hf, Wf, C = 3, 48, 64, 256
    counts = [170, 150, 161]
    feat, boxes, off = _case(7, B, Hf, Wf, C, counts)
    doff = torch.from_numpy(off).cuda()
    x = feat.permute(0, 2, 3, 1).contiguous().cuda()
    src = ops.to_split(x) if split else x
    monkeypatch.setenv("VBG_ROI_ROW", "1")
    ref, gref = ops.roi_align(src, torch.from_numpy(boxes).cuda(), doff, 0.25, 7, want_grid=True, split_out=split)
    monkeypatch.setenv("VBG_ROI_ROW", variant)
    got, ggot = ops.roi_align(src, torch.from_numpy(boxes).cuda(), doff, 0.25, 7, want_grid=True, split_out=split)
    torch.cuda.synchronize()
    assert torch.equal(ggot, gref)
    assert torch.equal(got.t if split else got, ref.t if split else ref)
    bidx = np.concatenate([np.full(c, b, np.int32) for b, c in enumerate(counts)])
    want, grids = oracle_ops.roi_align(feat.numpy(), boxes.astype(np.float32), bidx, 0.25, 7)
    assert np.array_equal(ggot.cpu().numpy(), grids)
    out = (got.float() if split else got).permute(0, 3, 1, 2).cpu().numpy()
    err = np.abs(out - want).max() / max(np.abs(want).max(), 1e-30)
    assert err < (2 ** -14 if split else 1e-5)
