"""The per-sequence CRF negative log-likelihood + gradient source of the CUDA kernels (csrc/vbg_crf_seq.h), compiled for the
host with g++ and held to (1) the oracle's float64 restatement and (2), in the build container, the UNMODIFIED reference's
``model.crf.CRF`` forward + autograd backward.  No GPU needed: the kernels include the very same header."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import oracle_ops

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


@pytest.fixture(scope="module")
def host_lib(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("crf") / "crf_host.so")
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-I", os.path.join(ROOT, "vibertgrid-pytorch_b200", "csrc"),
                    os.path.join(ROOT, "tests", "crf_host.cpp"), "-o", so], check=True)
    return ctypes.CDLL(so)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def host_crf(lib, feats, trans, tags, seg_off, dnll):
    K, T = feats.shape
    B = len(seg_off) - 1
    alpha = np.zeros((K, T), np.float32)
    logz, nll = np.zeros(B, np.float32), np.zeros(B, np.float32)
    lib.crf_host_nll_fwd(_p(feats), _p(trans), _p(tags), _p(seg_off), B, T, _p(alpha), _p(logz), _p(nll))
    dfeats = np.zeros((K, T), np.float32)
    dtr = np.zeros((B, T, T), np.float32)
    lib.crf_host_nll_bwd(_p(feats), _p(trans), _p(tags), _p(seg_off), B, T, _p(alpha), _p(dnll), _p(dfeats), _p(dtr))
    return nll, dfeats, dtr.sum(0)


def make_case(seed, lens, C):
    """Emissions / transitions / gold tags shaped like the `crf` head's (T = C + 2; START/STOP rows pinned at -10000)."""
    g = torch.Generator().manual_seed(seed)
    T = C + 2
    K = sum(lens)
    feats = torch.randn(K, T, generator=g) * 1.5
    trans = torch.randn(T, T, generator=g)
    trans[T - 2, :] = -10000.0
    trans[:, T - 1] = -10000.0
    tags = torch.randint(0, C, (K,), generator=g).to(torch.int32)
    seg_off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    return feats, trans, tags, seg_off


CASES = [(0, [7, 5], 4), (1, [64, 64], 4), (2, [1, 2, 130], 12), (3, [200], 30), (4, [65, 63, 128], 6), (5, [1024, 700], 12)]


@pytest.mark.parametrize("seed,lens,C", CASES)
def test_host_compiled_kernel_source_matches_oracle(host_lib, seed, lens, C):
    feats, trans, tags, seg_off = make_case(seed, lens, C)
    T = C + 2
    B = len(lens)
    dnll = (np.arange(B, dtype=np.float32) + 1.0) / B
    nll, dfeats, dtrans = host_crf(host_lib, feats.numpy().copy(), trans.numpy().copy(), tags.numpy().copy(), seg_off, dnll)
    f = feats.clone().double().requires_grad_(True)
    tr = trans.clone().double().requires_grad_(True)
    want = torch.stack([oracle_ops.crf_nll_torch(f[seg_off[b]:seg_off[b + 1]], tags[seg_off[b]:seg_off[b + 1]], tr, T - 2, T - 1)
                        for b in range(B)])
    (want * torch.from_numpy(dnll).double()).sum().backward()
    for b in range(B):      # the numpy oracle of the eval suite agrees with the differentiable one
        ref = oracle_ops.crf_nll(feats[seg_off[b]:seg_off[b + 1]].numpy(), tags[seg_off[b]:seg_off[b + 1]].numpy(),
                                 trans.numpy(), T - 2, T - 1)
        assert abs(ref - float(want[b])) <= 1e-9 * max(1.0, abs(ref))
    assert np.abs(nll - want.detach().numpy()).max() <= 2e-5 * max(1.0, float(want.abs().max()))
    assert np.abs(dfeats - f.grad.numpy()).max() <= 2e-5
    assert np.abs(dtrans - tr.grad.numpy()).max() <= 2e-5 * max(1.0, float(tr.grad.abs().max()))


@pytest.mark.skipif(not os.path.isdir(REF), reason="live reference only exists in the build container")
def test_host_compiled_kernel_source_matches_live_reference(host_lib):
    sys.path.insert(0, REF)
    try:
        for m in [m for m in sys.modules if m.split(".")[0] == "model"]:
            del sys.modules[m]
        from model.crf import CRF, START_TAG, STOP_TAG
    finally:
        sys.path.remove(REF)
    C, lens = 4, [9, 6, 11]
    feats, trans, tags, seg_off = make_case(7, lens, C)
    T = C + 2
    t2i = {f"c{i}": i for i in range(C)}
    t2i[START_TAG], t2i[STOP_TAG] = C, C + 1
    crf = CRF(t2i)
    with torch.no_grad():
        crf.transitions.copy_(trans)
    f = feats.clone().requires_grad_(True)
    # model/field_type_classification_head.py:686-699: score += crf(feat, tag) per sample, / batch size
    score = torch.zeros(1)
    for b in range(len(lens)):
        score = score + crf(feats=f[seg_off[b]:seg_off[b + 1]], tags=tags[seg_off[b]:seg_off[b + 1]].long())
    loss = score / len(lens)
    loss.backward()
    dnll = np.full(len(lens), 1.0 / len(lens), np.float32)
    nll, dfeats, dtrans = host_crf(host_lib, feats.numpy().copy(), trans.numpy().copy(), tags.numpy().copy(), seg_off, dnll)
    assert abs(float(nll.sum()) / len(lens) - float(loss)) <= 1e-5 * max(1.0, abs(float(loss)))
    assert np.abs(dfeats - f.grad.numpy()).max() <= 1e-5
    assert np.abs(dtrans - crf.transitions.grad.numpy()).max() <= 1e-5 * max(1.0, float(crf.transitions.grad.abs().max()))
