#!/usr/bin/env python
"""Tuning experiment: time the BERT-shaped bf16x3 GEMMs for the N-tile width forced by VBG_TC3_BN (one process per value)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vibertgrid_pytorch_b200 import ops

def timed(fn, reps=20):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

dev = "cuda"
for (M, N, K) in [(4128, 2304, 768), (4128, 768, 768), (4128, 3072, 768), (4128, 768, 3072), (131072, 256, 2304), (131072, 64, 576), (32768, 128, 1152), (8192, 256, 2304)]:
    A = torch.randn(M, K, device=dev); W = torch.randn(N, K, device=dev) * 0.02
    Ws = ops.split_bf16(W); out = torch.empty(M, N, device=dev)
    ep = ops.make_epilogue(None, torch.zeros(N, device=dev))
    ms = timed(lambda: ops.gemm(A, W, ep=ep, precision=ops.PREC_BF16X3, out=out, W_split=Ws))
    print(f"BN={os.environ.get('VBG_TC3_BN','auto'):>4} [{M}x{N}x{K}] {ms*1e3:8.1f} us {2.0*M*N*K/ms/1e9:7.1f} TF/s", flush=True)
