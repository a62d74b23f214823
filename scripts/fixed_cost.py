#!/usr/bin/env python
"""Tuning experiment: fixed per-launch cost of the persistent bf16x3 GEMM (M = 128 -> one tile per n-tile) vs M scaling."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vibertgrid_pytorch_b200 import ops

def timed(fn, reps=30):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

dev = "cuda"
for (N, K) in [(3072, 768), (768, 3072), (768, 768)]:
    W = torch.randn(N, K, device=dev) * 0.02; Ws = ops.split_bf16(W)
    ep = ops.make_epilogue(None, torch.zeros(N, device=dev))
    for M in (128, 1024, 2048, 4128, 8256, 16512, 33024):
        A = torch.randn(M, K, device=dev); out = torch.empty(M, N, device=dev)
        ms = timed(lambda: ops.gemm(A, W, ep=ep, precision=ops.PREC_BF16X3, out=out, W_split=Ws))
        print(f"[M={M:6d} N={N} K={K}] {ms*1e3:8.1f} us {2.0*M*N*K/ms/1e9:7.1f} TF/s", flush=True)
x = torch.zeros(1, device=dev)
print(f"[torch tiny kernel back-to-back] {timed(lambda: x.add_(1), 200)*1e3:.2f} us")
