#!/usr/bin/env python
"""ncu target: two launches of the pre-split GEMM -- an epilogue-dominated shape (K = 64) and the BERT FFN-up shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vibertgrid_pytorch_b200 import ops
dev = "cuda"
for (M, N, K) in [(131072, 256, 64), (4128, 3072, 768)]:
    A = ops.to_split(torch.randn(M, K, device=dev)); W = torch.randn(N, K, device=dev) * 0.02
    Ws = ops.split_bf16(W)
    ep = ops.make_epilogue(None, torch.zeros(N, device=dev))
    for so in (True, False):
        for _ in range(2):
            ops.gemm(A, W, ep=ep, precision=ops.PREC_BF16X3, W_split=Ws, split_out=so)
    torch.cuda.synchronize()
