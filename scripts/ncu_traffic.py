#!/usr/bin/env python
"""ncu target for profiles/r1_traffic.json: one launch each of the three roofline kernels at cfg2 shapes, in the storage
formats the engine uses (pre-split planes).  Run under `ncu --set full`; scripts/extract_traffic.py reads the report."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vibertgrid_pytorch_b200 import ops, synth
dev = "cuda"
cfg = synth.CONFIGS["cfg2"]
B, S = cfg.batch, cfg.segments
K = B * S
g = torch.Generator().manual_seed(1)
boxes = torch.cat([synth.make_boxes(S, cfg.height, cfg.width, g) for _ in range(B)], 0).int().to(dev)
seg_off = torch.arange(0, K + 1, S, dtype=torch.int32, device=dev)
emb = ops.to_split(torch.randn(K, 768, device=dev))
idx = ops.box_index_map(boxes, seg_off, B, 8, cfg.height // 8, cfg.width // 8)
feat = ops.to_split(torch.randn(B, cfg.height // 4, cfg.width // 4, 256, device=dev))
M = B * (cfg.seq_len + 2 * (cfg.seq_len // 510 + 1))
A = ops.to_split(torch.randn(M, 768, device=dev)); W = torch.randn(3072, 768, device=dev) * 0.03
Ws = ops.split_bf16(W); ep = ops.make_epilogue(None, torch.zeros(3072, device=dev), act=ops.ACT_GELU)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(2):
    flush.zero_(); ops.grid_scatter(emb, idx, seg_off)
    flush.zero_(); ops.roi_align(feat, boxes, seg_off, 0.25, 7, split_out=True)
    flush.zero_(); ops.gemm(A, W, ep=ep, precision=ops.PREC_BF16X3, W_split=Ws, split_out=True)
torch.cuda.synchronize()
