#!/usr/bin/env python
"""Tuning aid: clock64 milestones of one CTA (block 2000) of the windowed ROI-align kernel at the cfg2 shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vibertgrid_pytorch_b200 import ops, synth, _lib
dev = "cuda"
cfg = synth.CONFIGS["cfg2"]
B, S = cfg.batch, cfg.segments
K = B * S
g = torch.Generator().manual_seed(1)
boxes = torch.cat([synth.make_boxes(S, cfg.height, cfg.width, g) for _ in range(B)], 0).int().to(dev)
seg_off = torch.arange(0, K + 1, S, dtype=torch.int32, device=dev)
feat = ops.to_split(torch.randn(B, cfg.height // 4, cfg.width // 4, 256, device=dev))
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
buf = torch.zeros(16, dtype=torch.int64, device=dev)
f = lambda: ops.roi_align(feat, boxes, seg_off, 0.25, 7, split_out=True)
f(); torch.cuda.synchronize()
_lib.load().vbg_debug_set_timeline(buf.data_ptr())
flush.zero_(); f(); torch.cuda.synchronize()
_lib.load().vbg_debug_set_timeline(None)
t = buf.cpu().tolist()
names = ["entry", "geometry done", "tables built", "cp.async issued", "window landed", "merged + synced", "bins stored"]
bx = boxes[2000 // 4].tolist()
print(f"[roi_align cfg2] CTA 2000 (roi {2000 // 4}, box {bx}) milestones:")
for i, n in enumerate(names):
    d = t[i] - t[0]
    print(f"    {n:20s} {d:9d}  {d/1900.0:7.2f} us")
