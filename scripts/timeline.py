#!/usr/bin/env python
"""Tuning aid: pipeline milestones (clock64 of CTA 0) of one CTA-pair GEMM launch per shape -- where the fixed cost of a
small GEMM goes (setup, first-operand latency, mainloop, epilogue, teardown)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vibertgrid_pytorch_b200 import ops, _lib
ops.TUNE = ops.TUNE_PAIRS_ON
dev = "cuda"
buf = torch.zeros(16, dtype=torch.int64, device=dev)
names = ["entry", "setup_done", "after_pdl_wait", "first_tma_issued", "first_operands_landed", "last_mma_committed(tile0)",
         "epi_start(tile0)", "epi_end(tile0)", "before_final_sync", "exit"]
for (M, N, K, act) in [(256, 768, 768, 0), (4128, 768, 768, 0), (4128, 2304, 768, 0), (4128, 3072, 768, 0), (4128, 3072, 768, ops.ACT_GELU),
                       (4128, 768, 3072, 0)]:
    A = ops.to_split(torch.randn(M, K, device=dev)); W = torch.randn(N, K, device=dev) * 0.02
    Ws = ops.split_bf16(W); ep = ops.make_epilogue(None, torch.zeros(N, device=dev), act=act)
    f = lambda: ops.gemm(A, W, ep=ep, precision=ops.PREC_BF16X3, W_split=Ws, split_out=True)
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): f()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    _lib.load().vbg_debug_set_timeline(buf.data_ptr())
    f(); torch.cuda.synchronize()
    _lib.load().vbg_debug_set_timeline(None)
    t = buf.cpu().tolist()
    print(f"[{M}x{N}x{K} act={act}] back-to-back {us:.1f} us/launch; CTA 0 milestones (cycles since entry, us at 1.9 GHz):")
    for i, n in enumerate(names):
        d = t[i] - t[0]
        print(f"    {n:28s} {d:9d}  {d/1900.0:7.2f} us")

# ---- attention: milestones of CTA (q-tile 0, head 0, sequence 0) at the cfg2 BERT shape (8 x (512 + 4) packed rows)
import numpy as np
lens = [512, 4] * 8
cu = np.zeros(len(lens) + 1, np.int32); cu[1:] = np.cumsum(lens)
R = int(cu[-1])
qkv = ops.to_split(torch.randn(R, 2304, device=dev))
cud = torch.from_numpy(cu).to(dev)
f = lambda: ops.attention_split(qkv, cud, len(lens), 512, 12, split_out=True)
for _ in range(3): f()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): f()
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) / 20 * 1e3
_lib.load().vbg_debug_set_timeline(buf.data_ptr())
f(); torch.cuda.synchronize()
_lib.load().vbg_debug_set_timeline(None)
t = buf.cpu().tolist()
an = ["entry", "producer_start(setup done)", "Q landed (MMA)", "all S MMAs issued", "all PV MMAs issued", "S complete (softmax starts)",
      "row max done", "all P written", "O complete", "output stored", "exit"]
print(f"[attention 8x(512+4) rows, 12 heads] back-to-back {us:.1f} us/launch; CTA(0,0,0) milestones:")
for i, n in enumerate(an):
    d = t[i] - t[0]
    print(f"    {n:30s} {d:9d}  {d/1900.0:7.2f} us")
