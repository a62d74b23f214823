#!/usr/bin/env python
"""Tuning experiment: pre-split bf16x3 GEMM / conv (TMA-fed planes) vs the fp32-A converter kernel, per ring-stage
variant (ops.TUNE: 64-element K stages / SWIZZLE_128B, single-CTA or CTA-pair tiles; TUNE_KB32: 32 / SWIZZLE_64B)."""
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
GEMMS = [(8256, 2304, 768), (8256, 3072, 768), (4128, 2304, 768), (4128, 768, 768), (4128, 3072, 768), (4128, 768, 3072), (33024, 3072, 768), (131072, 256, 256),
         (131072, 256, 64), (32768, 256, 128), (1024, 1024, 12544)]
for (M, N, K) in GEMMS:
    A = torch.randn(M, K, device=dev); W = torch.randn(N, K, device=dev) * 0.02
    Ws = ops.split_bf16(W); As = ops.to_split(A)
    ep = ops.make_epilogue(None, torch.zeros(N, device=dev))
    ms = timed(lambda: ops.gemm(A, W, ep=ep, precision=ops.PREC_BF16X3, W_split=Ws))
    line = f"[gemm {M}x{N}x{K}] fp32-A {ms*1e3:7.1f} us {2.0*M*N*K/ms/1e9:6.1f} TF/s |"
    for kb, cg in (("64", "0"), ("64", "1"), ("32", "0")):
        ops.TUNE = (ops.TUNE_KB32 if kb == "32" else 0) | (ops.TUNE_PAIRS_ON if cg == "1" else ops.TUNE_PAIRS_OFF)
        ms = timed(lambda: ops.gemm(As, W, ep=ep, precision=ops.PREC_BF16X3, W_split=Ws, split_out=True))
        line += f" ps{kb}{'x2' if cg == '1' else ''} {ms*1e3:7.1f} us {2.0*M*N*K/ms/1e9:6.1f} TF/s |"
    print(line, flush=True)

CONVS = [(8, 128, 128, 256, 256, 3, 1), (8, 128, 128, 64, 64, 3, 1), (8, 64, 64, 128, 128, 3, 1), (8, 32, 32, 256, 256, 3, 1),
         (8, 16, 16, 512, 512, 3, 1), (1024, 7, 7, 256, 256, 3, 1), (8, 128, 128, 64, 128, 3, 2)]
for (B, H, Wd, Cin, Cout, k, s) in CONVS:
    x = torch.randn(B, H, Wd, Cin, device=dev); w = torch.randn(Cout, k, k, Cin, device=dev) * 0.02
    ws = ops.split_bf16(w); xs = ops.to_split(x)
    ep = ops.make_epilogue(torch.ones(Cout, device=dev), torch.zeros(Cout, device=dev), act=ops.ACT_RELU)
    Ho, Wo = (H + 2 - k) // s + 1, (Wd + 2 - k) // s + 1
    fl = 2.0 * B * Ho * Wo * Cout * Cin * k * k
    ms = timed(lambda: ops.conv2d(x, w, s, 1, ep=ep, precision=ops.PREC_BF16X3, W_split=ws))
    line = f"[conv B{B} {H}x{Wd} {Cin}->{Cout} k{k} s{s}] fp32-A {ms*1e3:7.1f} us {fl/ms/1e9:6.1f} TF/s |"
    for kb, cg in (("64", "0"), ("64", "1"), ("32", "0")):
        ops.TUNE = (ops.TUNE_KB32 if kb == "32" else 0) | (ops.TUNE_PAIRS_ON if cg == "1" else ops.TUNE_PAIRS_OFF)
        ms = timed(lambda: ops.conv2d(xs, w, s, 1, ep=ep, precision=ops.PREC_BF16X3, W_split=ws, split_out=True))
        line += f" ps{kb}{'x2' if cg == '1' else ''} {ms*1e3:7.1f} us {fl/ms/1e9:6.1f} TF/s |"
    print(line, flush=True)
