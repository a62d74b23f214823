#!/usr/bin/env python
"""GPU probe of the tcgen05 (VBG_PREC_TF32) GEMM / implicit-GEMM conv path: correctness against
torch fp32 on CPU-free references, what the tensor core does with the low 13 mantissa bits, and
per-shape throughput.  Diagnostic tool (scripts/), not part of the product path.

    python scripts/tc_probe.py [--quick]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
import torch.nn.functional as F

from vibertgrid_pytorch_b200 import ops, _lib


def trunc_tf32(x):
    return (x.view(torch.int32) & ~0x1FFF).view(torch.float32)


def rn_tf32(x):
    i = x.view(torch.int32)
    return ((i + 0x0FFF + ((i >> 13) & 1)) & ~0x1FFF).view(torch.float32)


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def timed(fn, reps=20):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    quick = "--quick" in sys.argv
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    ok = ops.tc_available()
    print("tc_available:", ok, "|", _lib.last_error())
    if not ok:
        return 1
    dev = "cuda"
    g = torch.Generator(device="cpu").manual_seed(0)

    # 1. structural check: A = row id pattern, W = identity-like -> C must reproduce A columns exactly
    M, N, K = 128, 64, 32
    A = (torch.arange(M)[:, None] * 64 + torch.arange(K)[None, :]).float()
    Wm = torch.zeros(N, K); Wm[torch.arange(K), torch.arange(K)] = 1.0
    C = ops.gemm(A.to(dev), Wm.to(dev), precision=ops.PREC_TF32)
    torch.cuda.synchronize()
    want = A @ Wm.t()
    bad = (C.cpu() != want).nonzero()
    print(f"[identity 128x64x32] mismatches: {bad.shape[0]}", "" if bad.shape[0] == 0 else f"first: {bad[:8].tolist()} got {C.cpu()[bad[0,0], :8].tolist()}")

    # 2. numerics: what happens to the low mantissa bits
    for (M, N, K) in [(256, 128, 768), (4128, 3072, 768), (4128, 768, 3072)]:
        A = torch.randn(M, K, generator=g).to(dev); Wm = (torch.randn(N, K, generator=g) / K ** 0.5).to(dev)
        C = ops.gemm(A, Wm, precision=ops.PREC_TF32)
        exact = (A.double() @ Wm.double().t())
        e_exact = rel(C, exact)
        e_trunc = rel(C, trunc_tf32(A).double() @ trunc_tf32(Wm).double().t())
        e_rn = rel(C, rn_tf32(A).double() @ rn_tf32(Wm).double().t())
        C2 = ops.gemm(rn_tf32(A), rn_tf32(Wm), precision=ops.PREC_TF32)
        e_pre = rel(C2, exact)
        Cs = ops.gemm(A, Wm, precision=ops.PREC_FP32)
        print(f"[numerics {M}x{N}x{K}] vs exact {e_exact:.2e} | vs trunc-operands {e_trunc:.2e} | vs rn-operands {e_rn:.2e} | "
              f"pre-rounded(RN) inputs vs exact {e_pre:.2e} | simt fp32 vs exact {rel(Cs, exact):.2e}")

    # 3. epilogue + split-K-source + ragged M/N
    for (M, N, K, K1) in [(300, 768, 768, 768), (4100, 3072, 768, 768), (128, 1024, 1792, 1024), (1000, 136, 64, 64), (77, 64, 96, 32), (4096, 128, 896, 128)]:
        A = torch.randn(M, K, generator=g); Wm = torch.randn(N, K, generator=g) / K ** 0.5
        bias = torch.randn(N, generator=g); res = torch.randn(M, N, generator=g)
        want = F.gelu(trunc_tf32(A).double() @ trunc_tf32(Wm).double().t() + bias + res)
        a1, a2 = A[:, :K1].contiguous().to(dev), (A[:, K1:].contiguous().to(dev) if K1 < K else None)
        ep = ops.make_epilogue(None, bias.to(dev), res.to(dev), ops.RES_SAME, ldr=N, act=ops.ACT_GELU)
        got = ops.gemm(a1, Wm.to(dev), A2=a2, ep=ep, precision=ops.PREC_TF32)
        print(f"[gemm+epi {M}x{N}x{K} K1={K1}] vs trunc-operand ref {rel(got.cpu(), want):.2e}")

    # 4. conv (stride 1: tcgen05; others fall back)
    for (B, H, W, Cin, Cout, k, s, p) in [(2, 32, 48, 64, 64, 3, 1, 1), (8, 128, 128, 64, 64, 3, 1, 1), (8, 64, 64, 128, 128, 3, 1, 1),
                                          (8, 32, 32, 256, 256, 3, 1, 1), (8, 16, 16, 512, 512, 3, 1, 1), (9, 7, 7, 256, 256, 3, 1, 1),
                                          (1024, 7, 7, 256, 256, 3, 1, 1), (2, 20, 12, 64, 128, 3, 1, 1), (2, 16, 16, 128, 256, 3, 2, 1)]:
        x = torch.randn(B, Cin, H, W, generator=g); w = torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5
        scale = torch.rand(Cout, generator=g) + 0.5; shift = torch.randn(Cout, generator=g)
        xd, wd = x.to(dev), w.to(dev)
        want = F.relu(F.conv2d(trunc_tf32(xd).double(), trunc_tf32(wd).double(), None, s, p) * scale.to(dev)[None, :, None, None].double()
                      + shift.to(dev)[None, :, None, None].double())
        w_ohwi = ops.repack_oihw_to_ohwi(wd)
        ep = ops.make_epilogue(scale.to(dev), shift.to(dev), act=ops.ACT_RELU)
        xn = xd.permute(0, 2, 3, 1).contiguous()
        got = ops.conv2d(xn, w_ohwi, s, p, ep=ep, precision=ops.PREC_TF32)
        e = rel(got.permute(0, 3, 1, 2), want)
        ms = timed(lambda: ops.conv2d(xn, w_ohwi, s, p, ep=ep, precision=ops.PREC_TF32), 10)
        Ho, Wo = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
        fl = 2.0 * B * Ho * Wo * Cout * Cin * k * k
        print(f"[conv B{B} {H}x{W} {Cin}->{Cout} k{k}s{s}] err {e:.2e}  {ms*1e3:8.1f} us  {fl/ms/1e9:7.1f} TFLOP/s")
    if quick:
        return 0

    # 5. throughput sweep of the forward's GEMM shapes (cfg2, B=8), L2-warm back-to-back
    shapes = [(4128, 2304, 768), (4128, 768, 768), (4128, 3072, 768), (4128, 768, 3072), (32768, 128, 896), (131072, 256, 64),
              (131072, 256, 256), (1024, 1024, 12544), (1024, 1024, 1792), (1024, 512, 1024), (8192, 8192, 8192)]
    for (M, N, K) in shapes:
        A = torch.randn(M, K, device=dev); Wm = torch.randn(N, K, device=dev) * 0.02
        bias = torch.zeros(N, device=dev)
        ep = ops.make_epilogue(None, bias)
        out = torch.empty(M, N, device=dev)
        Ws = ops.split_bf16(Wm)
        ms = timed(lambda: ops.gemm(A, Wm, ep=ep, precision=ops.PREC_TF32, out=out), 10)
        ms3 = timed(lambda: ops.gemm(A, Wm, ep=ep, precision=ops.PREC_BF16X3, out=out, W_split=Ws), 10)
        e3 = rel(out, A.double() @ Wm.double().t()) if M * N * K < 4e11 else float("nan")
        torch.backends.cuda.matmul.allow_tf32 = True
        ms_lib = timed(lambda: torch.matmul(A, Wm.t()), 10)
        torch.backends.cuda.matmul.allow_tf32 = False
        fl = 2.0 * M * N * K
        print(f"[gemm {M}x{N}x{K}] tf32 {ms*1e3:8.1f} us {fl/ms/1e9:7.1f} TF/s | bf16x3 {ms3*1e3:8.1f} us {fl/ms3/1e9:7.1f} TF/s (err {e3:.1e}) | "
              f"cuBLAS tf32 {ms_lib*1e3:8.1f} us {fl/ms_lib/1e9:7.1f} TF/s")
    # 6. conv shapes of the backbone / heads in bf16x3
    for (B, H, W, Cin, Cout, k, s, p) in [(8, 128, 128, 64, 64, 3, 1, 1), (8, 64, 64, 128, 128, 3, 1, 1), (8, 32, 32, 256, 256, 3, 1, 1),
                                          (8, 16, 16, 512, 512, 3, 1, 1), (8, 128, 128, 256, 256, 3, 1, 1), (1024, 7, 7, 256, 256, 3, 1, 1),
                                          (8, 128, 128, 64, 128, 3, 2, 1), (8, 64, 64, 128, 256, 3, 2, 1), (8, 32, 32, 256, 512, 3, 2, 1)]:
        x = torch.randn(B, H, W, Cin, device=dev); w = torch.randn(Cout, k, k, Cin, device=dev) / (Cin * k * k) ** 0.5
        ws = ops.split_bf16(w)
        ep = ops.make_epilogue(torch.ones(Cout, device=dev), torch.zeros(Cout, device=dev), act=ops.ACT_RELU)
        ms3 = timed(lambda: ops.conv2d(x, w, s, p, ep=ep, precision=ops.PREC_BF16X3, W_split=ws), 10)
        ms1 = timed(lambda: ops.conv2d(x, w, s, p, ep=ep, precision=ops.PREC_TF32), 10)
        Ho, Wo = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
        fl = 2.0 * B * Ho * Wo * Cout * Cin * k * k
        print(f"[conv B{B} {H}x{W} {Cin}->{Cout} k{k}s{s}] bf16x3 {ms3*1e3:8.1f} us {fl/ms3/1e9:7.1f} TF/s | tf32 {ms1*1e3:8.1f} us {fl/ms1/1e9:7.1f} TF/s")
    x4 = torch.randn(8, 518, 518, 4, device=dev); w774, w256 = ops.stem_pack_weights(torch.randn(64, 3, 7, 7, device=dev))
    ws = ops.split_bf16(w256)
    ep = ops.make_epilogue(torch.ones(64, device=dev), torch.zeros(64, device=dev), act=ops.ACT_RELU)
    ms3 = timed(lambda: ops.stem_conv(x4, w774, ep=ep, precision=ops.PREC_BF16X3, W_split=ws), 10)
    ms0 = timed(lambda: ops.stem_conv(x4, w774, ep=ep, precision=ops.PREC_FP32), 5)
    print(f"[stem B8 512x512] bf16x3 {ms3*1e3:8.1f} us | fp32 simt {ms0*1e3:8.1f} us")
    return 0


if __name__ == "__main__":
    sys.exit(main())
