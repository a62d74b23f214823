"""ROI-align kernels side by side at a BASELINE shape: the persistent TMA row-streaming kernel (the product path) against the
row-per-warp kernel (round-1 default) and the direct per-sample kernel -- sample grids bit-equal, values within fp32
re-association -- then CUDA-event timings with an L2 flush between launches (the same recipe as bench.py's
roofline_hbm_kernels).  `python scripts/roi_compare.py cfg2 [ncu]` (with `ncu`: one launch per variant, for a profiler)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vibertgrid_pytorch_b200 import ops, synth

cfg = synth.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "cfg2"]
ncu = len(sys.argv) > 2 and sys.argv[2] == "ncu"
dev = torch.device("cuda")
B, S = cfg.batch, cfg.segments
K = B * S
g = torch.Generator().manual_seed(1)
boxes = torch.cat([synth.make_boxes(S, cfg.height, cfg.width, g) for _ in range(B)], 0).int().to(dev)
seg_off = torch.arange(0, K + 1, S, dtype=torch.int32, device=dev)
Hf, Wf = cfg.height // 4, cfg.width // 4
feat32 = torch.randn(B, Hf, Wf, 256, device=dev)
feat_s = ops.to_split(feat32)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
by = B * 256 * Hf * Wf * 4 + K * 256 * 49 * 4 + K * 20
NAMES = {ops.ROI_STREAM: "stream (TMA rows)", ops.ROI_ROW: "row-per-warp 128 ", ops.ROI_DIRECT: "direct per-sample"}


def run(variant, split):
    return ops.roi_align(feat_s if split else feat32, boxes, seg_off, 0.25, 7, want_grid=True, split_out=split, variant=variant)


def timed(variant, split, reps=10):
    run(variant, split)
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(variant, split); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sum(ts) / len(ts), min(ts)


for split in (True, False):
    a, ga = run(ops.ROI_ROW, split)
    ta = (a.float() if split else a)
    for v in (ops.ROI_STREAM, ops.ROI_DIRECT):
        b, gb = run(v, split)
        tb = (b.float() if split else b)
        torch.cuda.synchronize()
        md = float((ta - tb).abs().max() / ta.abs().max())
        print(f"[{cfg.name} planes={split}] {NAMES[v]} vs row kernel: grid equal {torch.equal(ga, gb)}, max-rel diff {md:.2e}, "
              f"finite {bool(torch.isfinite(tb).all())}", flush=True)
    if ncu:
        continue
    for v in (ops.ROI_STREAM, ops.ROI_ROW, ops.ROI_DIRECT):
        ms, best = timed(v, split)
        print(f"[{cfg.name} planes={split}] {NAMES[v]} {ms * 1e3:7.1f} us avg, {best * 1e3:7.1f} us best"
              f" -> {by / ms / 1e6:7.0f} GB/s = {by / ms / 1e6 / 6548.8:.3f} of measured HBM peak", flush=True)
