"""ROI-align kernels side by side at a BASELINE shape: the 64-channel windowed kernel (default) against the row-per-warp kernel
(VBG_ROI_ROW=1 / 2) and its persistent double-buffered forms (3 / 4): bit-equality of the outputs in both storage formats, then CUDA-event timings with an L2 flush between launches
(the same recipe as bench.py's roofline_hbm_kernels)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vibertgrid_pytorch_b200 import ops, synth

cfg = synth.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "cfg2"]
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


NAMES = {0: "windowed-64     ", 1: "row-per-warp 128", 2: "row-per-warp 64 ", 3: "persistent 64   ", 4: "persistent 128  "}


def run(row, split):
    os.environ["VBG_ROI_ROW"] = str(int(row))
    f = feat_s if split else feat32
    return ops.roi_align(f, boxes, seg_off, 0.25, 7, want_grid=True, split_out=split)


def timed(row, split, reps=10):
    run(row, split)
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(row, split); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sum(ts) / len(ts), min(ts)


for split in (True, False):
    a, ga = run(0, split)
    ta = a.t if split else a
    for v in (1, 2, 3, 4):
        b, gb = run(v, split)
        tb = b.t if split else b
        same = torch.equal(ta, tb) and torch.equal(ga, gb)
        md = float((ta.float() - tb.float()).abs().max())
        print(f"[{cfg.name} planes={split}] {NAMES[v]} == windowed-64: {same} (max |diff| {md:.3e}), finite: {bool(torch.isfinite(tb.float()).all())}")
    for row in (0, 1, 2, 3, 4):
        ms, best = timed(row, split)
        print(f"[{cfg.name} planes={split}] {NAMES[row]} {ms * 1e3:7.1f} us avg, {best * 1e3:7.1f} us best"
              f" -> {by / ms / 1e6:7.0f} GB/s = {by / ms / 1e6 / 6548.8:.3f} of measured HBM peak")
