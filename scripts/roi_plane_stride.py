"""Does the distance between the hi and lo planes matter to the streaming ROI-align kernel?  At cfg2 the planes of P_fuse are
exactly 2^26 bytes apart; both are streamed at once.  Times the kernel with the lo plane displaced by a pad."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vibertgrid_pytorch_b200 import ops, synth, _lib as L

cfg = synth.CONFIGS["cfg2"]
dev = torch.device("cuda")
B, S = cfg.batch, cfg.segments
K = B * S
g = torch.Generator().manual_seed(1)
boxes = torch.cat([synth.make_boxes(S, cfg.height, cfg.width, g) for _ in range(B)], 0).int().to(dev)
seg_off = torch.arange(0, K + 1, S, dtype=torch.int32, device=dev)
Hf, Wf = cfg.height // 4, cfg.width // 4
feat = ops.to_split(torch.randn(B, Hf, Wf, 256, device=dev))
numel = feat.plane
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
by = B * 256 * Hf * Wf * 4 + K * 256 * 49 * 4 + K * 20
lib = L.load()
for pad in (0, 8, 1024, 4096 + 512, 65536 + 1024, (1 << 20) + 4096):
    buf = torch.empty(2 * numel + pad, dtype=torch.bfloat16, device=dev)
    buf[:numel].copy_(feat.t[0].reshape(-1)); buf[numel + pad:].copy_(feat.t[1].reshape(-1))
    for opad in (0, 4096 + 512):
        onum = K * 49 * 256
        obuf = torch.empty(2 * onum + opad, dtype=torch.bfloat16, device=dev)
        ts = []
        for rep in range(12):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = lib.vbg_roi_align_sel(buf.data_ptr(), numel + pad, B, Hf, Wf, 256, boxes.data_ptr(), seg_off.data_ptr(), K, 0.25, 7,
                                       obuf.data_ptr(), onum + opad, None, ops.ROI_STREAM, torch.cuda.current_stream().cuda_stream)
            e1.record(); torch.cuda.synchronize()
            assert rc == 0
            ts.append(e0.elapsed_time(e1) * 1e3)
        ts = ts[2:]
        print(f"in-plane pad {pad * 2:8d} B, out-plane pad {opad * 2:6d} B: " + " ".join(f"{t:5.1f}" for t in ts) +
              f" | avg {sum(ts) / len(ts):5.1f} us -> {by / (sum(ts) / len(ts)) / 1e3 / 6548.8:.3f} of HBM peak", flush=True)
