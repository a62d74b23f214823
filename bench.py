#!/usr/bin/env python
"""Headline benchmark: doc-images/s of the ViBERTgrid joint forward (BASELINE.json metric) on
N B200s, one process per GPU, documents sharded across ranks (no data-path collective).

    python bench.py --gpus 1 --steps 20 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # CPU arm: the UNMODIFIED reference (staged under oracle/_ref) on the host cores
    python bench.py --mode train ...          # the line's value is the training step (forward + backward + all-reduce + optimizers)

Prints ONE JSON line on rank 0 (contract in the task statement / DESIGN.md section 6).  Objects of the line:
  value / ms_per_step   eval-mode joint forward, inputs resident in HBM, CUDA events, max over ranks
  e2e                   the same through DevicePrefetcher -> net() -> HostResultQueue from pinned host batches (H2D + D2H timed)
  input_pipeline        e2e fed by the shard reader (ShardLoader: native collate, one uint8 H2D per step) + host-only collate rate
  train_step            whole-step CUDA graph + arena all-reduce (N > 1) + fused SGD / AdamW
  roofline              FFN-up GEMM: operands beyond L2 (primary), flush per launch, L2-warm; scatter / ROI-align in
                        roofline_hbm_kernels (flush per launch + back to back)          -- DESIGN.md section 4
  library_bar           reference eager on this GPU, cuBLAS, torchvision roi_align      (N = 1)
  serving               inference() latency for one cfg1 document, host to host         (N = 1)
  cpu_baseline          the unmodified reference on the host cores, bounded sample      (N = 1)
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

METRIC = "doc_images_per_sec"
UNIT = "images/s"


def fwd_flops_as_executed(cfg) -> float:
    """Forward FLOPs per image AS THE REFERENCE EXECUTES THEM: every 510-token window
    is a full 512-row BERT pass; the fuse / seg-head 1x1 convs run at full resolution."""
    hw = cfg.height * cfg.width / (512.0 * 512.0)
    wn = cfg.seq_len // 510 + 1
    bert = wn * 96.6e9 * (cfg.bert_layers / 12.0)
    backbone = (74.2e9 if "34" in cfg.backbone else 54.9e9) * hw
    seg = 39.7e9 * hw
    late = cfg.segments * 0.145e9
    return bert + backbone + seg + late


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self._stop_evt = index, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        sm = sorted(float(s[0]) for s in self.samples if s and s[0].replace(".", "").isdigit())
        reasons = set()
        for s in self.samples:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        smax = max((float(s[1]) for s in self.samples if len(s) > 1 and s[1].replace(".", "").isdigit()), default=None)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


def build_net(cfg, device=None):
    from vibertgrid_pytorch_b200 import synth
    from vibertgrid_pytorch_b200.net import ViBERTgridNet
    import contextlib
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        try:
            synth.write_bert_dir(cfg, tmp)
            kw = synth.model_kwargs(cfg, "eval")
            with contextlib.redirect_stdout(sys.stderr):      # the constructor prints like the reference's; stdout carries ONE JSON line
                net = ViBERTgridNet(**kw)
        finally:
            os.chdir(cwd)
    synth.fill_state_dict_(net, 0)
    if device is not None:
        net = net.to(device)
    return net.eval(), kw


def to_device(batch, dev, non_blocking=True):
    return [tuple(t.to(dev, non_blocking=non_blocking) for t in x) if isinstance(x, tuple) else x.to(dev, non_blocking=non_blocking)
            for x in batch]


def pin(batch):
    return [tuple(t.pin_memory() for t in x) if isinstance(x, tuple) else x.pin_memory() for x in batch]


def nbytes(batch):
    return sum(sum(t.numel() * t.element_size() for t in x) if isinstance(x, tuple) else x.numel() * x.element_size() for x in batch)


def _staged_reference():
    """The UNMODIFIED reference staged under oracle/_ref/reference (oracle/stage_reference.py), or None."""
    ref = os.path.join(ROOT, "oracle", "_ref", "reference")
    return ref if os.path.isfile(os.path.join(ref, "model", "ViBERTgrid_net.py")) else None


def cpu_reference_arm(cfg, steps, warmup, sample_images=None):
    """The reference's own CPU implementation of the path on all host threads: the unmodified ``ViBERTgridNet`` from the
    staged tree (kind "reference"; eval mode, no_grad, the seeded state dict of the GPU arm loaded with strict=True), or --
    only when the staged tree is absent -- the oracle restatement (kind "port").  One step = ``sample_images`` documents of
    the workload (default: the workload's own batch)."""
    import contextlib
    import dataclasses
    from vibertgrid_pytorch_b200 import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n_img = sample_images or cfg.batch
    c1 = dataclasses.replace(cfg, batch=n_img)
    net, kw = build_net(c1)
    sd = {k: v for k, v in net.state_dict().items()}
    ref_dir = _staged_reference()
    if ref_dir is not None:
        kind = "reference"
        sys.path.insert(0, ref_dir)
        cwd = os.getcwd()
        with tempfile.TemporaryDirectory() as tmp:
            os.chdir(tmp)
            try:
                synth.write_bert_dir(c1, tmp)
                with contextlib.redirect_stdout(sys.stderr):
                    from model.ViBERTgrid_net import ViBERTgridNet as RefNet
                    assert os.path.abspath(sys.modules[RefNet.__module__].__file__).startswith(ref_dir)
                    ref = RefNet(**synth.model_kwargs(c1, "eval"))
            finally:
                os.chdir(cwd)
        ref.load_state_dict(sd, strict=True)
        ref.eval()

        def fwd(batch):
            with torch.no_grad():
                return ref(*batch)
    else:
        kind = "port"
        from oracle import oracle_net
        ocfg = oracle_net.OracleConfig(backbone=c1.backbone, classifier_mode=c1.classifier_mode, num_classes=c1.num_classes,
                                       min_size=kw["test_image_min_size"], max_size=kw["image_max_size"])

        def fwd(batch):
            return oracle_net.forward(sd, ocfg, *batch)
    times = []
    for i in range(warmup + steps):
        batch = synth.make_batch(c1, i)
        t0 = time.perf_counter()
        fwd(batch)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    total = sum(times)
    what = "the unmodified reference ViBERTgridNet (oracle/_ref/reference)" if kind == "reference" else "oracle restatement (oracle/oracle_net.py)"
    return {"value": n_img * len(times) / total, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": f"{n_img} of the {cfg.batch} documents of a {cfg.name} step, {len(times)} timed forwards of {what}, "
                      f"fp32, eval, no_grad, torch {torch.__version__} CPU, {cores} threads",
            "ms_per_step": 1e3 * total / len(times)}


def time_region(fn, steps, stream_sync, world):
    """barrier + synchronize on both sides, CUDA events on the launching stream, max over ranks."""
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    from vibertgrid_pytorch_b200 import shard
    ms, _ = shard.aggregate_throughput(e0.elapsed_time(e1), 0, device="cuda")      # MAX over ranks
    return ms


def kernel_rooflines(net, cfg, dev, peaks):
    """Live CUDA-event timings of the dominant kernels, each launched alone on this stream with an L2 flush
    (512 MiB write) between launches.  Algorithmic bytes / FLOPs per launch: DESIGN.md section 5."""
    from vibertgrid_pytorch_b200 import ops, _lib
    out = {}
    # 512 MiB write between launches: flushes the 126 MB L2 and keeps the GPU busy (~80 us) while the host enqueues the
    # event + launch behind it, so the event pair brackets the kernel alone, not the Python launch latency
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` captures, KEYED BY KERNEL
    # NAME: a figure is attached only to the kernel that is actually launched below (profiles/r2_traffic.json)
    traffic = {}
    tp = os.path.join(ROOT, "profiles", "r2_traffic.json")
    if os.path.isfile(tp):
        with open(tp) as f:
            traffic = {k: v.get("dram_bytes") for k, v in json.load(f).items() if isinstance(v, dict)}

    def timed(fn, reps=10):
        fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return sum(ts) / len(ts)

    B, Hg, Wg, C, S = cfg.batch, cfg.height // 8, cfg.width // 8, 768, cfg.segments
    K = B * S
    g = torch.Generator().manual_seed(1)
    from vibertgrid_pytorch_b200 import synth
    boxes = torch.cat([synth.make_boxes(S, cfg.height, cfg.width, g) for _ in range(B)], 0).int().to(dev)
    seg_off = torch.arange(0, K + 1, S, dtype=torch.int32, device=dev)
    emb = torch.randn(K, C, device=dev)
    idx = ops.box_index_map(boxes, seg_off, B, 8, Hg, Wg)
    ps = net._get_engine()._ps()        # pre-split mode: the kernels run in the storage formats the engine uses (same bytes)
    emb_src = ops.to_split(emb) if ps else emb
    grid_out = ops.grid_scatter(emb_src, idx, seg_off)       # result buffers are allocated once: the event pair times the kernel
    ms = timed(lambda: ops.grid_scatter(emb_src, idx, seg_off, out=grid_out))
    by = B * C * Hg * Wg * 4 + K * C * 4 + K * 16 + B * Hg * Wg * 4
    sc_name = "grid_scatter_planes_kernel" if ps else "grid_scatter_kernel"
    out["grid_scatter"] = {"bound": "hbm", "achieved": by / ms / 1e6, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                           "frac": by / ms / 1e6 / peaks["hbm_gbs"], "traffic": traffic.get(sc_name), "ms": ms, "bytes": by, "kernel": sc_name}
    # the same kernel back to back over a working set larger than the 126 MB L2 (4 output grids of 100 MB, no flush in
    # between): the single-launch figure above ends with most of its output still in L2; this one is sustained HBM write-back
    try:
        grids = [ops.grid_scatter(emb_src, idx, seg_off) for _ in range(4)]
        torch.cuda.synchronize()
        reps = 5
        gs_graph = torch.cuda.CUDAGraph()           # replayed from a graph: a 21 us kernel is shorter than a Python launch
        with torch.cuda.graph(gs_graph):
            for _ in range(reps):
                for gbuf in grids:
                    ops.grid_scatter(emb_src, idx, seg_off, out=gbuf)
        gs_graph.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); gs_graph.replay(); e1.record(); torch.cuda.synchronize()
        ms_b2b = e0.elapsed_time(e1) / (reps * len(grids))
        out["grid_scatter"]["back_to_back"] = {"ms": ms_b2b, "achieved": by / ms_b2b / 1e6, "frac": by / ms_b2b / 1e6 / peaks["hbm_gbs"],
                                               "how": f"{reps * len(grids)} launches rotating 4 output grids ({4 * by / 1e6:.0f} MB > L2) replayed from one CUDA graph, no flush"}
        del grids, gs_graph
    except Exception as exc:      # diagnostics only
        print(f"[bench] back-to-back scatter timing skipped: {exc}", file=sys.stderr)
    Hf, Wf = cfg.height // 4, cfg.width // 4
    feat = torch.randn(B, Hf, Wf, 256, device=dev)
    if ps:
        feat = ops.to_split(feat)
    roi_out = ops.roi_align(feat, boxes, seg_off, 0.25, 7, split_out=ps)
    ms = timed(lambda: ops.roi_align(feat, boxes, seg_off, 0.25, 7, split_out=ps, out=roi_out))
    by = B * 256 * Hf * Wf * 4 + K * 256 * 49 * 4 + K * 20
    roi_name = ops.ROI_KERNEL_NAMES[ops.roi_variant(7, 256)]
    out["roi_align"] = {"bound": "hbm", "achieved": by / ms / 1e6, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                        "frac": by / ms / 1e6 / peaks["hbm_gbs"], "traffic": traffic.get(roi_name), "ms": ms, "bytes": by, "kernel": roi_name}
    try:        # the same kernel back to back over rotating feature maps (4 x 134 MB > L2), no flush: the second view, as for the scatter
        feats = [feat] + [ops.to_split(torch.randn(B, Hf, Wf, 256, device=dev)) if ps else torch.randn(B, Hf, Wf, 256, device=dev) for _ in range(3)]
        outs = [ops.roi_align(f_, boxes, seg_off, 0.25, 7, split_out=ps) for f_ in feats]
        torch.cuda.synchronize()
        gq = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gq):
            for _ in range(4):
                for f_, o_ in zip(feats, outs):
                    ops.roi_align(f_, boxes, seg_off, 0.25, 7, split_out=ps, out=o_)
        gq.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); gq.replay(); e1.record(); torch.cuda.synchronize()
        ms_b = e0.elapsed_time(e1) / 16
        out["roi_align"]["back_to_back"] = {"ms": ms_b, "achieved": by / ms_b / 1e6, "frac": by / ms_b / 1e6 / peaks["hbm_gbs"],
                                            "how": "16 launches rotating 4 feature maps (537 MB > L2) replayed from one CUDA graph, no flush"}
        del feats, outs, gq
    except Exception as exc:      # diagnostics only
        print(f"[bench] back-to-back ROI-align timing skipped: {exc}", file=sys.stderr)
    # dominant tensor-bound kernel: the BERT FFN-up GEMM of the packed batch (M = real rows, N=3072, K=768)
    eng = net._get_engine()
    prec = eng._prec()
    M = B * (cfg.seq_len + 2 * (cfg.seq_len // 510 + 1))
    A = torch.randn(M, 768, device=dev)
    Wt = torch.randn(3072, 768, device=dev) * 0.03
    bias = torch.zeros(3072, device=dev)
    ep = ops.make_epilogue(None, bias, act=ops.ACT_GELU)
    Ws = ops.split_bf16(Wt) if prec == ops.PREC_BF16X3 else None
    if ps:
        A = ops.to_split(A)
    ms = timed(lambda: ops.gemm(A, Wt, ep=ep, precision=prec, W_split=Ws, split_out=ps))
    # the same launch replayed back to back from a CUDA graph (no host gaps; operands L2-resident as they are inside the
    # forward, where the LayerNorm that produces A ran just before): reported beside the cold-L2 figure, not instead of it
    ms_warm = None
    try:
        out_w = ops.gemm(A, Wt, ep=ep, precision=prec, W_split=Ws, split_out=ps)
        torch.cuda.synchronize()
        gph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gph):
            for _ in range(20):
                out_w = ops.gemm(A, Wt, ep=ep, precision=prec, W_split=Ws, split_out=ps)
        gph.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); gph.replay(); e1.record(); torch.cuda.synchronize()
        ms_warm = e0.elapsed_time(e1) / 20
    except Exception as exc:      # diagnostics only
        print(f"[bench] warm GEMM timing skipped: {exc}", file=sys.stderr)
    # third view: operands LARGER THAN L2 instead of a flush -- 12 rotating (A, W, C) sets (~560 MB), 24 launches replayed from one
    # graph: every launch finds its operands in HBM, and the L2 holds what it holds inside the real step (the previous kernels'
    # outputs) instead of 126 MB of dirty flush lines whose write-back competes with the kernel's own fills
    ms_rot = None
    try:
        n_sets = 12
        sets = []
        for i in range(n_sets):
            Ai = torch.randn(M, 768, device=dev)
            Wi = torch.randn(3072, 768, device=dev) * 0.03
            sets.append((ops.to_split(Ai) if ps else Ai, Wi, ops.split_bf16(Wi) if prec == ops.PREC_BF16X3 else None))
        keep = [ops.gemm(a_, w_, ep=ep, precision=prec, W_split=ws_, split_out=ps) for a_, w_, ws_ in sets]
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            keep = [ops.gemm(*[(a_, w_)[j] for j in range(2)], ep=ep, precision=prec, W_split=ws_, split_out=ps) for _ in range(2) for a_, w_, ws_ in sets]
        gr.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
        ms_rot = e0.elapsed_time(e1) / (2 * n_sets)
        del keep, sets, gr
    except Exception as exc:      # diagnostics only
        print(f"[bench] rotating-operand GEMM timing skipped: {exc}", file=sys.stderr)
    fl = 2.0 * M * 3072 * 768
    # bf16x3: every fp32-equivalent product costs three bf16 tensor-core products, so the mode's ceiling is bf16 peak / 3
    peak, path = {ops.PREC_TF32: (peaks["tf32_tflops"], "tcgen05 kind::tf32 (measured cuBLAS TF32 peak)"),
                  ops.PREC_BF16X3: (peaks["bf16_tflops"] / 3.0, "tcgen05 kind::f16, 3 bf16 products per fp32-equivalent product"
                                                                  + (", operands pre-split in HBM (TMA-fed, no in-kernel conversion), CTA-pair tiles (cta_group::2)" if ps else "")
                                                                  + "; peak = measured bf16 peak / 3; frac == tensor-pipe share of bf16 peak"),
                  ops.PREC_FP32: (peaks["fp32_simt_tflops"], "CUDA-core fp32 FFMA")}[prec]
    # Primary figure: the average launch duration over a timed region of back-to-back launches whose operands are larger than
    # L2 (the timing rule's second option), i.e. every launch streams A and W from HBM.  The flush-per-launch figure is kept
    # beside it: there each launch also pays for writing back the 126 MB of dirty lines the flush leaves in L2.
    ms_flush = ms
    if ms_rot:
        ms = ms_rot
    out["gemm_ffn_up"] = {"bound": "tensor", "achieved": fl / ms / 1e9, "peak": peak, "unit": "TFLOP/s",
                          "frac": fl / ms / 1e9 / peak, "traffic": traffic.get("gemm_ps2_kernel<256, 64, 3>") if ps else None, "ms": ms, "flops": fl, "path": path,
                          "executed_tensor_tflops": (3.0 if prec == ops.PREC_BF16X3 else 1.0) * fl / ms / 1e9,
                          "shape": [M, 3072, 768],
                          "timing": ("operands beyond L2: 12 rotating (A, W, C) sets (~560 MB > 126 MB L2), 24 launches replayed from one CUDA graph, "
                                     "CUDA events around the region, no flush") if ms_rot else "cold L2 (512 MiB flush before every launch)",
                          "flush_cold": {"ms": ms_flush, "achieved": fl / ms_flush / 1e9, "frac": fl / ms_flush / 1e9 / peak,
                                         "how": "one launch at a time, 512 MiB write (L2 flush) before each: the launch also evicts 126 MB of dirty flush lines"},
                          "warm_l2": None if not ms_warm else {"ms": ms_warm, "achieved": fl / ms_warm / 1e9, "frac": fl / ms_warm / 1e9 / peak,
                                                               "how": "20 launches on ONE operand set replayed back to back from one CUDA graph"}}
    return out


def library_bar(cfg, dev, steps=3, warmup=2):
    """The library-kernel bar: what stock library kernels reach on this B200 for the same work --
      * the UNMODIFIED reference ViBERTgridNet in eager mode on the GPU (HF BertModel on cuBLAS, cuDNN convolutions,
        torchvision's CUDA roi_align, its own Python glue with its device->host syncs), fp32 and with TF32 allowed;
      * cuBLAS on the dominant GEMM's shape (fp32 / TF32 / bf16) and torchvision.ops.roi_align on the ROI kernel's shape.
    Library calls appear ONLY here, as the bar the hand-written path is held against; never on the hot path."""
    import contextlib
    import dataclasses
    from vibertgrid_pytorch_b200 import synth
    out = {}
    ev = lambda: torch.cuda.Event(enable_timing=True)

    def t_ms(fn, reps=10):
        fn(); torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            e0, e1 = ev(), ev()
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return sum(ts) / len(ts)

    M = cfg.batch * (cfg.seq_len + 2 * (cfg.seq_len // 510 + 1))
    a, w = torch.randn(M, 768, device=dev), torch.randn(3072, 768, device=dev)
    fl = 2.0 * M * 3072 * 768
    gemm = {}
    for name, tf32 in (("fp32", False), ("tf32", True)):
        torch.backends.cuda.matmul.allow_tf32 = tf32
        ms = t_ms(lambda: torch.nn.functional.linear(a, w))
        gemm[name] = {"ms": ms, "tflops": fl / ms / 1e9}
    torch.backends.cuda.matmul.allow_tf32 = False
    ab, wb = a.bfloat16(), w.bfloat16()
    ms = t_ms(lambda: torch.nn.functional.linear(ab, wb))
    gemm["bf16"] = {"ms": ms, "tflops": fl / ms / 1e9}
    out["cublas_ffn_up"] = {"shape": [M, 3072, 768], **gemm, "note": "warm L2, no bias / GELU epilogue"}
    try:
        import torchvision
        B, S = cfg.batch, cfg.segments
        g = torch.Generator().manual_seed(1)
        boxes = torch.cat([synth.make_boxes(S, cfg.height, cfg.width, g) for _ in range(B)], 0).float().to(dev)
        bidx = torch.arange(B, device=dev).repeat_interleave(S).float()[:, None]
        feat = torch.randn(B, 256, cfg.height // 4, cfg.width // 4, device=dev)
        rois = torch.cat([bidx, boxes], 1)
        ms = t_ms(lambda: torchvision.ops.roi_align(feat, rois, output_size=7, spatial_scale=0.25, sampling_ratio=-1, aligned=False))
        out["torchvision_roi_align"] = {"ms": ms, "layout": "NCHW fp32 (the reference's call, model/grid_roi_align.py:81)"}
    except Exception as exc:
        out["torchvision_roi_align"] = {"error": str(exc)[:120]}
    ref_dir = _staged_reference()
    if ref_dir is not None:
        try:
            sys.path.insert(0, ref_dir)
            net, kw = build_net(cfg)
            sd = {k: v for k, v in net.state_dict().items()}
            cwd = os.getcwd()
            with tempfile.TemporaryDirectory() as tmp:
                os.chdir(tmp)
                try:
                    synth.write_bert_dir(cfg, tmp)
                    with contextlib.redirect_stdout(sys.stderr):
                        from model.ViBERTgrid_net import ViBERTgridNet as RefNet
                        ref = RefNet(**synth.model_kwargs(cfg, "eval"))
                finally:
                    os.chdir(cwd)
            ref.load_state_dict(sd, strict=True)
            ref = ref.to(dev).eval()
            batches = [to_device(synth.make_batch(cfg, i), dev, False) for i in range(2)]
            res = {}
            for name, tf32 in (("fp32", False), ("tf32", True)):
                torch.backends.cuda.matmul.allow_tf32 = tf32
                torch.backends.cudnn.allow_tf32 = tf32
                with torch.no_grad():
                    for i in range(warmup):
                        ref(*batches[i % 2])
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    for i in range(steps):
                        ref(*batches[i % 2])
                    torch.cuda.synchronize()
                    dt = time.perf_counter() - t0
                res[name] = {"images_per_s": cfg.batch * steps / dt, "ms_per_step": 1e3 * dt / steps}
            torch.backends.cuda.matmul.allow_tf32 = False
            torch.backends.cudnn.allow_tf32 = False
            out["reference_eager_cuda"] = {**res, "what": f"unmodified reference ViBERTgridNet on this GPU, eager, batch {cfg.batch}, {steps} timed steps "
                                                          "(wall clock: its Python glue synchronises the device thousands of times per step)"}
            del ref
            torch.cuda.empty_cache()
        except Exception as exc:
            out["reference_eager_cuda"] = {"error": f"{type(exc).__name__}: {exc}"[:200]}
    return out


def serving_latency(dev, n=200):
    """the deployment path -- ``model.inference(image, seg_indices, coors, corpus, mask)`` (reference
    deployment/module_load.py -> model/ViBERTgrid_net.py:470-499) for ONE document of BASELINE configs[0] (cfg1: r18 +
    bert-base, 512x512, 512 tokens, 128 boxes), CUDA-graph replay, host buffers in pinned memory: per-request latency from
    the first H2D byte to the predictions read back on the host."""
    import dataclasses
    from vibertgrid_pytorch_b200 import synth
    cfg = synth.CONFIGS["cfg1"]
    net, _ = build_net(cfg, dev)
    host = [pin(synth.make_batch(cfg, 100 + i)) for i in range(4)]
    out_host = torch.empty((cfg.segments, cfg.num_classes), dtype=torch.float32).pin_memory()
    lat = []
    for i in range(n + 10):
        t0 = time.perf_counter()
        img, seg, cls, coors, corpus, mask = to_device(host[i % 4], dev, True)
        pred = net.inference(img, seg, coors, corpus, mask)
        out_host.copy_(pred, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        if i >= 10:
            lat.append((time.perf_counter() - t0) * 1e3)
    lat.sort()
    return {"workload": "cfg1, one document per request, inference() entry point, host -> host", "requests": n,
            "p50_ms": lat[len(lat) // 2], "p99_ms": lat[int(len(lat) * 0.99) - 1], "mean_ms": sum(lat) / len(lat),
            "documents_per_s_one_stream": 1e3 / (sum(lat) / len(lat)), "graph_replays": net._get_engine().graph_replays}


def input_pipeline_arm(net, cfg, dev, rank, world, steps):
    """the same e2e step fed by the package's input pipeline instead of ready-made host batches -- documents come
    out of a shard (pre-tokenised, decoded uint8 pixels; shards.py / csrc/vbg_shard.cpp), the native collate gathers each
    step's batch into ONE pinned staging buffer on a background thread, ONE host->device copy per step on a side stream, the
    decode kernel applies ToTensor's / 255 on the device.  Also the host-only rate of the reader + collate (no GPU in it)."""
    from vibertgrid_pytorch_b200 import shards, synth
    from vibertgrid_pytorch_b200.prefetch import HostResultQueue
    n_batches_distinct = 4
    docs = []
    for i in range(n_batches_distinct):
        img, seg, cls, coors, corpus, mask = synth.make_batch(cfg, 7000 + 1000 * rank + i)
        for b in range(len(img)):
            n = int(seg[b].shape[0])
            docs.append(dict(image=(img[b].permute(1, 2, 0) * 255.0).round().clamp(0, 255).to(torch.uint8).numpy(),
                             corpus=corpus[b, :n].numpy(), seg_ids=seg[b].numpy(), classes=cls[b].numpy(), coors=coors[b].numpy()))
    tmp = tempfile.mkdtemp(prefix="vbg_shard_")
    path = os.path.join(tmp, f"bench_rank{rank}.vbgshard")
    shards.write_shard(path, docs)
    sh = shards.Shard(path)
    B = cfg.batch
    plan = lambda n: [[(i % n_batches_distinct) * B + b for b in range(B)] for i in range(n)]
    results, sink = HostResultQueue(), {}

    def run(n, timed):
        ld = shards.ShardLoader(sh, batches=plan(n), device=dev, depth=3, threads=4)
        it = iter(ld)

        def step(i):
            loss, pm, ps, gt, pred = net(*next(it))
            results.push(pred, loss)
            while len(results) > (0 if i == n - 1 else 1):
                sink["pred"], sink["loss"] = results.pop()
        if not timed:
            for i in range(n):
                step(i)
            return None, ld
        return time_region(step, n, True, world), ld

    run(3, False)                                    # eager, capture, replay of the uint8 batch signature
    ms, ld = run(steps, True)
    out = {"value": B * steps * world / (ms / 1e3), "unit": UNIT, "ms_per_step": ms / steps,
           "h2d_bytes_per_step": ld.h2d_bytes // steps, "h2d_copies_per_step": 1,
           "what": "ShardLoader(device) -> net() -> HostResultQueue: shard -> native collate (background thread, pinned staging) -> one "
                   "H2D per step -> uint8 decode kernel -> joint forward -> D2H of pred_label + loss"}
    if rank == 0:                                    # host-only: reader + collate into pinned staging
        lay = sh.layout(plan(1)[0])
        staging = torch.empty(lay[0], dtype=torch.uint8).pin_memory()
        for threads in (1, 4):
            sh.collate_into(plan(1)[0], staging, threads)
            t0, n = time.perf_counter(), 0
            while time.perf_counter() - t0 < 0.5:
                sh.collate_into(plan(n + 1)[n], staging, threads)
                n += 1
            dt = time.perf_counter() - t0
            out[f"host_collate_{threads}t"] = {"documents_per_s": B * n / dt, "gb_per_s": lay[0] * n / dt / 1e9, "batch_bytes": int(lay[0])}
    sh.close()
    try:
        os.remove(path); os.rmdir(tmp)
    except OSError:
        pass
    return out


def measured_peaks(dev):
    """MEASURED_PEAKS.json (driver-written) + a live cuBLAS TF32 / fp32 GEMM measured the same way
    (library call used ONLY as the roofline denominator, never on the hot path)."""
    peaks = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "source": "fallback (B200_PROFILING.md)"}
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as f:
            j = json.load(f)
        peaks.update(hbm_gbs=j["hbm_gbs"], bf16_tflops=j["bf16_tflops"], bf16_tflops_sustained=j.get("bf16_tflops_sustained"),
                     source="measured (MEASURED_PEAKS.json)")
    n = 8192
    a, b = torch.randn(n, n, device=dev), torch.randn(n, n, device=dev)
    best = {}
    for name, allow in (("tf32_tflops", True), ("fp32_simt_tflops", False)):
        torch.backends.cuda.matmul.allow_tf32 = allow
        torch.matmul(a, b)
        torch.cuda.synchronize()
        t = []
        for _ in range(5 if allow else 2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); torch.matmul(a, b); e1.record()
            torch.cuda.synchronize()
            t.append(e0.elapsed_time(e1))
        best[name] = 2.0 * n ** 3 / min(t) / 1e9
    torch.backends.cuda.matmul.allow_tf32 = False
    peaks.update(best)
    peaks["tf32_how"] = "torch.matmul fp32 8192^3 with allow_tf32 (cuBLAS), best of 5, this run"
    return peaks


def train_step_arm(net, cfg, resident, n_rot, world, steps, warmup=3):
    """One training step = the reference's loop body (pipeline/train_val_utils.py:265-281): loss = model(batch) in train mode,
    zero_grad, loss.backward(), SGD step for the CNN / heads and AdamW step for the ``bert_model`` parameters
    (train_SROIE.py:217-235); at N > 1 the gradients are averaged over ranks with NCCL (shard.allreduce_gradients) before the
    optimizer steps.  Timed like the forward: CUDA events, barrier + synchronize on both sides, max over ranks."""
    from vibertgrid_pytorch_b200 import _lib, shard
    net.train()
    bert = [p for n, p in net.named_parameters() if "bert_model" in n]
    cnn = [p for n, p in net.named_parameters() if "bert_model" not in n]
    params = list(net.parameters())
    # the reference's two optimizers (train_SROIE.py:217-235) as one multi-tensor kernel launch each (optim.py)
    from vibertgrid_pytorch_b200.optim import FusedAdamW, FusedSGD
    opt_cnn = FusedSGD(cnn, lr=1e-4, momentum=0.9, weight_decay=5e-4)
    opt_bert = FusedAdamW(bert, lr=1e-6, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2)
    losses = []

    ar = {"ms": 0.0, "bytes": 0, "n": 0, "on": False}

    def step(i):
        loss = net(*resident[i % n_rot])
        opt_cnn.zero_grad()
        opt_bert.zero_grad()
        if world > 1 and ar["on"]:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        # graphed step: every gradient already sits in one flat arena -> in-place NCCL AVG over slices of it, then backward()
        # hands the averaged views to the parameters; eager step (first sighting of a batch signature): bucketed all-reduce
        # of .grad after backward()
        nbytes_ar = shard.allreduce_step_arena(net)
        loss.backward()
        if nbytes_ar == 0:
            shard.allreduce_gradients(params)
        if world > 1 and ar["on"] and nbytes_ar:
            e1.record()
            ar.setdefault("events", []).append((e0, e1))
            ar["bytes"] = nbytes_ar
        opt_cnn.step()
        opt_bert.step()
        losses.append(loss.detach())

    warmup = max(warmup, 2 * n_rot)        # every rotating batch signature is seen twice: eager, then captured
    for i in range(warmup):
        step(i)
    eng = net._train_engine
    c0, r0, k0 = _lib.launch_count, eng.graph_replays, eng.kernel_launches
    ar["on"] = True
    ms = time_region(step, steps, True, world)
    ar["on"] = False
    # kernels of this library executed in the timed region: host-issued C-ABI launches + those replayed from the step graphs
    # (the capture itself happens in the warm-up; its host-side calls are not counted twice)
    launches = (_lib.launch_count - c0) + (eng.kernel_launches - k0)
    vals = [float(l) for l in losses]
    out = {"value": cfg.batch * steps * world / (ms / 1e3), "unit": UNIT, "ms_per_step": ms / steps, "steps": steps, "warmup": warmup,
           "includes": "train-mode forward (batch-stat BN, hidden + attention dropout) + backward + fused multi-tensor SGD/AdamW steps"
                       + (" + NCCL gradient all-reduce" if world > 1 else ""),
           "whole_step_cuda_graph_replays": eng.graph_replays - r0,
           "gpu_launches": launches, "loss_first": vals[0], "loss_last": vals[-1],
           "peak_mem_gib": torch.cuda.max_memory_allocated() / 2 ** 30}
    if ar.get("events"):
        t = [a.elapsed_time(b) for a, b in ar["events"]]
        msa = sum(t) / len(t)
        out["allreduce"] = {"exposed_ms_per_step": msa, "bytes": ar["bytes"], "how": "in-place NCCL AVG over slices of the step's flat gradient arena (shard.allreduce_step_arena)",
                            "bus_gbs": 2.0 * (world - 1) / world * ar["bytes"] / msa / 1e6}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg2")
    ap.add_argument("--precision", default=None, choices=[None, "fp32", "tf32", "bf16x3"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the training-step measurement (the `train_step` object)")
    ap.add_argument("--no-input-pipeline", action="store_true", help="skip the shard-fed e2e measurement (the `input_pipeline` object)")
    ap.add_argument("--no-serving", action="store_true", help="skip the serving-latency measurement (the `serving` object)")
    ap.add_argument("--no-library-bar", action="store_true", help="skip the library-kernel bar (reference eager on the GPU, cuBLAS, torchvision)")
    ap.add_argument("--mode", default="forward", choices=["forward", "train"],
                    help="train: the line's value / ms_per_step are the TRAINING step (forward + backward + all-reduce + optimizers)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    # stdout carries exactly ONE JSON line: libraries that write to fd 1 (NCCL prints its version banner there) go to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        os.write(json_fd, (json.dumps(line) + "\n").encode())

    from vibertgrid_pytorch_b200 import synth
    cfg = synth.CONFIGS[args.config]
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    workload = (f"{cfg.name} eval fwd: {cfg.backbone}+bert-base, {cfg.batch} docs/GPU/step, {cfg.height}x{cfg.width}, L={cfg.seq_len}, "
                f"S={cfg.segments}, {cfg.classifier_mode}")           # < 150 characters: the driver's record truncates longer strings
    # the SAME config object in both arms (the reference arm times the reference's CPU path on this workload)
    config = {"workload": workload, "global_batch": cfg.batch * world, "parallelism": f"dp{world}"}

    if args.impl == "reference":
        if rank != 0:
            return
        r = cpu_reference_arm(cfg, max(args.steps, 1), args.warmup)          # one step = the workload's own batch
        line = {"metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "impl": "reference", "config": config,
                "arm": "rank 0 only: the reference's CPU forward on the host cores, one workload batch per step",
                "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        emit(line)
        return

    assert torch.cuda.is_available(), "bench.py (impl=ours) needs a GPU; there is no CPU fallback"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from vibertgrid_pytorch_b200 import _lib, ops
    if args.precision:
        os.environ["VBG_PRECISION"] = args.precision
    net, kw = build_net(cfg, dev)
    eng = net._get_engine()
    prec = eng._prec()

    n_rot = 4      # rotate distinct documents; weights (0.6 GB) + activations exceed the 126 MB L2 anyway
    from vibertgrid_pytorch_b200 import shard
    host = [pin(synth.make_batch(cfg, shard.batch_seed(rank, world, i, n_rot))) for i in range(n_rot)]
    resident = [to_device(b, dev, False) for b in host]
    torch.cuda.synchronize()

    def step_resident(i):
        net(*resident[i % n_rot])

    sink = {}

    from vibertgrid_pytorch_b200.prefetch import DevicePrefetcher, HostResultQueue
    feed, results = {}, HostResultQueue()

    def step_e2e(i):
        # the package's input pipeline: every step's batch is copied from pinned host memory inside the timed region, on a
        # side stream, overlapping the previous step's kernels (the first next() of a feed uploads batches 0 and 1)
        batch = next(feed["it"])
        loss, pm, ps, gt, pred = net(*batch)
        results.push(pred, loss)                # D2H of THIS step's result (+ loss), asynchronously into pinned memory ...
        if len(results) > 1 or i == feed["n"] - 1:
            while len(results) > (0 if i == feed["n"] - 1 else 1):
                sink["pred"], sink["loss"] = results.pop()     # ... read on the host one step later (last step: at once)

    def new_feed(n):
        feed["it"], feed["n"] = iter(DevicePrefetcher((host[i % n_rot] for i in range(n)), dev)), n

    for i in range(args.warmup):
        step_resident(i)
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    c0 = eng.kernel_launches
    ms = time_region(step_resident, args.steps, True, world)
    launches = eng.kernel_launches - c0
    new_feed(2)
    for i in range(2):
        step_e2e(i)
    new_feed(args.steps)
    ms_e2e = time_region(step_e2e, args.steps, True, world)
    pipeline = None
    if not args.no_input_pipeline:
        try:
            pipeline = input_pipeline_arm(net, cfg, dev, rank, world, args.steps)
        except Exception as e:          # secondary measurement: never loses the headline line
            pipeline = {"error": f"{type(e).__name__}: {e}"[:300]}
    clocks = sampler.stop() if sampler else None

    train = None
    if not args.no_train:
        try:
            train = train_step_arm(net, cfg, resident, n_rot, world, args.steps if args.mode == "train" else min(args.steps, 10))
        except Exception as e:          # the headline (forward) line must survive a failure of the secondary measurement
            train = {"error": f"{type(e).__name__}: {e}"[:300]}
        net.eval()

    imgs = cfg.batch * args.steps * world
    value = imgs / (ms / 1e3)
    e2e = imgs / (ms_e2e / 1e3)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {ops.PREC_TF32: "tf32", ops.PREC_BF16X3: "bf16x3 (hi/lo split, f32 accumulate)", ops.PREC_FP32: "f32"}[prec],
            "data": "synthetic",
            "config": dict(config),
            "notes": {"sharding": "documents sharded over ranks, no data-path collective",
                      "l2": "working set (605 MB weights + activations) >> 126 MB L2; 4 input batches rotated",
                      "scope": "value / e2e: eval-mode joint forward (the reference arm's workload); the training step (forward + "
                               "backward + gradient all-reduce + optimizers, configs[1] 'forward+backward') is the train_step object"},
            "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": ms_e2e / args.steps, "h2d_bytes_per_step": nbytes(host[0]),
                    "d2h_bytes_per_step": int(sink["pred"].numel() * 4 + sink["loss"].numel() * sink["loss"].element_size()),
                    "d2h_contents": "pred_label [K, C] fp32 + loss: what eval_SROIE.py / validate() read; pred_mask / pred_ss stay on the device"},
            "gpu_launches": launches, "cuda_graph_replays": eng.graph_replays, "clocks": clocks}
    if args.mode == "train":
        assert train is not None and "ms_per_step" in train, f"--mode train: the training-step arm failed: {train}"
        line["forward"] = {"value": value, "ms_per_step": ms / args.steps, "e2e": e2e}
        line.update(metric="train_" + METRIC, value=train["value"], ms_per_step=train["ms_per_step"], gpu_launches=train["gpu_launches"])
        line["config"]["workload"] = workload.replace("eval fwd", "TRAIN step (fwd+bwd+allreduce+SGD/AdamW)")
    if pipeline is not None:
        line["input_pipeline"] = pipeline
        if "value" in pipeline:         # numeric copy where the driver's record keeps it
            line["e2e"]["from_shards_images_per_s"] = round(pipeline["value"], 1)
            line["e2e"]["from_shards_h2d_bytes_per_step"] = int(pipeline["h2d_bytes_per_step"])
    if train is not None:
        line["train_step"] = train
        if "ms_per_step" in train:      # numeric copies where the driver's record keeps them
            line["config"]["train_step_ms"] = round(train["ms_per_step"], 3)
            line["config"]["train_images_per_s"] = round(train["value"], 1)
            line["e2e"]["train_step_images_per_s"] = round(train["value"], 1)
    if rank == 0:
        fl = fwd_flops_as_executed(cfg)
        line["fwd_tflops_as_executed"] = fl * cfg.batch * args.steps / (ms / 1e3) / 1e12
        if not args.no_roofline:
            peaks = measured_peaks(dev)
            kr = kernel_rooflines(net, cfg, dev, peaks)
            line["roofline"] = kr["gemm_ffn_up"]
            line["roofline_hbm_kernels"] = {k: kr[k] for k in ("grid_scatter", "roi_align")}
            # numeric copies inside the object the driver's record keeps
            line["roofline"]["hbm_scatter_frac"] = round(kr["grid_scatter"]["frac"], 4)
            line["roofline"]["hbm_scatter_b2b_frac"] = round(kr["grid_scatter"].get("back_to_back", {}).get("frac", 0.0), 4)
            line["roofline"]["hbm_roi_align_frac"] = round(kr["roi_align"]["frac"], 4)
            line["roofline"]["hbm_roi_align_b2b_frac"] = round(kr["roi_align"].get("back_to_back", {}).get("frac", 0.0), 4)
            line["roofline"]["flush_cold_frac"] = round(kr["gemm_ffn_up"]["flush_cold"]["frac"], 4)
            line["peaks"] = peaks
        if world == 1 and args.mode == "forward" and not args.no_serving:
            try:
                line["serving"] = serving_latency(dev)
                line["e2e"]["serve_cfg1_p50_ms"] = round(line["serving"]["p50_ms"], 3)
                line["e2e"]["serve_cfg1_p99_ms"] = round(line["serving"]["p99_ms"], 3)
            except Exception as exc:
                line["serving"] = {"error": f"{type(exc).__name__}: {exc}"[:200]}
        if not args.no_library_bar and world == 1:
            try:
                line["library_bar"] = library_bar(cfg, dev)
                rb = line["library_bar"].get("reference_eager_cuda", {})
                if "fp32" in rb:       # numeric copies where the driver's record keeps them
                    line["e2e"]["reference_eager_cuda_fp32_images_per_s"] = round(rb["fp32"]["images_per_s"], 2)
                    line["e2e"]["reference_eager_cuda_tf32_images_per_s"] = round(rb["tf32"]["images_per_s"], 2)
            except Exception as exc:
                line["library_bar"] = {"error": f"{type(exc).__name__}: {exc}"[:200]}
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = {k: v for k, v in cpu_reference_arm(cfg, 5, 1, sample_images=min(2, cfg.batch)).items()
                                    if k != "ms_per_step"}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
