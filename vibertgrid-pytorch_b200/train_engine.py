"""Training-mode joint forward of ViBERTgridNet as a differentiable sequence of sm_100a kernel stages
(reference ``model/ViBERTgrid_net.py:512-544`` under ``model.train()``, driven by ``pipeline/train_val_utils.py:265-281``:
``loss = model(...)``, ``loss.backward()``, optimizer steps).

What differs from the eval engine (engine.py):
  * every stage is a ``torch.autograd.Function`` of ``autograd.py`` (fp32 channels-last tensors between stages), so
    ``loss.backward()`` fills ``.grad`` of the registered parameters -- DDP hooks, ``clip_grad_norm`` and the optimizers of the
    reference's training loop see ordinary gradients;
  * BatchNorm uses batch statistics and updates ``running_mean / running_var / num_batches_tracked`` (momentum form);
  * the BERT hidden-state dropouts are applied (counter-based mask kernel) and so is the dropout on the attention
    probabilities (inside the attention kernels, forward and backward);
  * the image min-size is drawn per image like ``transform.py:124-131,192-194`` (same torch CPU RNG consumption);
  * WHOLE-STEP CUDA GRAPH: a batch signature seen for the second time is captured -- train-mode forward AND the backward to
    every parameter gradient -- into one CUDA graph (``_graphed_loss``).  From then on ``model(batch)`` copies the inputs into
    the graph's static buffers, refreshes one device word (the dropout step seed) and replays the graph; the returned loss is
    the output of a one-node autograd Function whose backward hands the already-computed gradients (times the incoming
    ``grad_output``, e.g. GradScaler's scale) to the parameters, so ``loss.backward()``, DDP's hooks, ``clip_grad_norm`` and the
    optimizers of the reference's loop work unchanged while ~2 000 Python-issued launches per step collapse into one.

Heads: ``simp`` (BASELINE configs[1]-[3]), ``full`` (binary gate + C-1 binary heads) and ``crf`` (emissions + the CRF
negative log-likelihood kernel, BASELINE configs[4]).  The ``simp`` head with the default auxiliary loss uses the fused
label-paint + cross-entropy kernel on the low-resolution logits; every other auxiliary-loss configuration (the two-stage
segmentation head of ``full`` / ``crf``, sampled / OHEM / class-weighted losses) paints the label maps with vbg_label_paint and
applies the reference's loss arithmetic (losses.py) to the nearest-upsampled logits.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import autograd as A
from . import ops
from .plan import plan_batch, roberta_position_ids


FORCE_SYNC_BN_SINGLE_RANK = False      # tests only: run the SyncBatchNorm code path in a one-rank process group


def _rand_seed():
    return int(torch.empty((), dtype=torch.int64).random_(0, 2 ** 62).item())       # CPU generator: no device sync


class _parameters_as:
    """Context: every module attribute that resolves to a registered Parameter ``p`` with ``id(p)`` in ``aliases`` resolves to
    the alias tensor instead (an instance-dict entry shadows ``nn.Module.__getattr__``); the registration itself is untouched,
    so ``named_parameters()``, the optimizers and DDP never see a difference.  (torch's ``_reparametrize_module`` cannot be
    used: the BERT module is registered under two parents -- ``bert_model`` and ``BERTgrid_generator.model``, the reference's
    state-dict layout -- and its tie handling leaves such a module holding the aliases on exit.)"""

    def __init__(self, net, aliases):
        self.net, self.aliases, self.done = net, aliases, []

    def __enter__(self):
        for mod in self.net.modules():
            for name, p in mod._parameters.items():
                if p is not None and id(p) in self.aliases and name not in mod.__dict__:
                    mod.__dict__[name] = self.aliases[id(p)]
                    self.done.append((mod, name))
        return self

    def __exit__(self, *exc):
        for mod, name in self.done:
            del mod.__dict__[name]
        return False


class _StepGradsF(torch.autograd.Function):
    """The one autograd node of a graphed training step: ``forward`` returns the loss the replayed CUDA graph produced;
    ``backward`` hands each parameter its gradient -- already computed by the same replay -- times the incoming grad_output."""

    @staticmethod
    def forward(ctx, loss, n_grads, *rest):
        ctx.grads = rest[:n_grads]                     # static tensors of the graph (rewritten by the next replay)
        return loss.clone()

    @staticmethod
    def backward(ctx, go):
        scaled = torch._foreach_mul(list(ctx.grads), go.reshape(()))       # fresh tensors: .grad never aliases graph memory
        return (None, None) + (None,) * len(ctx.grads) + tuple(scaled)


class TrainEngine:
    def __init__(self, net):
        import os
        self.net = net
        self.use_graphs = os.environ.get("VBG_TRAIN_GRAPHS", "1") != "0"
        self.side_wgrad = os.environ.get("VBG_TRAIN_SIDE_WGRAD", "1") != "0"
        self.side_prep = os.environ.get("VBG_TRAIN_SIDE_PREP", "1") != "0"
        self.bert_planes = os.environ.get("VBG_TRAIN_PLANES", "1") != "0"
        self.max_graphs = 4
        self._graphs = {}
        self.graph_replays = 0
        self.kernel_launches = 0          # C-ABI kernel launches executed on the device by graph replays (bench.py: gpu_launches)
        self._step_seed = None
        self._graph_step = False
        self.capture_failures = 0
        self._side_streams = None
        self.grad_arena, self.arena_fresh = None, False

    def invalidate(self):
        self._graphs.clear()

    # ------------------------------------------------------------------ building blocks
    @staticmethod
    def _sync_group(bn):
        """``(process_group,)`` when ``bn`` is an nn.SyncBatchNorm (the reference's ``convert_sync_batchnorm`` under
        ``syncBN: True``, train_SROIE.py:203-205) in a job of more than one rank, else None -- torch's own rule
        (SyncBatchNorm falls back to per-rank statistics in a single-process job)."""
        if not isinstance(bn, nn.SyncBatchNorm):
            return None
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()):
            return None
        group = getattr(bn, "process_group", None)
        if dist.get_world_size(group) == 1 and not FORCE_SYNC_BN_SINGLE_RANK:
            return None
        return (group,)

    def _bn(self, x, bn: nn.modules.batchnorm._BatchNorm, relu=False, residual=None):
        stats = []
        y = A.BatchNormTrainF.apply(x, bn.weight, bn.bias, residual, relu, bn.eps, stats, self._sync_group(bn))
        if bn.track_running_stats and bn.running_mean is not None:
            mean, var, n = stats[0]

            def update():         # five tiny kernels per BatchNorm that nothing in the step reads: off the main stream when captured
                with torch.no_grad():
                    bn.num_batches_tracked += 1
                    m = bn.momentum if bn.momentum is not None else 1.0 / float(bn.num_batches_tracked)
                    bn.running_mean.mul_(1.0 - m).add_(mean, alpha=m)
                    if isinstance(n, torch.Tensor):               # SyncBatchNorm: the global row count lives on the device
                        bn.running_var.mul_(1.0 - m).add_(var * (n / (n - 1.0).clamp_min(1.0)).float(), alpha=m)
                    else:
                        bn.running_var.mul_(1.0 - m).add_(var, alpha=m * n / max(n - 1, 1))
            A._deferred(update, True, mean, var, n)
        return y

    @staticmethod
    def _conv(x, conv: nn.Conv2d):
        k, s, p = conv.kernel_size[0], conv.stride[0], conv.padding[0]
        if k == 1 and s == 1 and p == 0:
            B, H, W, Cin = x.shape
            y = A.linear(x.reshape(B * H * W, Cin), A.leaf_view(conv.weight, conv.out_channels, Cin), conv.bias)
            return y.view(B, H, W, conv.out_channels)
        return A.ConvPS.apply(x, conv.weight, conv.bias, s, p)

    def _block(self, x, conv1, bn1, conv2, bn2, shortcut):
        y = self._bn(self._conv(x, conv1), bn1, relu=True)
        if shortcut is None:
            sc = x
        else:
            kind, sconv, sbn = shortcut
            sc = self._bn(self._conv(A.AvgPoolF.apply(x) if kind == "avg" else x, sconv), sbn)
        return self._bn(self._conv(y, conv2), bn2, relu=True, residual=sc)

    def _our_block(self, x, blk):
        sc = None
        if blk.downsample:
            sc = ("avg", blk.conv_shortcut[1], blk.conv_shortcut[2]) if blk.d_variant \
                else ("conv", blk.conv_shortcut[0], blk.conv_shortcut[1])
        return self._block(x, blk.conv_1, blk.bn_1, blk.conv_2, blk.bn_2, sc)

    def _tv_block(self, x, blk):
        sc = ("conv", blk.downsample[0], blk.downsample[1]) if hasattr(blk, "downsample") else None
        return self._block(x, blk.conv1, blk.bn1, blk.conv2, blk.bn2, sc)

    # ------------------------------------------------------------------ stages
    def _bert(self, plan, dt, corpus):
        bm = self.net.bert_model
        e = bm.embeddings
        cfg = bm.cfg
        heads = cfg["num_attention_heads"]
        p_drop = float(getattr(self.net, "bert_hidden_dropout", 0.1))
        # dropout on the attention probabilities (HF attention_probs_dropout_prob), applied inside the attention kernels
        p_attn = float(getattr(self.net, "bert_attn_dropout", cfg.get("attention_probs_dropout_prob", 0.1)))
        cu = dt["cu"]
        ids, pos = ops.bert_assemble(corpus, dt["seq_tab"], cu, plan.nseq, plan.R)
        if cfg.get("roberta"):
            pos = roberta_position_ids(ids, pos, int(cfg["pad_token_id"]), cu)
        x = A.EmbedSumF.apply(e.word_embeddings.weight, e.position_embeddings.weight, e.token_type_embeddings.weight, ids, pos)
        x = A.LayerNormPS.apply(x, e.LayerNorm.weight, e.LayerNorm.bias, e.LayerNorm.eps)

        step_seed = self._seed_word()                      # device word, refreshed per step (None outside the graphed path)

        def drop(t):
            return A.DropoutF.apply(t, p_drop, _rand_seed(), step_seed) if p_drop > 0.0 else t

        x = drop(x)
        # QKV -> attention and FFN-up -> GELU -> FFN-down keep their wide tensors in the plane format (autograd.py "planes protocol")
        planes = bool(self.bert_planes and not getattr(self, "_test_standins", False) and not A._fp32()
                      and plan.max_len <= 512 and ops.tc_available() and cfg["hidden_size"] // heads == 64)
        for lyr in bm.encoder.layer:
            sa = lyr.attention.self
            wqkv = A.pack_rows(sa.query.weight, sa.key.weight, sa.value.weight)       # cat whose backward is three views
            bqkv = A.pack_rows(sa.query.bias, sa.key.bias, sa.value.bias)
            qkv = A.linear(x, wqkv, bqkv, out_planes=planes)
            ctx = A.AttentionF.apply(qkv, cu, plan.nseq, plan.max_len, heads, p_attn, _rand_seed() if p_attn > 0.0 else 0, step_seed)
            ao = lyr.attention.output
            a = drop(A.linear(ctx, ao.dense.weight, ao.dense.bias)) + x
            x = A.LayerNormPS.apply(a, ao.LayerNorm.weight, ao.LayerNorm.bias, ao.LayerNorm.eps)
            h = A.GeluF.apply(A.linear(x, lyr.intermediate.dense.weight, lyr.intermediate.dense.bias, out_planes=planes))
            o = drop(A.linear(h, lyr.output.dense.weight, lyr.output.dense.bias)) + x
            x = A.LayerNormPS.apply(o, lyr.output.LayerNorm.weight, lyr.output.LayerNorm.bias, lyr.output.LayerNorm.eps)
        return x

    def _backbone(self, img4, grid):
        bb = self.net.backbone
        if bb.pretrained_layout:
            r = bb.resnet
            x1 = A.MaxPoolF.apply(self._bn(A.StemF.apply(img4, r.conv1.weight), r.bn1, relu=True))
            for blk in r.layer1:
                x1 = self._tv_block(x1, blk)
            x2 = self._tv_block(x1, r.layer2[0])
            x2 = self._conv(torch.cat([x2, grid], -1), bb.early_fusion)
            for blk in list(r.layer2)[1:]:
                x2 = self._tv_block(x2, blk)
            x3 = x2
            for blk in r.layer3:
                x3 = self._tv_block(x3, blk)
            x4 = x3
            for blk in r.layer4:
                x4 = self._tv_block(x4, blk)
        else:
            x1 = A.MaxPoolF.apply(self._bn(A.StemF.apply(img4, bb.conv_1[0].weight), bb.conv_1[1], relu=True))
            for blk in bb.conv_2_x:
                x1 = self._our_block(x1, blk)
            x2 = self._our_block(x1, bb.conv_3_x.block_1)
            x2 = self._conv(torch.cat([x2, grid], -1), bb.conv_3_x.early_fusion)
            for blk in bb.conv_3_x.layers:
                x2 = self._our_block(x2, blk)
            x3 = x2
            for blk in bb.conv_4_x:
                x3 = self._our_block(x3, blk)
            x4 = x3
            for blk in bb.conv_5_x:
                x4 = self._our_block(x4, blk)
        # FPN top-down (no BN / bias / activation)
        x4 = self._conv(x4, bb.conv_6_x)
        x5 = self._conv(self._conv(x3, bb.skip_1) + A.Up2F.apply(x4), bb.merge_1)
        x6 = self._conv(self._conv(x2, bb.skip_2) + A.Up2F.apply(x5), bb.merge_2)
        x7 = self._conv(self._conv(x1, bb.skip_3) + A.Up2F.apply(x6), bb.merge_3)
        # fuse(cat[up8 x4, up4 x5, up2 x6, x7]) == chained K-slices of the fuse weight at native resolution (engine.py)
        wf = bb.fuse.weight.view(bb.fuse.out_channels, -1)
        Pc = wf.shape[1] // 4
        t = None
        for i, lvl in enumerate((x4, x5, x6, x7)):
            B, H, W, Cc = lvl.shape
            y = A.linear(lvl.reshape(B * H * W, Cc), wf[:, i * Pc:(i + 1) * Pc], None).view(B, H, W, -1)
            t = y if t is None else y + A.Up2F.apply(t)
        return t

    # ------------------------------------------------------------------ the step's forward
    def _graphable(self, dev):
        """Whole-step capture needs a step without host-side randomness inside the losses (index-sampled / OHEM losses draw
        with Python's ``random`` and upload indices, losses.py) and without collectives inside the tape (SyncBatchNorm)."""
        net = self.net
        cfg = net.loss_cfg
        if not self.use_graphs or dev.type != "cuda" or getattr(self, "_test_standins", False):
            return False
        if net.loss_weights is not None:
            return False
        if net.classifier_mode == "full":
            return False                     # the gated binary heads of `full` index rows by a device mask (losses.main_loss)
        if net.classifier_mode == "crf" and not self._two_stage_default(dev):
            return False                     # two-stage auxiliary head with sampled / OHEM settings: boolean-mask gathers
        if self._sampled_losses() and not self._device_sampling(dev):
            return False                     # host-side draws (Python `random`) and data-dependent shapes cannot be captured
        if any(self._sync_group(m) is not None for m in net.modules() if isinstance(m, nn.modules.batchnorm._BatchNorm)):
            return False
        if any(m.momentum is None for m in net.modules() if isinstance(m, nn.modules.batchnorm._BatchNorm)):
            return False                     # cumulative moving average: the factor is read back from the step counter
        return True

    def _two_stage_default(self, dev):
        """The two-stage auxiliary head (``full`` / ``crf``) with the plain-mean loss configuration on a GPU: computed by the
        fixed-shape form losses_device.two_stage_aux_default (no host sync); ``net.loss_sampling = "host"`` keeps losses.py."""
        net, cfg = self.net, self.net.loss_cfg
        return (dev.type == "cuda" and not getattr(self, "_test_standins", False) and net.classifier_mode in ("full", "crf")
                and cfg["aux_sample_list"] is None and tuple(int(v) for v in cfg["aux"]) == (-1, -1) and net.loss_weights is None
                and getattr(net, "loss_sampling", "device") == "device")

    def _sampled_losses(self):
        cfg = self.net.loss_cfg
        return not (cfg["aux_sample_list"] is None and tuple(cfg["aux"]) == (-1, -1)
                    and all(int(v) == -1 or int(v) >= (1 << 20) for v in tuple(cfg["main_1"]) + tuple(cfg["main_2"])))

    def _device_sampling(self, dev):
        """Sampled / OHEM losses drawn on the device (losses_device.py: no host randomness, no syncs, capturable) -- the default
        for the `simp` head on a GPU; ``net.loss_sampling = "host"`` keeps the reference's own Python-``random`` draws."""
        from . import losses_device
        net = self.net
        return (dev.type == "cuda" and not getattr(self, "_test_standins", False) and net.classifier_mode == "simp"
                and getattr(net, "loss_sampling", "device") == "device" and losses_device.supported(net.loss_cfg))

    def loss(self, image, seg_indices, seg_classes, coors, corpus, mask):
        net = self.net
        dev = corpus.device
        if dev.type != "cuda" and not getattr(self, "_test_standins", False):      # the flag exists for tests/mock_autograd.py only
            raise RuntimeError("ViBERTgridNet (B200) trains on CUDA tensors only; there is no CPU fallback")
        if net.classifier_mode not in ("simp", "full", "crf"):
            raise ValueError(f"unknown classifier_mode {net.classifier_mode!r}")
        sizes = list(net.image_min_size)
        # transform.py:124-131,192-194: one draw per image from the training min-size list
        min_sizes = [float(sizes[int(torch.empty(1).uniform_(0.0, float(len(sizes))).item())]) for _ in image]
        standins = getattr(self, "_test_standins", False)
        plan = plan_batch([ops.image_hw(im) if not standins else tuple(im.shape[-2:]) for im in image], [int(s.shape[0]) for s in seg_indices],
                          [int(c.shape[0]) for c in coors], int(corpus.shape[1]), min_sizes, float(net.image_max_size))
        # pinned staging + asynchronous copy: a pageable H2D copy would hold the host until the stream has drained, i.e. until
        # the previous step's kernels are done -- in a launch-bound step that bubble is paid in full
        tab = torch.from_numpy(plan.table)
        tab = (tab.pin_memory() if dev.type == "cuda" else tab).to(dev, non_blocking=True)
        st = dict(image=[im.contiguous() for im in image],
                  coors=torch.cat([c.reshape(-1, 4) for c in coors], 0).to(torch.int64).contiguous(),
                  seg_ids=torch.cat([s.reshape(-1) for s in seg_indices], 0).to(torch.int32).contiguous(),
                  cls=torch.cat([c.reshape(-1) for c in seg_classes], 0).to(torch.int32).contiguous(),
                  corpus=corpus.contiguous(), mask=None if mask is None else mask.to(torch.int32).contiguous(), tab=tab)
        if self._graphable(dev):
            return self._graphed_loss(plan, st, tuple(min_sizes))
        return self._eager(plan, st)

    def _seed_word(self):
        """The device seed word of graphed steps, or None on the eager path (by-value seeds).  The tensor itself lives as long
        as the engine: every captured graph has its ADDRESS baked in, so it must never be dropped and re-created."""
        return self._step_seed if self._graph_step else None

    def _eager(self, plan, st):
        self._graph_step = False
        self.arena_fresh = False
        return self._forward(plan, st)

    # ------------------------------------------------------------------ whole-step CUDA graph
    def _graphed_loss(self, plan, st, min_sizes):
        net = self.net
        dev = st["corpus"].device
        named = [(n, p) for n, p in net.named_parameters() if p.requires_grad]
        params = [p for _, p in named]
        key = (tuple((tuple(im.shape), im.dtype) for im in st["image"]), tuple(st["coors"].shape), tuple(st["seg_ids"].shape), tuple(st["corpus"].shape),
               st["mask"] is None, min_sizes, dev.index, tuple(p.data_ptr() for p in params),
               tuple(b.data_ptr() for b in net.buffers()), float(net.bert_hidden_dropout), float(net.bert_attn_dropout),
               tuple(plan.seg_counts), tuple(int(v) for v in plan.view("tok_off")))
        ent = self._graphs.get(key)
        if ent is None:                        # first sighting: an ordinary eager step (also warms lazy kernel attributes up)
            self._graphs[key] = {"graph": None}
            if len(self._graphs) > self.max_graphs:
                self._graphs.pop(next(iter(self._graphs)))
            return self._eager(plan, st)
        if ent.get("never"):
            return self._eager(plan, st)
        if self._step_seed is None or self._step_seed.device != dev:
            self._step_seed = torch.zeros(1, dtype=torch.int64, device=dev)
            # pinned staging ring: the host runs a step or more ahead of the device, so the word of step t must not be
            # rewritten before its (asynchronous) upload has executed -- each slot is guarded by an event
            self._seed_ring = [[torch.zeros(1, dtype=torch.int64).pin_memory(), None] for _ in range(8)]
            self._seed_i = 0
        if ent["graph"] is None:               # second sighting: capture forward + backward
            try:
                self._capture(ent, plan, st, named, params, dev)
            except RuntimeError as e:
                # A capture can be invalidated from outside the step (any thread's synchronising CUDA call while the stream
                # records): never fatal -- this step runs eagerly and the capture is tried again at the next sighting.
                if "capture" not in str(e).lower():
                    raise
                for cleanup in (A.side_join, A.prep_end):
                    try:
                        cleanup()
                    except RuntimeError:
                        pass
                torch.cuda.synchronize()
                self.capture_failures += 1
                ent["failed"] = ent.get("failed", 0) + 1
                if ent["failed"] >= 3:
                    ent["never"] = True
                import warnings
                warnings.warn(f"[vibertgrid_b200] whole-step capture failed ({type(e).__name__}: {str(e)[:120]}); eager step, "
                              + ("capture disabled for this signature" if ent.get("never") else "will retry"))
                return self._eager(plan, st)
        else:
            sd = ent["static"]
            for dst, src in zip(sd["image"], st["image"]):
                dst.copy_(src, non_blocking=True)
            for k in ("coors", "seg_ids", "cls", "corpus", "mask"):
                if sd[k] is not None:
                    sd[k].copy_(st[k], non_blocking=True)
        self._seed_i = (self._seed_i + 1) % len(self._seed_ring)
        slot = self._seed_ring[self._seed_i]
        if slot[1] is not None:
            slot[1].synchronize()
        slot[0].random_(0, 2 ** 62)
        self._step_seed.copy_(slot[0], non_blocking=True)
        slot[1] = torch.cuda.Event()
        slot[1].record()
        self._graph_step = True
        ent["graph"].replay()
        self.graph_replays += 1
        self.kernel_launches += ent["launches"]
        self.last = ent["last"]
        self.grad_arena, self.arena_fresh = ent["arena"], True
        return _StepGradsF.apply(ent["loss"], len(ent["views"]), *ent["views"], *ent["params"])

    def _capture(self, ent, plan, st, named, params, dev):
        net = self.net
        static = {k: ([t.clone() for t in v] if isinstance(v, list) else (None if v is None else v.clone())) for k, v in st.items()}
        if self._side_streams is None:         # the two helper streams of the captured step, created once, outside any capture
            self._side_streams = (torch.cuda.Stream(dev), torch.cuda.Stream(dev))
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        c0 = ops.L.launch_count
        self._graph_step = True
        with torch.cuda.graph(g):
            # The tape is built over ALIASES of the parameters (detached views of the same storage, made leaves inside the
            # capture): the real Parameters' AccumulateGrad nodes were created by earlier eager steps / by DDP on the default
            # stream and stay alive, and routing a captured backward through them makes the engine synchronise the capturing
            # stream with the legacy stream, which invalidates the capture.  The aliases read the parameter storage, so replays
            # follow every in-place optimizer update.
            aliases = {id(p_): p_.detach().requires_grad_() for _, p_ in named}
            with _parameters_as(net, aliases):
                with torch.enable_grad():
                    # parameter-only preparation (weight planes, transposes, repacks) runs ahead on a prep stream
                    # (autograd.py "_PrepWork"); joined after the backward, when its last consumer has been enqueued
                    if self.side_prep:
                        A.prep_begin(dev, self._side_streams[0])
                    try:
                        # weight gradients / bias sums / table gradients (backward) and the BatchNorm running-statistics
                        # updates (forward) run on a side stream beside the main chain and are joined once, below
                        # (autograd.py "_SideWork")
                        if self.side_wgrad:
                            A.side_begin(dev, self._side_streams[1])
                        try:
                            loss = self._forward(plan, static)
                            grads = torch.autograd.grad(loss, [aliases[id(p_)] for p_ in params], allow_unused=True)
                        finally:
                            A.side_join()
                    finally:
                        A.prep_end()
            used = [(p_, g_) for p_, g_ in zip(params, grads) if g_ is not None]
            # one flat arena, gradients as views (the N > 1 bench averages it with in-place all-reduces over its slices)
            arena = torch.empty(sum(g_.numel() for _, g_ in used), dtype=torch.float32, device=dev)
            views, off = [], 0
            for _, g_ in used:
                views.append(arena[off:off + g_.numel()].view(g_.shape))
                off += g_.numel()
            torch._foreach_copy_(views, [g_ for _, g_ in used])
            loss_out = loss.detach().clone()
        ent.update(graph=g, static=static, loss=loss_out, views=views, arena=arena, params=[p_ for p_, _ in used], last=self.last,
                   launches=ops.L.launch_count - c0)       # C-ABI kernel launches recorded into the graph = executed per replay

    def _forward(self, plan, st):
        """The kernel sequence of one training forward over staged inputs (capturable: no host sync)."""
        net = self.net
        image, coors_cat, seg_ids, cls_cat, corpus, mask, tab = (st["image"], st["coors"], st["seg_ids"], st["cls"], st["corpus"],
                                                                 st["mask"], st["tab"])
        dev = corpus.device
        default_aux = (net.classifier_mode == "simp" and net.loss_cfg["aux_sample_list"] is None
                       and tuple(net.loss_cfg["aux"]) == (-1, -1) and net.loss_weights is None)
        dt = {k: tab[s:s + n] for k, (s, n) in plan.offsets.items()}
        dt["ratios"] = dt["ratios"].view(torch.float32)
        seg_off = dt["seg_off"]
        B = plan.B
        status = torch.zeros(1, dtype=torch.int32, device=dev)

        # a1 transform
        boxes = ops.resize_coords(coors_cat, seg_off, dt["ratios"], B)
        img4 = torch.zeros((B, plan.H + 6, plan.W + 6, 4), dtype=torch.float32, device=dev)
        for b, im in enumerate(image):
            ops.normalize_resize_pad(im, img4, b, plan.sizes[b][0], plan.sizes[b][1], net.image_mean, net.image_std)

        # a2 - a4 BERT, segment aggregation, BERTgrid
        hidden = self._bert(plan, dt, corpus)
        if mask is not None and dev.type == "cuda":                 # input contract of `mask` (prefix of n_tok ones): status bit 2
            ops.mask_check(mask, dt["tok_off"], status)
        seg_start = ops.segment_starts(seg_ids, dt["tok_off"], B, plan.K, status)
        seg_emb = A.SegmentReduceF.apply(hidden, dt["tok_row"], seg_start, plan.K,
                                         ops.AGG_MEAN if net.grid_mode == "mean" else ops.AGG_FIRST)
        gs = net.early_fusion_downsampling_ratio
        idx = ops.box_index_map(boxes, seg_off, B, gs, int(plan.H / gs), int(plan.W / gs))
        grid = A.GridScatterF.apply(seg_emb, idx, boxes, seg_off, B, gs)

        # a5 backbone
        p_fuse = self._backbone(img4, grid)

        # a6 auxiliary segmentation head: conv-BN-ReLU x2, the packed 1x1 heads at low resolution, fused CE
        enc = net.semantic_segmentation_head.encoder
        s = self._bn(self._conv(p_fuse, enc.conv_1), enc.bn_1, relu=True)
        s = self._bn(self._conv(s, enc.conv_2), enc.bn_2, relu=True)
        seg_w = torch.cat([enc.conv_3_1.weight.flatten(1), enc.conv_3_2.weight.flatten(1)], 0)
        seg_b = torch.cat([enc.conv_3_1.bias, enc.conv_3_2.bias], 0)
        Bs, Hs, Ws, Cs = s.shape
        lg = A.linear(s.reshape(Bs * Hs * Ws, Cs), seg_w, seg_b).view(Bs, Hs, Ws, -1)
        from . import losses
        ctx = None
        if self._sampled_losses() and self._device_sampling(dev):
            from .losses_device import SamplingCtx
            ctx = SamplingCtx(self._seed_word(), base_seed=_rand_seed() & 0xFFFFFFFF)
        if default_aux:
            aux = A.SegCEF.apply(lg, boxes, seg_off, cls_cat, B, plan.H, plan.W, net.p_fuse_downsampling_ratio, 3)
            loss_aux = aux[0] + aux[1]
        else:
            # semantic_segmentation_head.py:73-78 (nearest upsample to the padded image size), :199-214 (label painting),
            # :216-233 / :343-352 (sampled / OHEM losses, the binary second stage of the two-stage head)
            full = torch.nn.functional.interpolate(lg.permute(0, 3, 1, 2), scale_factor=int(net.p_fuse_downsampling_ratio),
                                                   mode="nearest")
            pos_neg_labels, class_labels = ops.label_paint(boxes, seg_off, cls_cat, B, plan.H, plan.W)
            if self._two_stage_default(dev):
                # full / crf heads with the plain-mean loss configuration: the fixed-shape, sync-free form (capturable)
                from .losses_device import two_stage_aux_default
                loss_aux = two_stage_aux_default(net, full[:, :3], full[:, 3:], pos_neg_labels, class_labels)
            else:
                loss_aux = losses.aux_loss(net, {"pred_mask": full[:, :3], "pred_ss": full[:, 3:],
                                                 "pos_neg_labels": pos_neg_labels, "class_labels": class_labels}, ctx)

        # a7 / a8 ROI align, late fusion
        roi = A.RoiAlignF.apply(p_fuse, boxes, seg_off, 1.0 / float(net.p_fuse_downsampling_ratio), net.roi_shape)
        rn = net.late_fusion_net.ROI_embedding_net
        r = self._bn(self._conv(roi, rn.conv_1), rn.bn_1, relu=True)
        r = self._bn(self._conv(r, rn.conv_2), rn.bn_2, relu=True)
        Cc, Pp = net.p_fuse_channel, net.roi_shape
        # the FC weight is stored for the (C, H, W) flatten of the reference; the kernels flatten (H, W, C)
        w_fc = rn.linear.weight.view(-1, Cc, Pp, Pp).permute(0, 2, 3, 1).reshape(rn.linear.out_features, -1)
        roi_emb = A.linear(r.reshape(plan.K, -1), w_fc, rn.linear.bias)
        fl = net.late_fusion_net.fuse_embedding_net.linear
        late = A.linear(torch.cat([roi_emb, seg_emb], 1), fl.weight, fl.bias)

        # a9 field-type head + losses (tiny [K, C] tensors: torch ops, see losses.py)
        head = net.field_type_classification_head

        def mlp(m, x):
            if not hasattr(m, "linear_1"):                # layer_mode "single"
                return A.linear(x, m.linear.weight, m.linear.bias)
            h = torch.relu(A.linear(x, m.linear_1.weight, m.linear_1.bias))
            return A.linear(h, m.linear_2.weight, m.linear_2.bias)

        out = {"gt_label": cls_cat, "plan": plan}
        if net.classifier_mode == "simp":
            out["logits"] = mlp(head.category_classification_net, late)
            if hasattr(head, "pos_neg_classification_net"):
                out["pos_neg_logits"] = mlp(head.pos_neg_classification_net, late)
            loss_c = losses.main_loss(net, out, ctx)
        elif net.classifier_mode == "crf":
            # field_type_classification_head.py:683-699: emissions, then the mean over samples of the per-sample NLL
            feats = mlp(head.category_classification_net, late)
            nll = A.CrfNllF.apply(feats, head.crf_layer.transitions, cls_cat, seg_off, B)
            loss_c = (nll.sum() / B).reshape(1)
        else:
            # field_type_classification_head.py:355-400: the binary heads run on every row here; main_loss gates the rows
            # with the detached pos/neg decision, which leaves the same loss and gradients as running them on the gated rows
            out["pos_neg_logits"] = mlp(head.pos_neg_classification_net.layer, late)
            out["logits"] = torch.cat([mlp(getattr(head, f"category_classification_net_{i}").layer, late)
                                       for i in range(net.num_tokens - 1)], 1)
            loss_c = losses.main_loss(net, out)
        self.last = dict(status=status, loss_aux=loss_aux.detach(), loss_c=loss_c.detach())
        return loss_c + net.loss_control_lambda * loss_aux
