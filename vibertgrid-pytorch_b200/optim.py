"""Fused multi-tensor optimizers for the training step: drop-in replacements for the two optimizers the reference builds
(train_SROIE.py:217-235) -- ``torch.optim.SGD`` (momentum, weight decay) for the CNN / heads and ``torch.optim.AdamW`` for the
parameters whose name contains ``bert_model`` -- that update EVERY tensor of a param group in one sm_100a kernel launch
(``vbg_sgd_step_mt`` / ``vbg_adamw_step_mt``) instead of torch's multi-pass foreach loops.

Same constructor arguments and ``param_groups`` keys, so the reference's per-iteration lr / weight-decay schedulers
(pipeline/train_val_utils.py:211-243 write ``param_group["lr"]`` / ``["weight_decay"]``) and ``GradScaler.step`` work
unchanged; same update rules and state names (``momentum_buffer``, ``exp_avg``, ``exp_avg_sq``, ``step``), so a checkpoint's
optimizer state loads into either.  No CPU path: parameters must live on a CUDA device.
"""
from __future__ import annotations

import math

import torch

from . import _lib as L
from .ops import _stream


class _FusedBase(torch.optim.Optimizer):
    def _table(self, rows, device):
        """Device table of {param, grad, state1, state2, numel, first_chunk} (6 int64 per tensor), staged through a pinned
        buffer: gradient tensors are fresh every step, so the table is rebuilt per step (a few hundred pointers)."""
        chunk = L.load().vbg_optim_chunk()
        n = len(rows)
        # Pinned staging: a RING of host tables, each guarded by an event recorded after its asynchronous upload.  The host runs
        # ahead of the device (a whole training step can be queued in front of the copy), so a single buffer would be rewritten
        # with the NEXT call's pointers before the previous upload has executed (found by running the suite under
        # compute-sanitizer, which slows the device down enough to expose it).
        ring = getattr(self, "_host_ring", None)
        if ring is None:
            ring = self._host_ring = {"slots": [None] * 4, "i": 0}
        ring["i"] = (ring["i"] + 1) % len(ring["slots"])
        slot_h = ring["slots"][ring["i"]]
        if slot_h is None or slot_h["buf"].shape[0] < n:
            slot_h = ring["slots"][ring["i"]] = {"buf": torch.empty((max(n, 64), 6), dtype=torch.int64).pin_memory(), "ev": None}
        if slot_h["ev"] is not None:
            slot_h["ev"].synchronize()               # the upload that last read this buffer has completed
        host = slot_h["buf"]
        dev = getattr(self, "_dev_tabs", None)
        if dev is None:
            dev = self._dev_tabs = {}
        flat, first = [], 0
        for p, g, s1, s2 in rows:
            flat += [p.data_ptr(), g.data_ptr(), s1.data_ptr() if s1 is not None else 0, s2.data_ptr() if s2 is not None else 0, p.numel(), first]
            first += -(-p.numel() // chunk)
        host[:n].view(-1).copy_(torch.tensor(flat, dtype=torch.int64))
        key = (device, n)
        # two device tables used alternately: the copy of step t + 1 never overwrites what the kernel of step t may still read
        slot = self._dev_tabs.setdefault(key, [torch.empty((n, 6), dtype=torch.int64, device=device) for _ in range(2)] + [0])
        slot[2] ^= 1
        tab = slot[slot[2]]
        tab.copy_(host[:n], non_blocking=True)
        slot_h["ev"] = torch.cuda.Event()
        slot_h["ev"].record()
        return tab, first

    @staticmethod
    def _rows(group):
        for p in group["params"]:
            if p.grad is None:
                continue
            if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError("fused optimizers update contiguous fp32 CUDA parameters only (no CPU path)")
            g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
            if g.is_sparse or g.dtype != torch.float32:
                raise RuntimeError("fused optimizers need dense fp32 gradients")
            yield p, g


class FusedSGD(_FusedBase):
    """``torch.optim.SGD(params, lr, momentum, weight_decay)`` (dampening 0, no Nesterov) in one launch per param group."""

    def __init__(self, params, lr=1e-3, momentum=0.0, weight_decay=0.0, grad_scale=1.0):
        super().__init__(params, dict(lr=lr, momentum=momentum, weight_decay=weight_decay, dampening=0, nesterov=False, grad_scale=grad_scale))

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        lib = L.load()
        for group in self.param_groups:
            fresh, old = [], []          # a momentum buffer created this step is INITIALISED to the gradient (torch's first step)
            for p, g in self._rows(group):
                st = self.state[p]
                if group["momentum"] != 0 and "momentum_buffer" not in st:
                    st["momentum_buffer"] = torch.empty_like(p)
                    fresh.append((p, g, st["momentum_buffer"], None))
                else:
                    old.append((p, g, st.get("momentum_buffer"), None))
            for sub, first in ((fresh, 1), (old, 0)):
                if not sub:
                    continue
                tab, chunks = self._table(sub, sub[0][0].device)
                L.check(lib.vbg_sgd_step_mt(tab.data_ptr(), len(sub), chunks, float(group["lr"]), float(group["momentum"]),
                                            float(group["weight_decay"]), first, float(group["grad_scale"]), _stream()), "vbg_sgd_step_mt")
        return loss


class FusedAdamW(_FusedBase):
    """``torch.optim.AdamW(params, lr, betas, eps, weight_decay)`` (amsgrad off) in one launch per param group."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, grad_scale=1.0):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=False, grad_scale=grad_scale))

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        lib = L.load()
        for group in self.param_groups:
            by_step = {}
            for p, g in self._rows(group):
                st = self.state[p]
                if "exp_avg" not in st:
                    st["step"], st["exp_avg"], st["exp_avg_sq"] = 0, torch.zeros_like(p), torch.zeros_like(p)
                st["step"] = int(st["step"]) + 1
                by_step.setdefault(st["step"], []).append((p, g, st["exp_avg"], st["exp_avg_sq"]))
            b1, b2 = group["betas"]
            for t, rows in by_step.items():          # one launch per distinct step count (normally exactly one)
                tab, chunks = self._table(rows, rows[0][0].device)
                L.check(lib.vbg_adamw_step_mt(tab.data_ptr(), len(rows), chunks, float(group["lr"]), float(b1), float(b2), float(group["eps"]),
                                              float(group["weight_decay"]), 1.0 - b1 ** t, math.sqrt(1.0 - b2 ** t), float(group["grad_scale"]),
                                              _stream()), "vbg_adamw_step_mt")
        return loss
