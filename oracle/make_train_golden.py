"""Generate ``tests/golden/train_*.npz``: one training step (loss, gradients, BatchNorm running statistics) of the UNMODIFIED
reference (/root/reference, CPU, fp32, ``model.train()``) on seeded synthetic documents and the seeded weights of
``vibertgrid_pytorch_b200.synth``.  All nn.Dropout probabilities are set to 0 on the reference instance (dropout masks are not
comparable across implementations); everything else is the stock training forward + ``loss.backward()``
(pipeline/train_val_utils.py:265-277).

Run in the build container only:   python oracle/make_train_golden.py

Per parameter the fixture stores (sum, L2 norm, 64 evenly strided samples) of the gradient -- enough to catch any wrong or
missing gradient without storing 45M floats.
"""
from __future__ import annotations

import dataclasses
import json
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, ROOT)

from vibertgrid_pytorch_b200 import synth  # noqa: E402
from vibertgrid_pytorch_b200.net import ViBERTgridNet as OurNet  # noqa: E402  (parameter tree only)
from oracle.make_golden import seed_hub, write_bert_dir  # noqa: E402

CASES = {
    # fixture -> (config, weight seed, input seed[, classifier mode, extra constructor kwargs])
    "train_tiny": ("tiny", 0, 0),
    "train_tiny_d": ("tiny_d", 3, 3),
    "train_tiny_pre": ("tiny_pre", 4, 4),
    "train_mid": ("mid", 6, 6),                 # larger BatchNorm populations: the better-conditioned case
    # the other heads / loss configurations of the training step
    "train_tiny_crf": ("tiny", 7, 7, "crf", {}),                                  # CRF negative log-likelihood
    "train_tiny_crf_multi": ("tiny_d", 8, 8, "crf", {"layer_mode": "multi"}),
    "train_tiny_full": ("tiny", 9, 9, "full", {}),                                # two-stage heads, single-layer classifiers
    "train_tiny_full_multi": ("tiny_d", 10, 10, "full", {"layer_mode": "multi"}),
    # simp head with the rounding-robust loss knobs on: index-sampled aux-1 (Python `random` draws), class weights
    "train_tiny_sampled": ("tiny", 11, 11, "simp", {"loss_aux_sample_list": [300, 200, 100],
                                                    "loss_weights": [0.5, 1.0, 2.0, 1.5, 0.75]}),
    # OHEM everywhere.  The reference indexes the SORTED losses with ORIGINAL indices (custom_loss.py:174-176), which makes
    # the kept set -- and the loss -- jump under 1e-6 perturbations of tied pixel losses, so these two bind the host logic
    # only (tests/test_train_host_logic.py, same fp32 CPU arithmetic as the reference), not the GPU kernels.
    "train_tiny_ohem": ("tiny", 11, 11, "simp", {"loss_aux_sample_list": [300, 200, 100], "num_hard_positive_aux": 150,
                                                 "num_hard_negative_aux": 250, "num_hard_positive_main_1": 3,
                                                 "num_hard_negative_main_1": 2, "num_hard_positive_main_2": 4,
                                                 "num_hard_negative_main_2": 2,
                                                 "loss_weights": [0.5, 1.0, 2.0, 1.5, 0.75]}),
    "train_tiny_full_ohem": ("tiny_d", 10, 10, "full", {"layer_mode": "multi", "num_hard_positive_main_2": 2,
                                                        "num_hard_negative_main_2": 3}),
}
PY_RANDOM_SEED = 4321
N_SAMPLES = 64


def summarize(t):
    f = t.detach().double().reshape(-1)
    idx = torch.linspace(0, f.numel() - 1, min(N_SAMPLES, f.numel())).long()
    return np.concatenate([[float(f.sum()), float(f.norm())], f[idx].numpy()])


def run_one(name, cfg_name, wseed, iseed, mode="simp", extra=None, outdir=None):
    import random
    extra = dict(extra or {})
    cfg = dataclasses.replace(synth.CONFIGS[cfg_name], classifier_mode=mode)
    if mode == "crf" and cfg.tag_to_idx is None:
        cfg.tag_to_idx = {f"T{i}": i for i in range(cfg.num_classes)}
    with tempfile.TemporaryDirectory() as tmp:
        cwd = os.getcwd()
        os.chdir(tmp)
        try:
            write_bert_dir(cfg, tmp)
            seed_hub(tmp)
            kw = {**synth.model_kwargs(cfg, "eval"), **extra}
            ours = OurNet(**kw)
            synth.fill_state_dict_(ours, wseed)
            sd = {k: v.clone() for k, v in ours.state_dict().items()}
            sys.path.insert(0, REF)
            for m in [m for m in sys.modules if m.split(".")[0] in ("model", "pipeline")]:
                del sys.modules[m]
            from model.ViBERTgrid_net import ViBERTgridNet as RefNet
            ref = RefNet(**{**synth.model_kwargs(cfg, "eval"), **extra})
            ref.load_state_dict(sd, strict=True)
            sys.path.remove(REF)
            ref.train()
            n_drop = 0
            for m in ref.modules():
                if isinstance(m, torch.nn.Dropout):
                    m.p = 0.0
                    n_drop += 1
            for c in [getattr(m, "config", None) for m in ref.modules()]:
                if c is not None and hasattr(c, "attention_probs_dropout_prob"):
                    c.attention_probs_dropout_prob = 0.0
                    c.hidden_dropout_prob = 0.0
            batch = synth.make_batch(cfg, iseed)
            torch.manual_seed(0)
            random.seed(PY_RANDOM_SEED)
            loss = ref(*batch)
            assert isinstance(loss, torch.Tensor), "training mode must return the loss alone"
            loss.backward()
            bufs = {k: summarize(v.float()) for k, v in ref.named_buffers() if "running_" in k or "num_batches" in k}
            # determinism check (dropout really off): a second forward gives the same loss (BN running stats do not enter)
            torch.manual_seed(1)
            random.seed(PY_RANDOM_SEED)
            loss2 = ref(*batch)
            assert abs(float(loss2) - float(loss)) < 1e-6 * max(1.0, abs(float(loss))), (float(loss), float(loss2))
        finally:
            os.chdir(cwd)
    fx = dict(meta=json.dumps(dict(name=name, cfg=cfg_name, classifier_mode=mode, extra_kwargs=extra, py_random_seed=PY_RANDOM_SEED,
                                   weight_seed=wseed, input_seed=iseed,
                                   torch=torch.__version__, dropouts_zeroed=n_drop)),
              loss=np.asarray([float(loss)]), loss_shape=np.asarray(list(loss.shape), dtype=np.int64))
    names, no_grad = [], []
    for k, p in ref.named_parameters():
        if p.grad is None:
            no_grad.append(k)
            continue
        names.append(k)
        fx["g:" + k] = summarize(p.grad)
    for k, v in bufs.items():                 # BatchNorm running statistics after ONE training forward
        fx["b:" + k] = v
    fx["buffer_names"] = np.asarray(list(bufs))
    fx["grad_names"] = np.asarray(names)
    fx["no_grad_names"] = np.asarray(no_grad)
    path = os.path.join(outdir, f"{name}.npz")
    np.savez_compressed(path, **fx)
    print(f"[{name}] loss={float(loss):.6f} params with grad={len(names)} without={len(no_grad)} -> {path} "
          f"({os.path.getsize(path) / 1024:.0f} KiB)")


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count() or 1)
    out = os.path.join(ROOT, "tests", "golden")
    for n in (sys.argv[1:] or list(CASES)):
        run_one(n, *CASES[n], outdir=out)
