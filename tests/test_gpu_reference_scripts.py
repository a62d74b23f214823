"""The reference's OWN entry scripts, unmodified, driving the drop-in on the GPU (BASELINE configs[0] is "via
reference eval_SROIE.py").  The script files come from the staged copy ``oracle/_ref/reference`` (oracle/stage_reference.py;
git-ignored, travels to the GPU box); ``dropin/`` precedes it on PYTHONPATH so ``from model.ViBERTgrid_net import
ViBERTgridNet`` (eval_SROIE.py:11, train_SROIE.py:13) resolves to the B200 module while ``data.*`` / ``pipeline.*`` stay the
reference's.  tests/harness/run_reference_script.py is the launcher (its docstring lists the three harness-side patches).

  * eval_SROIE.main over a synthetic on-disk SROIE tree + checkpoint: the result file the script writes equals the one the
    live reference wrote on CPU (tests/golden/eval_sroie_cfg1.json, oracle/make_script_golden.py), and every document's
    ``pred_label`` is within 1e-3 of the reference's with identical argmax.
  * train_SROIE.train under torchrun with ``syncBN: True`` and ``amp: True``: SyncBatchNorm conversion, DDP
    (find_unused_parameters), autocast, GradScaler, SGD + AdamW, validate() before and after -- one rank here; two ranks
    (when the box has two GPUs) must leave bit-identical parameters on both ranks and reproduce the single-rank losses.
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR, ROOT

sys.path.insert(0, os.path.join(ROOT, "tests", "harness"))

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not os.path.isfile(os.path.join(ROOT, "oracle", "_ref", "reference", "eval_SROIE.py")),
                                 reason="staged reference absent (python oracle/stage_reference.py in the build container)")]
LAUNCHER = os.path.join(ROOT, "tests", "harness", "run_reference_script.py")


def _harness_lines(stdout):
    return [json.loads(l[len("VBG_HARNESS "):]) for l in stdout.splitlines() if l.startswith("VBG_HARNESS ")]


def test_staged_reference_is_the_reference():
    import hashlib
    man = json.load(open(os.path.join(ROOT, "oracle", "_ref", "MANIFEST.json")))
    for rel in ("eval_SROIE.py", "train_SROIE.py", "pipeline/train_val_utils.py", "data/SROIE_dataset.py", "model/ViBERTgrid_net.py"):
        got = hashlib.sha256(open(os.path.join(ROOT, "oracle", "_ref", "reference", rel), "rb").read()).hexdigest()
        assert got == man["files"][rel], rel
        live = os.path.join("/root/reference", rel)
        if os.path.isfile(live):
            assert got == hashlib.sha256(open(live, "rb").read()).hexdigest(), rel


@pytest.mark.parametrize("precision", ["bf16x3", "fp32"])
def test_eval_sroie_main_through_the_dropin(precision, tmp_path):
    import sroie_synth
    tmp = str(tmp_path)
    cpath, rpath = sroie_synth.prepare_eval_case(tmp, "cuda")
    env = sroie_synth.script_env(ROOT, with_dropin=True)
    env["VBG_HARNESS_DUMP"] = os.path.join(tmp, "preds.npz")
    env["VBG_PRECISION"] = precision
    out = subprocess.run([sys.executable, LAUNCHER, "eval_SROIE", "--config", cpath], cwd=tmp, env=env, capture_output=True,
                         text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    info = _harness_lines(out.stdout)[-1]
    # what ran: the drop-in's module, the in-tree sm_100a library, kernels launched through the C-ABI
    assert info["net_module"] == "vibertgrid_pytorch_b200.net", info
    assert info["model_module_file"].startswith(os.path.join(ROOT, "dropin")), info
    assert info["so"] and info["so"].endswith("libvbg_sm100a.so") and info["launches"] > 100, info   # host-side C-ABI calls (CUDA-graph replays re-run them without re-counting)
    got = json.load(open(rpath))
    want = json.load(open(os.path.join(GOLDEN_DIR, "eval_sroie_cfg1.json")))
    z, zr = np.load(env["VBG_HARNESS_DUMP"]), np.load(os.path.join(GOLDEN_DIR, "eval_sroie_cfg1_preds.npz"))
    assert int(z["n"]) == int(zr["n"]) == sroie_synth.EVAL_DOCS
    worst, flips = 0.0, 0
    for i in range(int(z["n"])):
        p, pr = z[f"pred_{i}"], zr[f"pred_{i}"]
        assert p.shape == pr.shape
        worst = max(worst, float(np.abs(p - pr).max() / np.abs(pr).max()))
        flips += int((p.argmax(1) != pr.argmax(1)).sum())
        assert abs(float(z[f"loss_{i}"][0]) - float(zr[f"loss_{i}"][0])) <= 1e-3 * abs(float(zr[f"loss_{i}"][0]))
    print(f"[eval_SROIE.py through the drop-in, {precision}] pred_label max-rel {worst:.1e} vs the reference on CPU, "
          f"{flips} argmax flips over {int(z['n'])} documents, {info['launches']} kernel launches")
    assert worst < (1e-3 if precision == "bf16x3" else 1e-4) and flips == 0
    assert got == want                          # the file the script itself writes: per-document key strings + metrics


def _train_case(tmp, batch_size, n_train=12, sync_bn=True):
    import dataclasses
    import sroie_synth
    from vibertgrid_pytorch_b200 import synth
    cfg = dataclasses.replace(synth.CONFIGS["mid"], name="train_case", batch=batch_size, height=256, width=256, segments=16,
                              ragged=False)
    sroie_synth.write_split(os.path.join(tmp, "data", "train"), n_train, cfg, seed=2, tokens_per_seg=3)
    sroie_synth.write_split(os.path.join(tmp, "data", "test"), 2, cfg, seed=3, tokens_per_seg=3)
    sroie_synth.write_bert(cfg, tmp, with_weights=True, seed=0, dropout=0.0)     # dropout off: runs must be comparable
    cpath = os.path.join(tmp, "train.yaml")
    sroie_synth.write_config(cpath, cfg, os.path.join(tmp, "data"), batch_size=batch_size, end_epoch=1, sync_bn=sync_bn, amp=True)
    return cpath


def _run_train(tmp, nproc, batch_size, port, sync_bn=True):
    import sroie_synth
    cpath = _train_case(tmp, batch_size, sync_bn=sync_bn)
    env = sroie_synth.script_env(ROOT, with_dropin=True)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), LAUNCHER, "train_SROIE", "-c", cpath]
    out = subprocess.run(cmd, cwd=tmp, env=env, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-5000:]
    infos = sorted(_harness_lines(out.stdout), key=lambda i: i["rank"])
    assert len(infos) == nproc
    return infos, out.stdout


def test_train_sroie_one_rank_ddp_syncbn_amp(tmp_path):
    infos, stdout = _run_train(str(tmp_path), 1, 4, 29611)
    info = infos[0]
    assert info["net_module"] == "vibertgrid_pytorch_b200.net" and info["launches"] > 1000, info
    assert info["bn_classes"] == ["SyncBatchNorm"], info                # train_SROIE.py:203-205 converted every BatchNorm
    assert len(info["losses"]) == 3 and all(np.isfinite(info["losses"])), info     # 12 documents / batch 4
    assert info["sgd_steps"] == 3 and info["adamw_steps"] == 3, info    # GradScaler found finite gradients and stepped both
    assert info["params_with_grad"] > 100
    assert "train_loss" in stdout and "validate_loss" in stdout          # train_one_epoch and validate() both printed
    # one rank: SyncBatchNorm falls back to per-rank statistics (torch's own rule), so the step has no collective inside the tape
    # and is captured -- under DDP, autocast and GradScaler: step 0 eager, step 1 capture + replay, step 2 replay
    assert info["train_graph_replays"] >= 2, info
    print(f"[train_SROIE.py, 1 rank, syncBN + amp] losses {info['losses']}")


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_train_sroie_two_ranks_match_one_rank(tmp_path):
    one, _ = _run_train(str(tmp_path / "w1"), 1, 4, 29612)
    two, _ = _run_train(str(tmp_path / "w2"), 2, 2, 29613)
    a, b = two
    assert a["checksums_all_ranks"][0] == a["checksums_all_ranks"][1], a     # DDP kept the replicas bit-identical
    assert a["sgd_steps"] == b["sgd_steps"] == 3 and a["adamw_steps"] == 3
    l2 = np.mean(np.asarray(a["losses_all_ranks"], np.float64), 0)           # same 4 documents per step as the one-rank run
    l1 = np.asarray(one[0]["losses"], np.float64)
    print(f"[train_SROIE.py] one rank x batch 4: {l1.tolist()}  two ranks x batch 2 (mean over ranks): {l2.tolist()}")
    assert np.allclose(l1, l2, rtol=2e-3), (l1, l2)
    assert abs(one[0]["param_checksum"] - a["param_checksum"]) <= 1e-5 * abs(a["param_checksum"])


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_train_sroie_two_ranks_graphed_step_under_ddp(tmp_path):
    """``syncBN: False``: no collective inside the tape, so each rank replays its whole-step CUDA graph and DDP's bucket hooks fire
    from the step's one autograd node -- replicas stay bit-identical and the losses equal the one-rank run of the same global
    batch up to the per-rank BatchNorm statistics (two documents per rank instead of four)."""
    two, _ = _run_train(str(tmp_path / "w2"), 2, 2, 29615, sync_bn=False)
    a, b = two
    assert a["bn_classes"] == ["BatchNorm2d"], a
    assert a["train_graph_replays"] >= 2 and b["train_graph_replays"] >= 2, (a["train_graph_replays"], b["train_graph_replays"])
    assert a["checksums_all_ranks"][0] == a["checksums_all_ranks"][1], a     # DDP kept the replicas bit-identical
    assert a["sgd_steps"] == b["sgd_steps"] == 3 and a["adamw_steps"] == 3
    assert all(np.isfinite(a["losses"])) and all(np.isfinite(b["losses"]))
    print(f"[train_SROIE.py, 2 ranks, graphed step under DDP] losses rank0 {a['losses']} rank1 {b['losses']}")
