"""Test harness (NOT product code): run one of the reference's own entry scripts -- unmodified, from the staged tree
``oracle/_ref/reference`` -- in this process, with whichever ``model.ViBERTgrid_net`` comes first on PYTHONPATH
(``dropin/`` ahead of the reference = the B200 drop-in; the reference alone = the CPU baseline that makes the golden file).

    python tests/harness/run_reference_script.py eval_SROIE  --config cfg.yaml
    torchrun --nproc-per-node 2 tests/harness/run_reference_script.py train_SROIE -c cfg.yaml

Harness-side patches, none of which touches the script files or the hot path:
  1. ``re.compile``: eval_SROIE.py:26 puts the inline flag ``(?i)`` in the middle of its date pattern, an error since
     Python 3.11 -- the flag is hoisted into the ``flags`` argument.
  2. ``eval_SROIE.SROIE_result_filter`` returns None for a date / total string that fails its regex and the caller then takes
     ``len(None)`` (eval_SROIE.py:60-76,199): with untrained weights that is every document.  None is mapped to "".
  3. ``seqeval`` / ``matplotlib`` stand-ins on PYTHONPATH (tests/harness/stubs), imported by the reference at module level.
Instrumentation: the module's training losses, optimizer step counts and a parameter checksum (all-gathered over ranks) are
printed as one ``VBG_HARNESS {json}`` line per rank; with ``VBG_HARNESS_DUMP=<file.npz>`` every eval-mode forward's loss and
``pred_label`` are saved there (the numbers behind the strings the script itself writes).
"""
import argparse
import gc
import importlib
import json
import os
import re
import sys

_orig_compile = re.compile


def _compile(pattern, flags=0):
    if isinstance(pattern, str) and "(?i)" in pattern and not pattern.startswith("(?i)"):
        pattern, flags = pattern.replace("(?i)", ""), flags | re.IGNORECASE
    return _orig_compile(pattern, flags)


def main():
    script, rest = sys.argv[1], sys.argv[2:]
    re.compile = _compile
    import torch
    rec = {"losses": [], "sgd_steps": 0, "adamw_steps": 0, "preds": []}

    import model.ViBERTgrid_net as M
    net_cls = M.ViBERTgridNet
    orig_forward = net_cls.forward

    def forward(self, *a, **k):
        out = orig_forward(self, *a, **k)
        if self.training and isinstance(out, torch.Tensor):
            rec["losses"].append(out.detach().float().reshape(-1)[:1].clone())
        elif not self.training and isinstance(out, tuple) and os.environ.get("VBG_HARNESS_DUMP"):
            rec["preds"].append((out[0].detach().float().reshape(-1)[:1].cpu(), out[4].detach().float().cpu()))
        return out
    net_cls.forward = forward
    for opt, key in ((torch.optim.SGD, "sgd_steps"), (torch.optim.AdamW, "adamw_steps")):
        orig_step = opt.step

        def step(self, *a, _o=orig_step, _k=key, **k):
            rec[_k] += 1
            return _o(self, *a, **k)
        opt.step = step

    mod = importlib.import_module(script)
    if script.startswith("eval_"):
        flt = mod.SROIE_result_filter if hasattr(mod, "SROIE_result_filter") else None
        if flt is not None:
            mod.SROIE_result_filter = lambda s, c: (lambda r: "" if r is None else r)(flt(s, c))
        ap = argparse.ArgumentParser()
        ap.add_argument("--config", required=True)
        mod.main(ap.parse_args(rest))
    else:
        ap = argparse.ArgumentParser()
        ap.add_argument("-c", "--config_path", required=True)
        ap.add_argument("--dist-url", default="env://")
        mod.train(ap.parse_args(rest))

    # ---- what ran?
    nets = [o for o in gc.get_objects() if isinstance(o, net_cls)]
    info = {"script": script, "net_module": net_cls.__module__, "net_file": sys.modules[net_cls.__module__].__file__,
            "model_module_file": M.__file__, "losses": [float(l) for l in rec["losses"]],
            "sgd_steps": rec["sgd_steps"], "adamw_steps": rec["adamw_steps"], "rank": int(os.environ.get("RANK", 0))}
    try:
        from vibertgrid_pytorch_b200 import _lib
        info["launches"], info["so"] = int(_lib.launch_count), _lib.LIB_PATH if _lib._lib is not None else None
    except Exception:
        info["launches"], info["so"] = 0, None
    if nets:
        net = nets[0]
        info["bn_classes"] = sorted({type(m).__name__ for m in net.modules() if isinstance(m, torch.nn.modules.batchnorm._BatchNorm)})
        eng = getattr(net, "_train_engine", None)
        info["train_graph_replays"] = int(getattr(eng, "graph_replays", 0) + getattr(eng, "capture_failures", 0)) if eng is not None else 0
        with torch.no_grad():
            chk = float(sum(p.detach().double().abs().sum() for p in net.parameters()))
            grads = sum(1 for p in net.parameters() if p.grad is not None)
        info["param_checksum"], info["params_with_grad"] = chk, grads
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            allc = [None] * torch.distributed.get_world_size()
            torch.distributed.all_gather_object(allc, (chk, info["losses"]))
            info["checksums_all_ranks"] = [c for c, _ in allc]
            info["losses_all_ranks"] = [l for _, l in allc]
    if os.environ.get("VBG_HARNESS_DUMP") and rec["preds"]:
        import numpy as np
        np.savez_compressed(os.environ["VBG_HARNESS_DUMP"], n=len(rec["preds"]),
                            **{f"loss_{i}": l.numpy() for i, (l, _) in enumerate(rec["preds"])},
                            **{f"pred_{i}": p.numpy() for i, (_, p) in enumerate(rec["preds"])})
    sys.__stdout__.write("VBG_HARNESS " + json.dumps(info) + "\n")
    sys.__stdout__.flush()
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
