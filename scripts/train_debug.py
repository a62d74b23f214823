"""GPU debugging aid: one training step against a train_* fixture, printing every gradient's deviation."""
import os, sys, tempfile
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import build_case, load_golden
from test_gpu_train_step import summarize, _to_dev

print("VBG_PRECISION =", os.environ.get("VBG_PRECISION"))
for name in sys.argv[1:] or ["train_tiny"]:
    fx = load_golden(name)
    os.chdir(tempfile.mkdtemp())
    cfg, kw, net, batch = build_case(fx["meta"])
    net = net.cuda(); net.train(); net.bert_hidden_dropout = net.bert_attn_dropout = 0.0
    try:
        loss = net(*_to_dev(batch))
        print(name, "loss", float(loss), "want", float(fx["loss"][0]), {k: float(v) for k, v in net._train_engine.last.items()})
        loss.backward(); torch.cuda.synchronize()
    except Exception:
        import traceback; traceback.print_exc(); continue
    params = dict(net.named_parameters())
    for k in fx["grad_names"]:
        k = str(k); ref = fx["g:" + k]
        if params[k].grad is None:
            print("NOGRAD", k); continue
        got = summarize(params[k].grad)
        scale = max(np.abs(ref[2:]).max(), ref[1] / np.sqrt(params[k].numel()), 1e-12)
        err = np.abs(got[2:] - ref[2:]).max() / scale
        nerr = abs(got[1] - ref[1]) / max(ref[1], 1e-12)
        flag = "BAD " if (err > 5e-3 or nerr > 5e-3) else "ok  "
        print(f"{flag}{k:90s} err={err:.2e} norm_err={nerr:.2e} norm={ref[1]:.3e}")
