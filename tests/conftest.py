import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, f"{name}.npz"))
    fx = {k: z[k] for k in z.files}
    fx["meta"] = json.loads(str(fx["meta"]))
    return fx


@pytest.fixture
def bert_dir(tmp_path, monkeypatch):
    """chdir into a temp dir holding the stand-in ``bert-base-uncased/`` directory."""
    from vibertgrid_pytorch_b200 import synth

    def make(cfg):
        synth.write_bert_dir(cfg, str(tmp_path))
        monkeypatch.chdir(tmp_path)
        return str(tmp_path)
    return make


def build_case(meta):
    """(cfg, kwargs, model with seeded weights, batch) for a golden fixture's meta."""
    import dataclasses
    from vibertgrid_pytorch_b200 import synth
    from vibertgrid_pytorch_b200.net import ViBERTgridNet
    cfg = dataclasses.replace(synth.CONFIGS[meta["cfg"]], classifier_mode=meta["classifier_mode"])
    if meta["classifier_mode"] == "crf" and cfg.tag_to_idx is None:
        cfg.tag_to_idx = {f"T{i}": i for i in range(cfg.num_classes)}
    synth.write_bert_dir(cfg, os.getcwd())
    kw = {**synth.model_kwargs(cfg, "eval"), **meta.get("extra_kwargs", {})}
    net = ViBERTgridNet(**kw)
    synth.fill_state_dict_(net, meta["weight_seed"])
    batch = synth.make_batch(cfg, meta["input_seed"])
    return cfg, kw, net, batch


def relerr(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
