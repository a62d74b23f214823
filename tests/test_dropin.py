"""The drop-in boundary against the LIVE reference (this container only: skipped where /root/reference is absent, e.g. on
the GPU box).  (1) Putting ``dropin/`` ahead of the reference on PYTHONPATH shadows exactly ``model.ViBERTgrid_net``; the
rest of the reference's ``model`` / ``pipeline`` namespace still resolves to the reference.  (2) For every classifier mode
our parameter tree has the reference's state-dict keys and shapes: a reference module loads ours with strict=True and
vice versa (reference ``train_SROIE.py:280`` / ``eval_SROIE.py:336-337`` checkpoint round trips)."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "model")), reason="live reference not mounted")


def test_dropin_shadows_only_the_model_module():
    env = dict(os.environ, PYTHONSAFEPATH="1", PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "dropin"), ROOT, REF]))
    code = ("from model.ViBERTgrid_net import ViBERTgridNet as N; import model.crf as c, pipeline.transform as t; "
            "print(N.__module__); print(c.__file__); print(t.__file__)")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.strip().splitlines() if not l.startswith("[vibertgrid_b200]")]
    assert lines[0] == "vibertgrid_pytorch_b200.net"
    assert lines[1].startswith(REF) and lines[2].startswith(REF)


@pytest.mark.parametrize("mode", ["simp", "full", "crf"])
def test_state_dict_round_trips_with_the_live_reference(mode, tmp_path, monkeypatch):
    import dataclasses
    import torch
    from vibertgrid_pytorch_b200 import synth
    from vibertgrid_pytorch_b200.net import ViBERTgridNet as Ours
    monkeypatch.chdir(tmp_path)
    cfg = dataclasses.replace(synth.CONFIGS["tiny"], classifier_mode=mode)
    if mode == "crf":
        cfg.tag_to_idx = {f"T{i}": i for i in range(cfg.num_classes)}
    synth.write_bert_dir(cfg, str(tmp_path))
    ours = Ours(**synth.model_kwargs(cfg, "eval"))
    synth.fill_state_dict_(ours, 7)
    monkeypatch.syspath_prepend(REF)
    for m in [k for k in sys.modules if k == "model" or k.startswith("model.")]:
        monkeypatch.delitem(sys.modules, m)
    from model.ViBERTgrid_net import ViBERTgridNet as Ref          # the reference's own module (no dropin on the path)
    assert Ref.__module__ == "model.ViBERTgrid_net" and Ref is not Ours
    kw = synth.model_kwargs(cfg, "eval")
    if mode == "crf":
        kw["tag_to_idx"] = {f"T{i}": i for i in range(cfg.num_classes)}     # the CRF head mutates it in place
    ref = Ref(**kw)
    sd_ours, sd_ref = ours.state_dict(), ref.state_dict()
    assert set(sd_ours) == set(sd_ref)
    assert all(tuple(sd_ours[k].shape) == tuple(sd_ref[k].shape) for k in sd_ref)
    ref.load_state_dict(sd_ours, strict=True)
    ours.load_state_dict(ref.state_dict(), strict=True)
    assert [n for n, _ in ours.named_parameters()] == [n for n, _ in ref.named_parameters()]


BAD_KWARGS = [
    ("work_mode", {"work_mode": "deploy"}),
    ("image_mean length", {"image_mean": [0.5, 0.5]}),
    ("image_std type", {"image_std": "0.1"}),
    ("image_min_size type", {"image_min_size": 512.0}),
    ("image_max_size type", {"image_max_size": 800.0}),
    ("bert name", {"bert_model": "bert-large-uncased"}),
    ("backbone", {"backbone": "resnet_50_fpn"}),
    ("grid_mode", {"grid_mode": "max"}),
    ("classifier_mode", {"classifier_mode": "softmax"}),
    ("crf without tags", {"classifier_mode": "crf", "tag_to_idx": None}),
    ("crf tag format", {"classifier_mode": "crf", "tag_to_idx": {"a": 0, "b": 5}}),
    ("loss_weights type", {"loss_weights": 3}),
    ("tokenizer class", {"tokenizer": object()}),
]


@pytest.mark.parametrize("what,bad", BAD_KWARGS, ids=[w for w, _ in BAD_KWARGS])
def test_constructor_rejects_what_the_reference_rejects(what, bad, tmp_path, monkeypatch):
    """Error behaviour of the constructor: the same invalid arguments raise the same exception class in the
    reference and in the drop-in."""
    from vibertgrid_pytorch_b200 import synth
    from vibertgrid_pytorch_b200.net import ViBERTgridNet as Ours
    monkeypatch.chdir(tmp_path)
    cfg = synth.CONFIGS["tiny"]
    synth.write_bert_dir(cfg, str(tmp_path))
    monkeypatch.syspath_prepend(REF)
    for m in [k for k in sys.modules if k == "model" or k.startswith("model.")]:
        monkeypatch.delitem(sys.modules, m)
    from model.ViBERTgrid_net import ViBERTgridNet as Ref
    errs = []
    for cls in (Ref, Ours):
        kw = {**synth.model_kwargs(cfg, "eval"), **{k: (dict(v) if isinstance(v, dict) else v) for k, v in bad.items()}}
        try:
            cls(**kw)
            errs.append(None)
        except Exception as e:          # noqa: BLE001 - the exception class is what is compared
            errs.append(type(e))
    assert errs[0] is not None, f"the reference accepts {what}: not a rejection case"
    assert errs[1] is errs[0], f"{what}: reference raises {errs[0].__name__}, drop-in {getattr(errs[1], '__name__', None)}"
