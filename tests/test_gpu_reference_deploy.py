"""the reference's DEPLOYMENT path, unmodified, through the drop-in on the GPU -- ``deployment.module_load.
inference_init`` (yaml -> ViBERTgridNet(work_mode="inference") -> from_pretrained BERT -> checkpoint) and
``deployment.inference_SROIE.inference_pipe`` (JPEG bytes -> generate_batch -> model.inference -> SROIE_postprocessing) -- against
the same path with the reference's own module (eager, same device): identical key dictionary, ``pred_label`` within 1e-3, the
drop-in replaying its CUDA graph from the third request on.  The only stand-in is the external OCR service (an HTTP call in the
reference): tests/harness/run_reference_deploy.py returns a recorded OCR result instead."""
import dataclasses
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch
import yaml

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "tests", "harness"))

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not os.path.isfile(os.path.join(ROOT, "oracle", "_ref", "reference", "deployment", "module_load.py")),
                                 reason="staged reference absent (python oracle/stage_reference.py in the build container)")]
LAUNCHER = os.path.join(ROOT, "tests", "harness", "run_reference_deploy.py")


def _case(tmp):
    """One synthetic receipt (JPEG + the OCR service's answer), a stand-in BERT directory WITH weights (work_mode="inference"
    calls from_pretrained), a seeded checkpoint and deployment/config/network_config.yaml's keys."""
    import sroie_synth
    from PIL import Image
    from vibertgrid_pytorch_b200 import synth
    cfg = dataclasses.replace(synth.CONFIGS["cfg1"], bert_layers=2)
    g = torch.Generator().manual_seed(77)
    img = (torch.rand(cfg.height, cfg.width, 3, generator=g) * 255).to(torch.uint8).numpy()
    ipath = os.path.join(tmp, "receipt.jpg")
    Image.fromarray(img, "RGB").save(ipath, quality=92)
    boxes = synth.make_boxes(cfg.segments, cfg.height, cfg.width, g)
    ids = torch.randint(1000, cfg.vocab_size, (cfg.segments, 4), generator=g)
    texts = [" ".join(f"tok{int(i)}" for i in ids[s]) for s in range(cfg.segments)]
    texts[3], texts[7] = "", "   "                                   # dropped by generate_batch's filter
    json.dump({"text": texts, "coors": [[int(v) for v in b] for b in boxes]}, open(os.path.join(tmp, "ocr.json"), "w"))
    sroie_synth.write_bert(cfg, tmp, with_weights=True, seed=3)
    ck = os.path.join(tmp, "deploy_seed0.pth")
    sroie_synth.write_checkpoint(ck, cfg, tmp, seed=0)
    hyp = dict(ocr_url="http://ocr.invalid/api", parse_mode="eng_line", weights=ck, num_classes=cfg.num_classes,
               image_mean=[0.9248, 0.9224, 0.9215], image_std=[0.1532, 0.1545, 0.1536], bert_version=cfg.bert_name,
               backbone=cfg.backbone, grid_mode="mean", early_fusion_downsampling_ratio=8, roi_shape=7, p_fuse_downsampling_ratio=4,
               late_fusion_fuse_embedding_channel=1024, layer_mode="single", classifier_mode=cfg.classifier_mode)
    cpath = os.path.join(tmp, "network_config.yaml")
    yaml.safe_dump(hyp, open(cpath, "w"))
    return cpath, ipath, os.path.join(tmp, "ocr.json")


def _run(tmp, with_dropin, cpath, ipath, opath, tag):
    import sroie_synth
    env = sroie_synth.script_env(ROOT, with_dropin=with_dropin)
    env["VBG_ALLOW_RANDOM_INIT"] = "0"
    out = os.path.join(tmp, f"result_{tag}.json")
    r = subprocess.run([sys.executable, LAUNCHER, "--config", cpath, "--img", ipath, "--ocr", opath, "--out", out], cwd=tmp, env=env,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    info = json.load(open(out))
    preds = np.load(out + ".npz")
    return info, [preds[f"pred_{i}"] for i in range(len(preds.files))]


def test_inference_pipe_through_the_dropin(tmp_path):
    tmp = str(tmp_path)
    cpath, ipath, opath = _case(tmp)
    ours, p_ours = _run(tmp, True, cpath, ipath, opath, "dropin")
    ref, p_ref = _run(tmp, False, cpath, ipath, opath, "reference")
    # what ran
    assert ours["net_module"] == "vibertgrid_pytorch_b200.net" and ours["so"] and ours["so"].endswith("libvbg_sm100a.so")
    assert ours["launches"] > 100 and ours["device"].startswith("cuda") and ours["work_mode"] in ("inference", "train")
    assert ours["graph_replays"] >= 1, "the third identical request is expected to replay the captured graph"
    assert ref["net_module"] == "model.ViBERTgrid_net" and "oracle/_ref/reference" in ref["net_file"].replace(os.sep, "/")
    # what came out: 3 requests each, same bytes in -> same answer out
    assert len(p_ours) == len(p_ref) == 3
    for a in p_ours[1:]:
        assert np.array_equal(a, p_ours[0]), "eager, captured and replayed requests must agree exactly"
    a, b = p_ours[0], p_ref[0]
    assert a.shape == b.shape and a.shape[0] == 126                          # two of the 128 OCR lines are blank
    err = float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
    print(f"[deployment path] pred_label max-rel {err:.1e}; keys {ours['results'][0]}")
    assert err < 1e-3 and np.array_equal(a.argmax(1), b.argmax(1))
    assert ours["results"] == ref["results"]
