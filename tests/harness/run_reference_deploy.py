"""Test harness (NOT product code): the reference's DEPLOYMENT path, unmodified, from the staged tree oracle/_ref/reference --
``deployment.module_load.inference_init`` (yaml -> tokenizer -> ViBERTgridNet(work_mode="inference") -> checkpoint) and
``deployment.inference_SROIE.inference_pipe`` (image bytes -> generate_batch -> model.inference -> SROIE_postprocessing) --
with whichever ``model.ViBERTgrid_net`` comes first on PYTHONPATH (``dropin/`` first = the B200 module).

    python tests/harness/run_reference_deploy.py --config net.yaml --img doc.jpg --ocr ocr.json --out result.json

Harness-side patches (script files untouched): the external OCR HTTP call (``ocr_extraction``, inference_preporcessing.py:116)
returns the recorded result in ``ocr.json`` instead of POSTing to ``ocr_url``; ``re.compile`` hoists the mid-pattern ``(?i)``
of the date filter (an error since Python 3.11); ``ltp`` is a stand-in package (tests/harness/stubs)."""
import argparse
import json
import re
import sys

_orig_compile = re.compile


def _compile(pattern, flags=0):
    if isinstance(pattern, str) and "(?i)" in pattern and not pattern.startswith("(?i)"):
        pattern, flags = pattern.replace("(?i)", ""), flags | re.IGNORECASE
    return _orig_compile(pattern, flags)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", required=True)
    ap.add_argument("--img", required=True)
    ap.add_argument("--ocr", required=True)
    ap.add_argument("--out", required=True)
    ap.add_argument("--repeat", type=int, default=3)
    args = ap.parse_args()
    re.compile = _compile
    import numpy as np
    import torch
    import deployment.inference_preporcessing as P
    ocr = json.load(open(args.ocr))
    P.ocr_extraction = lambda image_bytes, ocr_url, parse_mode: (200, list(ocr["text"]), [list(c) for c in ocr["coors"]])
    import deployment.inference_SROIE as D
    from deployment.module_load import inference_init
    import model.ViBERTgrid_net as M
    model, ocr_url, tokenizer, device, num_classes, parse_mode = inference_init(dir_config=args.config)
    preds = []
    orig = model.inference

    def inference(*a, **k):
        out = orig(*a, **k)
        preds.append(out.detach().float().cpu())
        return out
    model.inference = inference
    image_bytes = open(args.img, "rb").read()
    results = [D.inference_pipe(model, ocr_url, tokenizer, device, num_classes, image_bytes=image_bytes, parse_mode=parse_mode)
               for _ in range(args.repeat)]                 # repeated requests: eager, graph capture, graph replay in the drop-in
    info = {"net_module": M.ViBERTgridNet.__module__, "net_file": sys.modules[M.ViBERTgridNet.__module__].__file__,
            "device": str(device), "work_mode": getattr(model, "work_mode", None), "results": results}
    try:
        from vibertgrid_pytorch_b200 import _lib
        info["launches"], info["so"] = int(_lib.launch_count), _lib.LIB_PATH if _lib._lib is not None else None
        eng = getattr(model, "_engine", None)
        info["graph_replays"] = int(getattr(eng, "graph_replays", 0)) if eng is not None else 0
    except Exception:
        info["launches"], info["so"] = 0, None
    np.savez_compressed(args.out + ".npz", **{f"pred_{i}": p.numpy() for i, p in enumerate(preds)})
    json.dump(info, open(args.out, "w"))
    sys.__stdout__.write("VBG_HARNESS " + json.dumps({k: v for k, v in info.items() if k != "results"}) + "\n")


if __name__ == "__main__":
    main()
