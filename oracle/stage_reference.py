"""ORACLE tooling (test / bench infrastructure only) -- stage the UNMODIFIED reference under ``oracle/_ref/reference``.

The reference is pure Python (no build step): "compiling" it is copying its source files, byte for byte, from where they lie
under /root/reference into the git-ignored ``oracle/_ref/`` so that they travel to the GPU box with the snapshot (the box has
no /root/reference).  Nothing is copied into tracked paths; ``oracle/_ref/MANIFEST.json`` records the sha256 of every staged
file so a test can prove the staged tree is the reference's.  Used by
  * tests/test_gpu_reference_scripts.py -- the reference's own eval_SROIE.py / train_SROIE.py driven through the drop-in,
  * bench.py --impl reference            -- the reference's own CPU forward as the baseline arm (kind: "reference").

    python oracle/stage_reference.py            # no-op when /root/reference is absent (GPU box: uses the staged files)
"""
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("VBG_REFERENCE_SRC", "/root/reference")
DST = os.path.join(ROOT, "oracle", "_ref", "reference")
KEEP = (".py", ".yaml", ".txt", ".md")


def staged() -> bool:
    return os.path.isfile(os.path.join(DST, "model", "ViBERTgrid_net.py"))


def stage(verbose=True) -> bool:
    if not os.path.isdir(os.path.join(SRC, "model")):
        if verbose:
            print(f"[stage_reference] {SRC} absent; staged copy {'present' if staged() else 'ABSENT'}")
        return staged()
    manifest = {}
    for d, _, files in os.walk(SRC):
        if ".git" in d.split(os.sep):
            continue
        for f in files:
            if not f.endswith(KEEP):
                continue
            src = os.path.join(d, f)
            rel = os.path.relpath(src, SRC)
            dst = os.path.join(DST, rel)
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            shutil.copyfile(src, dst)
            manifest[rel] = hashlib.sha256(open(src, "rb").read()).hexdigest()
    with open(os.path.join(ROOT, "oracle", "_ref", "MANIFEST.json"), "w") as fh:
        json.dump({"source": SRC, "files": manifest}, fh, indent=1, sort_keys=True)
    if verbose:
        print(f"[stage_reference] staged {len(manifest)} files -> {DST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)
