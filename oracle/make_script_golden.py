"""ORACLE tooling (test infrastructure): run the UNMODIFIED reference's ``eval_SROIE.main`` on CPU over the synthetic SROIE
tree of tests/harness/sroie_synth.py (BASELINE configs[0]: "CPU forward via reference eval_SROIE.py") and commit what it
writes -- ``result/<weights>.json``, the per-document predicted key strings -- as ``tests/golden/eval_sroie_cfg1.json``.
tests/test_gpu_reference_scripts.py then runs the same script file over the same tree with ``dropin/`` ahead of the
reference on PYTHONPATH (the B200 path) and compares.  Build container only.

    python oracle/stage_reference.py && python oracle/make_script_golden.py
"""
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests", "harness")]
import sroie_synth  # noqa: E402

if __name__ == "__main__":
    with tempfile.TemporaryDirectory() as tmp:
        cpath, rpath = sroie_synth.prepare_eval_case(tmp, "cpu")
        env = sroie_synth.script_env(ROOT, with_dropin=False)
        env["VBG_HARNESS_DUMP"] = os.path.join(ROOT, "tests", "golden", "eval_sroie_cfg1_preds.npz")
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "harness", "run_reference_script.py"), "eval_SROIE",
                              "--config", cpath], cwd=tmp, env=env, capture_output=True, text=True)
        print(out.stdout[-3000:], out.stderr[-3000:])
        assert out.returncode == 0
        info = json.loads([l for l in out.stdout.splitlines() if l.startswith("VBG_HARNESS ")][-1][len("VBG_HARNESS "):])
        assert info["net_module"] == "model.ViBERTgrid_net" and "oracle/_ref/reference" in info["net_file"], info
        res = json.load(open(rpath))
    dst = os.path.join(ROOT, "tests", "golden", "eval_sroie_cfg1.json")
    json.dump(res, open(dst, "w"), indent=1, sort_keys=True)
    print(f"wrote {dst}: {len(res['per_sample'])} documents, method {res['method']}")
