#!/bin/bash
# GPU job 57: verification of HEAD with the row-per-warp ROI-align kernel as the default: all gpu tests, then a short bench.
mkdir -p gpurun_out
timeout 100 python -m pytest tests -m gpu -q -x --timeout 100 -o faulthandler_timeout=90 2>&1 | grep -E "passed|failed|FAILED|Error|assert |mismatch|Timeout|gradients off" | tail -10 > gpurun_out/pytest_gpu.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
timeout 40 python bench.py --no-train --no-cpu-baseline > gpurun_out/bench_roi.json 2> gpurun_out/bench_roi.err; echo "bench exit $?" >> gpurun_out/bench_roi.err
python - <<'PY'
import json
try:
    j=json.loads(open('gpurun_out/bench_roi.json').read().strip().splitlines()[-1])
    print({k:j[k] for k in ('value','ms_per_step')}, j['e2e']['value'], j['roofline']['frac'], {k:(round(v['frac'],3),round(v['ms'],4)) for k,v in j['roofline_hbm_kernels'].items()})
except Exception as e:
    print("no bench line:", e)
PY
