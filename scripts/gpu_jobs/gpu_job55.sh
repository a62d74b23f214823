#!/bin/bash
# GPU job 55: verification of HEAD: all gpu tests, smoke, bench (forward + e2e + rooflines + cpu baseline + training step), then the
# cfg5 (CRF head, 1024 x 1024) training step at full size.
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q --timeout 150 -o faulthandler_timeout=140 2>&1 | grep -E "passed|failed|FAILED|Error|assert |mismatch|Timeout|gradients off" | tail -14 > gpurun_out/pytest_gpu.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log; tail -6 gpurun_out/pytest_gpu.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -3 gpurun_out/smoke.log
timeout 400 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench exit $?" >> gpurun_out/bench_final.err
python - <<'PY'
import json
j=json.loads(open('gpurun_out/bench_final.json').read().strip().splitlines()[-1])
print({k:j[k] for k in ('value','ms_per_step','gpu_launches','steps','warmup')}, j['e2e'], j['roofline']['frac'], j['roofline']['ms'], {k:(round(v['frac'],3),round(v['ms'],4)) for k,v in j['roofline_hbm_kernels'].items()}, j.get('cpu_baseline'), j['clocks'])
print('train_step', j.get('train_step'))
PY
tail -2 gpurun_out/bench_final.err
timeout 150 python scripts/train_bench.py cfg5 4 2>&1 | grep -E "^cfg5|Error|error" > gpurun_out/train_step_cfg5.log; cat gpurun_out/train_step_cfg5.log
