#!/bin/bash
# GPU job 49: verification of HEAD after the training path: all gpu tests, smoke (forward + training step), bench with train_step.
mkdir -p gpurun_out
timeout 1000 python -m pytest tests -m gpu -q --timeout 150 -o faulthandler_timeout=140 2>&1 | grep -E "passed|failed|FAILED|Error|assert |mismatch|Timeout|gradients off" | tail -14 > gpurun_out/pytest_gpu.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log; tail -8 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -4 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench exit $?" >> gpurun_out/bench_final.err
python - <<'PY'
import json
j=json.loads(open('gpurun_out/bench_final.json').read().strip().splitlines()[-1])
print({k:j[k] for k in ('value','ms_per_step','gpu_launches','steps','warmup')}, j['e2e'], j['roofline']['frac'], j['roofline']['ms'], {k:(round(v['frac'],3),round(v['ms'],4)) for k,v in j['roofline_hbm_kernels'].items()}, j.get('cpu_baseline'), j['clocks'])
print('train_step', j.get('train_step'))
PY
wc -l gpurun_out/bench_final.json; tail -2 gpurun_out/bench_final.err
