#!/bin/bash
# GPU job 34: final verification of HEAD: all gpu tests, smoke, bench (with cpu baseline), reference arm.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s --timeout 100 2>&1 | grep -E "^\[cfg|passed|failed|Error|assert |mismatch|Timeout" | tail -8 > gpurun_out/pytest_gpu.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -3 gpurun_out/smoke.log
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_final_ref.json 2> gpurun_out/bench_final_ref.err
timeout 500 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench exit $?" >> gpurun_out/bench_final.err
python - <<'PY'
import json
j=json.loads(open('gpurun_out/bench_final.json').read().strip().splitlines()[-1])
print({k:j[k] for k in ('value','ms_per_step','gpu_launches','steps','warmup')}, j['e2e'], j['roofline']['frac'], j['roofline']['ms'], j['roofline'].get('warm_l2'), {k:(round(v['frac'],3),round(v['ms'],4)) for k,v in j['roofline_hbm_kernels'].items()}, j.get('cpu_baseline'), j['clocks'])
r=json.loads(open('gpurun_out/bench_final_ref.json').read().strip().splitlines()[-1]); print('reference arm', r['value'], r['cpu_baseline'])
PY
wc -l gpurun_out/bench_final.json; tail -2 gpurun_out/bench_final.err
