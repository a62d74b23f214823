#!/bin/bash
# r2 job 7: full GPU suite + bench (forward + train step) with the streaming ROI kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | grep -E "passed|failed|FAILED|Error|error|assert |mismatch|Timeout" | tail -15 > gpurun_out/r2_pytest_gpu_full.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/r2_pytest_gpu_full.log; cat gpurun_out/r2_pytest_gpu_full.log
timeout 400 python bench.py > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err; echo "bench exit $?"
python - <<'PY'
import json
try:
    j=json.loads(open('gpurun_out/r2_bench_a.json').read().strip().splitlines()[-1])
    print({k:j[k] for k in ('value','ms_per_step')}, j['e2e']['value'], j['roofline']['frac'], {k:(round(v['frac'],3),round(v['ms'],4), v.get('back_to_back',{}).get('frac')) for k,v in j['roofline_hbm_kernels'].items()}, j.get('train_step',{}).get('ms_per_step'), j.get('cpu_baseline'))
except Exception as e:
    print("no bench line:", e)
PY
tail -3 gpurun_out/r2_bench_a.err
