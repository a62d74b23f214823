#!/bin/bash
mkdir -p gpurun_out
timeout 120 python scripts/timeline.py > gpurun_out/timeline4.log 2>&1; grep -A10 "3072x768" gpurun_out/timeline4.log | grep -E "^\[|epi_|exit"
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout 60 -k "presplit or tensor_core" 2>&1 | tail -2
timeout 400 python -m pytest tests/test_gpu_forward.py -m gpu -q -s --timeout 100 2>&1 | grep -E "^\[cfg|passed|failed" | tail -6
timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1_x.json 2> gpurun_out/bench_r1_x.err; echo "bench exit $?" >> gpurun_out/bench_r1_x.err
python - <<'PY'
import json
j=json.loads(open('gpurun_out/bench_r1_x.json').read().strip().splitlines()[-1])
print({k:j[k] for k in ('value','ms_per_step','gpu_launches')}, j['e2e'], j['roofline']['frac'], j['roofline']['ms'], j['roofline'].get('warm_l2'))
PY
