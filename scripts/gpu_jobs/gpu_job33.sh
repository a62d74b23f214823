#!/bin/bash
mkdir -p gpurun_out
timeout 120 python scripts/timeline.py > gpurun_out/timeline3.log 2>&1; grep -A12 "3072x768" gpurun_out/timeline3.log | grep -E "^\[|epilogue|epi_"
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout 60 -k "presplit or tensor_core or stem" 2>&1 | tail -2
timeout 400 python -m pytest tests/test_gpu_forward.py -m gpu -q --timeout 100 2>&1 | tail -2
timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1_w.json 2> gpurun_out/bench_r1_w.err; echo "bench exit $?" >> gpurun_out/bench_r1_w.err
python - <<'PY'
import json
j=json.loads(open('gpurun_out/bench_r1_w.json').read().strip().splitlines()[-1])
print({k:j[k] for k in ('value','ms_per_step','gpu_launches')}, j['e2e'], j['roofline']['frac'], j['roofline']['ms'], j['roofline'].get('warm_l2'))
PY
