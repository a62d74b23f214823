#!/bin/bash
mkdir -p gpurun_out
timeout 120 python scripts/timeline.py 2>&1 | tail -12 > gpurun_out/timeline_attn4.log; cat gpurun_out/timeline_attn4.log
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout 60 -k "attention" 2>&1 | tail -2
timeout 400 python -m pytest tests/test_gpu_forward.py -m gpu -q --timeout 100 2>&1 | tail -2
timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/bench_r1_y.json 2> gpurun_out/bench_r1_y.err; echo "bench exit $?" >> gpurun_out/bench_r1_y.err
python - <<'PY'
import json
j=json.loads(open('gpurun_out/bench_r1_y.json').read().strip().splitlines()[-1])
print({k:j[k] for k in ('value','ms_per_step','gpu_launches')}, j['e2e'])
PY
