#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/host_cost.py 2>/dev/null | tail -2
timeout 300 python -m pytest tests/test_gpu_forward.py -m gpu -q --timeout 100 -k "two_stream" 2>&1 | tail -2
for i in 1 2; do
timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/bench_r1_t$i.json 2> gpurun_out/bench_r1_t$i.err
python - gpurun_out/bench_r1_t$i.json <<'PY'
import json,sys
j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print({k:j[k] for k in ('value','ms_per_step','gpu_launches')}, j['e2e'])
PY
done
