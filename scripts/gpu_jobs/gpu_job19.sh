#!/bin/bash
# GPU job 19: GEMM timeline of CTA 0, prefetching e2e bench, quick tests.
mkdir -p gpurun_out
timeout 120 python scripts/timeline.py > gpurun_out/timeline.log 2>&1; cat gpurun_out/timeline.log
timeout 300 python -m pytest tests/test_gpu_forward.py -m gpu -q --timeout 100 -x 2>&1 | tail -3
timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1_m.json 2> gpurun_out/bench_r1_m.err; echo "bench exit $?" >> gpurun_out/bench_r1_m.err
python - <<'PY'
import json
j=json.loads(open('gpurun_out/bench_r1_m.json').read().strip().splitlines()[-1])
print({k:j[k] for k in ('value','ms_per_step','gpu_launches')}, j['e2e'])
PY
tail -2 gpurun_out/bench_r1_m.err
