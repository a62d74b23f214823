#!/bin/bash
mkdir -p gpurun_out
timeout 120 python scripts/timeline_roi.py > gpurun_out/timeline_roi3.log 2>&1; cat gpurun_out/timeline_roi3.log
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout 60 -k "roi or format_aware" 2>&1 | tail -2
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1_z.json 2> gpurun_out/bench_r1_z.err
python - <<'PY'
import json
j=json.loads(open('gpurun_out/bench_r1_z.json').read().strip().splitlines()[-1])
print({k:j[k] for k in ('value','ms_per_step')}, {k:(round(v['frac'],3),round(v['ms'],4)) for k,v in j['roofline_hbm_kernels'].items()})
PY
