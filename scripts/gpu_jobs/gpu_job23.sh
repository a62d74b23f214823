#!/bin/bash
# GPU job 23: separable windowed ROI-align; attention with both TMEM halves in flight.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout 60 -k "roi or attention or format_aware" 2>&1 | tail -3
timeout 600 python -m pytest tests/test_gpu_forward.py -m gpu -q -s --timeout 100 2>&1 | grep -E "^\[cfg|passed|failed|Error|assert |mismatch|Timeout" | tail -8
timeout 120 python scripts/timeline.py 2>&1 | tail -12 > gpurun_out/timeline_attn3.log; cat gpurun_out/timeline_attn3.log
timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1_q.json 2> gpurun_out/bench_r1_q.err; echo "bench exit $?" >> gpurun_out/bench_r1_q.err
python - <<'PY'
import json
j=json.loads(open('gpurun_out/bench_r1_q.json').read().strip().splitlines()[-1])
print({k:j[k] for k in ('value','ms_per_step','gpu_launches')}, j['e2e'], j['roofline']['frac'], j['roofline']['ms'], {k:(round(v['frac'],3),round(v['ms'],4)) for k,v in j['roofline_hbm_kernels'].items()})
PY
tail -2 gpurun_out/bench_r1_q.err
