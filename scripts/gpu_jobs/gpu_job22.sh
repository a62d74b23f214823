#!/bin/bash
# GPU job 22: attention fast paths (unmasked chunks, ex2.approx, 16-byte plane stores): tests, timeline, bench.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout 60 -k "attention" 2>&1 | tail -3
timeout 120 python scripts/timeline.py 2>&1 | tail -13 > gpurun_out/timeline_attn2.log; cat gpurun_out/timeline_attn2.log
timeout 600 python -m pytest tests/test_gpu_forward.py -m gpu -q -s --timeout 100 2>&1 | grep -E "^\[cfg|passed|failed|Error|assert |mismatch|Timeout" | tail -8
timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1_p.json 2> gpurun_out/bench_r1_p.err; echo "bench exit $?" >> gpurun_out/bench_r1_p.err
python - <<'PY'
import json
j=json.loads(open('gpurun_out/bench_r1_p.json').read().strip().splitlines()[-1])
print({k:j[k] for k in ('value','ms_per_step','gpu_launches')}, j['e2e'], j['roofline']['frac'], j['roofline']['ms'])
PY
tail -2 gpurun_out/bench_r1_p.err
