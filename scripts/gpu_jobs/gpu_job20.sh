#!/bin/bash
# GPU job 20: two-stream fork/join forward: tests, bench with 1 vs 2 streams.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -s --timeout 100 2>&1 | grep -E "^\[cfg|passed|failed|Error|assert |mismatch|Timeout" | tail -12 > gpurun_out/pytest_gpu.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log; tail -6 gpurun_out/pytest_gpu.log
VBG_STREAMS=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/bench_r1_n_1stream.json 2> gpurun_out/bench_r1_n_1stream.err
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/bench_r1_n.json 2> gpurun_out/bench_r1_n.err; echo "bench exit $?" >> gpurun_out/bench_r1_n.err
for f in gpurun_out/bench_r1_n_1stream.json gpurun_out/bench_r1_n.json; do python - $f <<'PY'
import json,sys
try:
    j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], {k:j[k] for k in ('value','ms_per_step','gpu_launches')}, j['e2e'])
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
tail -3 gpurun_out/bench_r1_n.err
