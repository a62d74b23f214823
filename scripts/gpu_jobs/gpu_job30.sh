#!/bin/bash
# GPU job 30: split-K for under-filled GEMMs: targeted tests first, then full tests, A/B bench.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -x --timeout 60 -k "split_k" 2>&1 | grep -E "^E   |passed|failed" | head -12 > gpurun_out/pytest_sk.log; echo "exit ${PIPESTATUS[0]}" >> gpurun_out/pytest_sk.log; cat gpurun_out/pytest_sk.log
if grep -q "^exit 0" gpurun_out/pytest_sk.log; then
timeout 600 python -m pytest tests -m gpu -q -s --timeout 100 2>&1 | grep -E "^\[cfg|passed|failed|Error|assert |mismatch|Timeout" | tail -8
VBG_PS_SPLITK=0 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/bench_r1_u0.json 2> gpurun_out/bench_r1_u0.err
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/bench_r1_u1.json 2> gpurun_out/bench_r1_u1.err
for f in gpurun_out/bench_r1_u0.json gpurun_out/bench_r1_u1.json; do python - $f <<'PY'
import json,sys
try:
    j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], {k:j[k] for k in ('value','ms_per_step','gpu_launches')}, j['e2e'])
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
tail -2 gpurun_out/bench_r1_u1.err
fi
