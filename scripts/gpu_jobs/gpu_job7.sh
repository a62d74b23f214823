#!/bin/bash
# GPU job 7: separable ROI-align, full-size TC-vs-fp32 tests, N-tile sweep.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_ops.log
echo "pytest ops exit ${PIPESTATUS[0]}" >> gpurun_out/pytest_ops.log
timeout 1500 python -m pytest tests/test_gpu_forward.py -m gpu -q -s 2>&1 | grep -E "^\[cfg|passed|failed|Error|assert |mismatch" | tail -40 > gpurun_out/pytest_fwd.log
for bn in auto 64 128 192 256; do
  if [ "$bn" = auto ]; then timeout 300 python scripts/bn_sweep.py; else VBG_TC3_BN=$bn timeout 300 python scripts/bn_sweep.py; fi
done > gpurun_out/bn_sweep.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r1_e.json 2> gpurun_out/bench_r1_e.err; echo "bench exit $?" >> gpurun_out/bench_r1_e.err
tail -8 gpurun_out/pytest_ops.log; tail -8 gpurun_out/pytest_fwd.log; cat gpurun_out/bn_sweep.log; python - <<'PY'
import json
j=json.load(open('gpurun_out/bench_r1_e.json'))
print({k:j[k] for k in ('value','ms_per_step','e2e','gpu_launches')}); print(j['roofline_hbm_kernels'])
PY
