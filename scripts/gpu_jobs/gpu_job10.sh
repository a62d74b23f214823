#!/bin/bash
# GPU job 10: pre-split (bf16 hi/lo planes end to end) GEMM/conv path: parity tests, shape sweep per ring variant, forward tests, bench.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "presplit or split_merge or format_aware or attention" 2>&1 | tail -30 > gpurun_out/pytest_ps.log
echo "pytest ps exit ${PIPESTATUS[0]}" >> gpurun_out/pytest_ps.log
timeout 600 python scripts/ps_sweep.py > gpurun_out/ps_sweep.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "^\[cfg|^\[|passed|failed|Error|assert |mismatch" | tail -60 > gpurun_out/pytest_gpu.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
for kb in 64 32; do
  VBG_PS_KB=$kb timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1_ps$kb.json 2> gpurun_out/bench_r1_ps$kb.err; echo "bench exit $?" >> gpurun_out/bench_r1_ps$kb.err
done
VBG_PRESPLIT=0 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/bench_r1_nops.json 2> gpurun_out/bench_r1_nops.err
tail -12 gpurun_out/pytest_ps.log; cat gpurun_out/ps_sweep.log; tail -12 gpurun_out/pytest_gpu.log
for f in gpurun_out/bench_r1_ps64.json gpurun_out/bench_r1_ps32.json gpurun_out/bench_r1_nops.json; do python - "$f" <<'PY'
import json,sys
try:
    j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], {k:j[k] for k in ('value','ms_per_step','gpu_launches')}, j['e2e']['value'], j.get('roofline',{}).get('frac'), {k:v['frac'] for k,v in j.get('roofline_hbm_kernels',{}).items()})
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
tail -3 gpurun_out/bench_r1_ps64.err
