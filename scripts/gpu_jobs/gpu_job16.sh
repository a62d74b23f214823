#!/bin/bash
# GPU job 16: CTA-pair (cta_group::2) pre-split GEMM: targeted parity first (tight timeouts), then sweep, full tests, bench.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -x --timeout 60 -k "presplit" 2>&1 | tail -15 > gpurun_out/pytest_cg2.log
echo "pytest cg2 exit ${PIPESTATUS[0]}" >> gpurun_out/pytest_cg2.log
tail -6 gpurun_out/pytest_cg2.log
if grep -q "pytest cg2 exit 0" gpurun_out/pytest_cg2.log; then
  timeout 300 python scripts/ps_sweep.py > gpurun_out/ps_sweep4.log 2>&1; cat gpurun_out/ps_sweep4.log
  timeout 600 python -m pytest tests -m gpu -q -s --timeout 100 2>&1 | grep -E "^\[cfg|passed|failed|Error|assert |mismatch|Timeout" | tail -12 > gpurun_out/pytest_gpu.log
  echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
  for cg in 0 1 auto; do
    if [ $cg = auto ]; then timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1_k_auto.json 2> gpurun_out/bench_r1_k_auto.err
    else VBG_PS_CG2=$cg timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1_k_cg$cg.json 2> gpurun_out/bench_r1_k_cg$cg.err; fi
  done
  for f in gpurun_out/bench_r1_k_cg0.json gpurun_out/bench_r1_k_cg1.json gpurun_out/bench_r1_k_auto.json; do python - $f <<'PY'
import json,sys
try:
    j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], {k:j[k] for k in ('value','ms_per_step')}, j['e2e']['value'], j['roofline']['frac'], j['roofline']['ms'], {k:(round(v['frac'],3),round(v['ms'],4)) for k,v in j['roofline_hbm_kernels'].items()})
except Exception as e: print(sys.argv[1], "ERR", e)
PY
  done
  tail -3 gpurun_out/bench_r1_k_auto.err
fi
