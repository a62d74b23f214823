#!/bin/bash
# GPU job 13: 8-warp swizzled epilogue, scale/shift prefetch, windowed ROI-align, plane-copy scatter.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "^\[cfg|^\[|passed|failed|Error|assert |mismatch" | tail -40 > gpurun_out/pytest_gpu.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
timeout 600 python scripts/ps_sweep.py > gpurun_out/ps_sweep3.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1_i.json 2> gpurun_out/bench_r1_i.err; echo "bench exit $?" >> gpurun_out/bench_r1_i.err
VBG_ROI_DIRECT=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1_i_roidirect.json 2> /dev/null
tail -8 gpurun_out/pytest_gpu.log; cat gpurun_out/ps_sweep3.log
for f in gpurun_out/bench_r1_i.json gpurun_out/bench_r1_i_roidirect.json; do python - $f <<'PY'
import json,sys
j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print({k:j[k] for k in ('value','ms_per_step','gpu_launches')}, j['e2e']['value'], j['roofline']['frac'], j['roofline']['ms'], {k:(v['frac'],v['ms']) for k,v in j['roofline_hbm_kernels'].items()})
PY
done
tail -2 gpurun_out/bench_r1_i.err
