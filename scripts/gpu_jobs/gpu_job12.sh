#!/bin/bash
# GPU job 12: lean tensor-core epilogue (hoisted flags, packed bf16 conversion, incremental row offsets): tests, sweep, bench, ncu.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "^\[cfg|^\[|passed|failed|Error|assert |mismatch" | tail -40 > gpurun_out/pytest_gpu.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
timeout 600 python scripts/ps_sweep.py > gpurun_out/ps_sweep2.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1_h.json 2> gpurun_out/bench_r1_h.err; echo "bench exit $?" >> gpurun_out/bench_r1_h.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_ps_kernel -c 8 -o gpurun_out/prof_ps2 \
   python scripts/ps_probe.py > gpurun_out/ncu_ps2.log 2>&1
VBG_CUDA_GRAPHS=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_r1_h.csv \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/ncu_bench12.log 2>&1
tail -8 gpurun_out/pytest_gpu.log; cat gpurun_out/ps_sweep2.log
python - <<'PY'
import json
j=json.loads(open('gpurun_out/bench_r1_h.json').read().strip().splitlines()[-1])
print({k:j[k] for k in ('value','ms_per_step','gpu_launches')}, j['e2e']['value'], j['roofline']['frac'], j['roofline']['ms'], {k:(v['frac'],v['ms']) for k,v in j['roofline_hbm_kernels'].items()})
PY
tail -2 gpurun_out/bench_r1_h.err
