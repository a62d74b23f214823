#!/bin/bash
# GPU job 17: ROI window loads with 4 in flight; pair rule; full tests; bench (+reference arm); traffic capture; launch list.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -s --timeout 100 2>&1 | grep -E "^\[cfg|passed|failed|Error|assert |mismatch|Timeout" | tail -12 > gpurun_out/pytest_gpu.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"grid_scatter|roi_align|gemm_ps" -c 6 -o gpurun_out/prof_traffic \
   python scripts/ncu_traffic.py > gpurun_out/ncu_traffic.log 2>&1
python scripts/extract_traffic.py gpurun_out/prof_traffic.ncu-rep profiles/r1_traffic.json > gpurun_out/traffic.log 2>&1; cp profiles/r1_traffic.json gpurun_out/; cat gpurun_out/traffic.log | head -40
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r1_l.json 2> gpurun_out/bench_r1_l.err; echo "bench exit $?" >> gpurun_out/bench_r1_l.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1_l_ref.json 2> gpurun_out/bench_r1_l_ref.err
python - <<'PY'
import json
j=json.loads(open('gpurun_out/bench_r1_l.json').read().strip().splitlines()[-1])
print({k:j[k] for k in ('value','ms_per_step','gpu_launches')}, j['e2e'], j['roofline']['frac'], j['roofline']['ms'], j['roofline']['traffic'], {k:(round(v['frac'],3),round(v['ms'],4),v['traffic']) for k,v in j['roofline_hbm_kernels'].items()}, j.get('cpu_baseline'), j['clocks'])
PY
head -c 600 gpurun_out/bench_r1_l_ref.json; tail -2 gpurun_out/bench_r1_l.err
VBG_CUDA_GRAPHS=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_r1_l.csv \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/ncu_bench17.log 2>&1
