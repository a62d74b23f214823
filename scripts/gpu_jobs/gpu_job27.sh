#!/bin/bash
# GPU job 27: consolidated verification of the current tree: all gpu tests, smoke, traffic capture, bench (+cpu baseline), launch list.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s --timeout 100 2>&1 | grep -E "^\[cfg|passed|failed|Error|assert |mismatch|Timeout" | tail -12 > gpurun_out/pytest_gpu.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -3 gpurun_out/smoke.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"grid_scatter|roi_align|gemm_ps" -c 6 -o gpurun_out/prof_traffic2 \
   python scripts/ncu_traffic.py > gpurun_out/ncu_traffic2.log 2>&1
python scripts/extract_traffic.py gpurun_out/prof_traffic2.ncu-rep gpurun_out/r1_traffic.json > gpurun_out/traffic2.log 2>&1; cp gpurun_out/r1_traffic.json profiles/r1_traffic.json
timeout 500 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r1_s.json 2> gpurun_out/bench_r1_s.err; echo "bench exit $?" >> gpurun_out/bench_r1_s.err
python - <<'PY'
import json
j=json.loads(open('gpurun_out/bench_r1_s.json').read().strip().splitlines()[-1])
print({k:j[k] for k in ('value','ms_per_step','gpu_launches')}, j['e2e'], j['roofline']['frac'], j['roofline']['ms'], j['roofline']['traffic'], {k:(round(v['frac'],3),round(v['ms'],4),v['traffic']) for k,v in j['roofline_hbm_kernels'].items()}, j.get('cpu_baseline',{}).get('value'), j['clocks'])
PY
tail -2 gpurun_out/bench_r1_s.err
VBG_CUDA_GRAPHS=0 VBG_STREAMS=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_r1_s.csv \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/ncu_bench27.log 2>&1
