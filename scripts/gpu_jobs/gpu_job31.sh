#!/bin/bash
# GPU job 31: batched normalise, seg logits on tensor cores, warm roofline figure: tests + bench + launch list.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s --timeout 100 2>&1 | grep -E "^\[cfg|passed|failed|Error|assert |mismatch|Timeout" | tail -8 > gpurun_out/pytest_gpu.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log; tail -7 gpurun_out/pytest_gpu.log
timeout 500 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1_v.json 2> gpurun_out/bench_r1_v.err; echo "bench exit $?" >> gpurun_out/bench_r1_v.err
python - <<'PY'
import json
j=json.loads(open('gpurun_out/bench_r1_v.json').read().strip().splitlines()[-1])
print({k:j[k] for k in ('value','ms_per_step','gpu_launches')}, j['e2e'], j['roofline']['frac'], j['roofline']['ms'], j['roofline'].get('warm_l2'), {k:(round(v['frac'],3),round(v['ms'],4)) for k,v in j['roofline_hbm_kernels'].items()})
PY
tail -3 gpurun_out/bench_r1_v.err
VBG_CUDA_GRAPHS=0 VBG_STREAMS=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_r1_v.csv \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/ncu_bench31.log 2>&1
