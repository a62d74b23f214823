#!/bin/bash
# GPU job 15: after the warp-uniform epilogue fix: all gpu tests (per-test timeout), PDL on/off, ROI windowed vs direct, launch list.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s --timeout 100 2>&1 | grep -E "^\[cfg|^\[|passed|failed|Error|assert |mismatch|Timeout" | tail -40 > gpurun_out/pytest_gpu.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
VBG_PDL=0 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1_j_nopdl.json 2> gpurun_out/bench_r1_j_nopdl.err
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1_j.json 2> gpurun_out/bench_r1_j.err; echo "bench exit $?" >> gpurun_out/bench_r1_j.err
VBG_ROI_DIRECT=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1_j_roidirect.json 2> /dev/null
for f in gpurun_out/bench_r1_j_nopdl.json gpurun_out/bench_r1_j.json gpurun_out/bench_r1_j_roidirect.json; do python - $f <<'PY'
import json,sys
try:
    j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], {k:j[k] for k in ('value','ms_per_step','gpu_launches')}, j['e2e']['value'], j['roofline']['frac'], j['roofline']['ms'], {k:(round(v['frac'],3),round(v['ms'],4)) for k,v in j['roofline_hbm_kernels'].items()})
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
tail -3 gpurun_out/bench_r1_j.err
VBG_CUDA_GRAPHS=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_r1_j.csv \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/ncu_bench15.log 2>&1
