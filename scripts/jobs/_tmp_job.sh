#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_ops.py tests/test_gpu_train_step.py tests/test_gpu_train_graph.py -m gpu -q -x --timeout 600 2>&1 | grep -E "passed|failed|FAILED|Error|error|assert |^E " | tail -15
timeout 600 python bench.py --mode train --steps 15 --no-roofline --no-cpu-baseline --no-library-bar --no-input-pipeline > gpurun_out/r2_bench_train_n1_c.json 2> gpurun_out/r2_bench_train_n1_c.err; echo "bench exit $?"
python - <<PY
import json
j = json.loads(open('gpurun_out/r2_bench_train_n1_c.json').read().strip().splitlines()[-1])
print({k: j['train_step'].get(k) for k in ('value','ms_per_step','gpu_launches','peak_mem_gib')})
PY
VBG_TRAIN_GRAPHS=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'attn_bwd_tc_kernel' -c 4 -o gpurun_out/r2_ncu_attn_bwd_tc -f python scripts/train_bench.py cfg2 2 > gpurun_out/r2_ncu_attn_bwd_tc.log 2>&1; echo "ncu exit $?"
tail -2 gpurun_out/r2_ncu_attn_bwd_tc.log
