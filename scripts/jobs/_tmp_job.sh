#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_autograd.py tests/test_gpu_train_step.py tests/test_gpu_train_graph.py -m gpu -q -x --timeout 600 2>&1 | grep -E "passed|failed|FAILED|Error|error|assert |^E " | tail -12
timeout 600 python bench.py --mode train --steps 15 --no-roofline --no-cpu-baseline --no-library-bar --no-input-pipeline > gpurun_out/r2_bench_train_n1_d.json 2> gpurun_out/r2_bench_train_n1_d.err; echo "bench exit $?"
python - <<PY
import json
j = json.loads(open('gpurun_out/r2_bench_train_n1_d.json').read().strip().splitlines()[-1])
print({k: j['train_step'].get(k) for k in ('value','ms_per_step','gpu_launches','peak_mem_gib')})
PY
