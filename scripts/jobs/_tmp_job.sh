#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_graph.py tests/test_gpu_train_step.py -m gpu -q -x --timeout 600 2>&1 | grep -E "passed|failed|FAILED|Error|error|assert |^E " | tail -15
timeout 600 python bench.py --config cfg5 --steps 10 --no-roofline --no-cpu-baseline --no-library-bar --no-serving --no-input-pipeline > gpurun_out/r2_bench_cfg5.json 2> gpurun_out/r2_bench_cfg5.err; echo "bench cfg5 exit $?"
python - <<PY
import json
j = json.loads(open('gpurun_out/r2_bench_cfg5.json').read().strip().splitlines()[-1])
print('cfg5', {k: j[k] for k in ('value','ms_per_step')}, 'train', {k: j.get('train_step',{}).get(k) for k in ('value','ms_per_step','whole_step_cuda_graph_replays','error')})
PY
tail -3 gpurun_out/r2_bench_cfg5.err
