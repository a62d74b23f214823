#!/bin/bash
mkdir -p gpurun_out
for ch in 1 2; do
VBG_ARENA_CHUNKS=$ch timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2971$ch bench.py --gpus 8 --steps 10 --warmup 3 --mode train --no-roofline --no-cpu-baseline --no-input-pipeline > gpurun_out/r2_bench_train_n8_ch$ch.json 2> gpurun_out/r2_bench_train_n8_ch$ch.err; echo "exit $?"
python - <<PY
import json
j = json.loads(open('gpurun_out/r2_bench_train_n8_ch$ch.json').read().strip().splitlines()[-1])
print('chunks=$ch', {k: j.get('train_step', {}).get(k) for k in ('value', 'ms_per_step', 'allreduce')})
PY
done
