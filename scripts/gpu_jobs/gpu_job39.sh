#!/bin/bash
# GPU job 39: the other BASELINE configurations' shapes through the same bench (information only; cfg2 is the headline).
mkdir -p gpurun_out
for c in cfg1 cfg4 cfg5; do
  timeout 300 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/bench_r1_$c.json 2> gpurun_out/bench_r1_$c.err; echo "$c exit $?"
  python - gpurun_out/bench_r1_$c.json <<'PY'
import json,sys
try:
    j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(j['config']['workload']); print({k:j[k] for k in ('value','ms_per_step','gpu_launches')}, j['e2e']['value'])
except Exception as e: print("ERR", e)
PY
done
tail -3 gpurun_out/bench_r1_cfg4.err
