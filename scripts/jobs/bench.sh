#!/bin/bash
# default bench line (forward + e2e + train step + rooflines + library bar + CPU reference) and the reference arm
mkdir -p gpurun_out
( time timeout 900 python bench.py "$@" > gpurun_out/bench.json 2> gpurun_out/bench.err ) 2>&1 | grep real; echo "bench exit $?"
python - <<'PY'
import json
try:
    j = json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
    print({k: j[k] for k in ('value', 'ms_per_step', 'gpu_launches')}, 'e2e', j['e2e']['value'])
    print('roofline', j.get('roofline', {}).get('frac'), {k: (round(v['frac'], 3), round(v['ms'], 4)) for k, v in j.get('roofline_hbm_kernels', {}).items()})
    print('train', {k: j.get('train_step', {}).get(k) for k in ('value', 'ms_per_step', 'gpu_launches')})
    print('cpu', j.get('cpu_baseline'))
except Exception as e:
    print("no bench line:", e)
PY
tail -3 gpurun_out/bench.err
( time timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err ) 2>&1 | grep real
tail -c 900 gpurun_out/bench_ref.json
