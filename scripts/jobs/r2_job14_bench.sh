#!/bin/bash
# r2 job 14: full default bench (forward + train step + rooflines + library bar + CPU reference) and the reference arm
mkdir -p gpurun_out
( time timeout 900 python bench.py > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err ) 2>&1 | grep real; echo "bench exit $?"
python - <<'PY'
import json
try:
    j=json.loads(open('gpurun_out/r2_bench_b.json').read().strip().splitlines()[-1])
    print({k:j[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', j['e2e']['value'])
    print('roofline', j['roofline']['frac'], {k:(round(v['frac'],3),round(v['ms'],4), v.get('back_to_back',{}).get('frac'), v.get('traffic')) for k,v in j['roofline_hbm_kernels'].items()})
    print('train', {k:j['train_step'].get(k) for k in ('value','ms_per_step','gpu_launches','whole_step_cuda_graph_replays')})
    print('library_bar', json.dumps(j.get('library_bar'))[:1500])
    print('cpu', j.get('cpu_baseline'))
except Exception as e:
    print("no bench line:", e)
PY
tail -3 gpurun_out/r2_bench_b.err
( time timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err ) 2>&1 | grep real
tail -c 900 gpurun_out/r2_bench_ref.json
