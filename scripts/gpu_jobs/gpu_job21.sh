#!/bin/bash
# GPU job 21: attention timeline; bench with pipelined result read + larger flush.
mkdir -p gpurun_out
timeout 120 python scripts/timeline.py 2>&1 | tail -14 > gpurun_out/timeline_attn.log; cat gpurun_out/timeline_attn.log
timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1_o.json 2> gpurun_out/bench_r1_o.err; echo "bench exit $?" >> gpurun_out/bench_r1_o.err
python - <<'PY'
import json
j=json.loads(open('gpurun_out/bench_r1_o.json').read().strip().splitlines()[-1])
print({k:j[k] for k in ('value','ms_per_step','gpu_launches')}, j['e2e'], j['roofline']['frac'], j['roofline']['ms'], {k:(round(v['frac'],3),round(v['ms'],4)) for k,v in j['roofline_hbm_kernels'].items()})
PY
tail -2 gpurun_out/bench_r1_o.err
