#!/bin/bash
# forward and training-step bench lines at N GPUs (gpurun --gpus N -- 'bash scripts/jobs/bench_n.sh N')
N=${1:-2}
mkdir -p gpurun_out
for mode in forward train; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus $N --steps 10 --warmup 3 --mode $mode --no-roofline --no-cpu-baseline > gpurun_out/bench_${mode}_n$N.json 2> gpurun_out/bench_${mode}_n$N.err; echo "bench $mode n$N exit $?"
  tail -c 1500 gpurun_out/bench_${mode}_n$N.json; tail -3 gpurun_out/bench_${mode}_n$N.err
done
